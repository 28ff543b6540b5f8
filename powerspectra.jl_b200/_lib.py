"""ctypes binding of libpsb200.so (include/psb200.h).

The library is the product; this module only loads it and declares the prototypes.
There is no fallback of any kind: if the shared object is missing, or no CUDA device
is visible at call time, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PSB200_LIB lets a developer A/B another in-tree build of the same library (tools/); default: the product
LIB_PATH = os.environ.get("PSB200_LIB") or os.path.join(_HERE, "libpsb200.so")

DP = C.POINTER(C.c_double)
DPP = C.POINTER(DP)

# symbol -> (restype, argtypes); must list every function include/psb200.h declares
PROTOTYPES = {
    "psb200_mcm": (C.c_int, [C.c_int, C.c_int, C.c_int, DP, C.c_int, DP, C.c_long, DP, C.c_int]),
    "psb200_cov": (C.c_int, [C.c_int, C.c_int, C.c_int, DPP, C.c_int, DPP, C.c_int, DPP, C.c_int, C.c_int,
                             DP, C.c_long, C.c_int]),
    "psb200_mcm_master": (C.c_int, [C.c_int, C.c_int, DP, DP, DP, DP, C.c_int, DP, DP, DP, DP, DP, C.c_long, C.c_int]),
    "psb200_mcm_master_dev": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.POINTER(C.c_void_p), C.c_long, C.c_int, C.c_int, C.c_void_p]),
    "psb200_last_error": (C.c_char_p, []),
    "psb200_device_count": (C.c_int, []),
    "psb200_version": (C.c_char_p, []),
    "psb200_mcm_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_long,
                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "psb200_cov_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                 C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                 C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]),
    "psb200_mcm_dev_bands": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_long,
                                       C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "psb200_cov_dev_bands": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                       C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                       C.c_void_p, C.c_long, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "psb200_finish_dev": (C.c_int, [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "psb200_band_edges": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "psb200_host_band_edges": (C.c_int, [C.c_int] * 6 + [C.POINTER(C.c_int)]),
    "psb200_terms": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "psb200_dfma_peak": (C.c_double, [C.c_int]),
    "psb200_job_stats": (C.c_int, [C.c_int] * 6 + [C.POINTER(C.c_longlong)]),
    "psb200_quickpol_xi": (C.c_int, [C.c_int] * 5 + [DP, C.c_int, C.c_int, C.c_int, DP, C.c_long, C.c_int]),
    "psb200_quickpol_xi_dev": (C.c_int, [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long,
                                                         C.c_int, C.c_int, C.c_void_p]),
    "psb200_mcm_solve": (C.c_int, [C.c_int, C.c_int, C.c_int, DP, C.c_int, DP, C.c_long, C.c_int, DP, C.c_long, C.c_int]),
    "psb200_master_solve": (C.c_int, [C.c_int, C.c_int, DP, DP, DP, DP, C.c_int, DP, C.c_long, DP, C.c_long, C.c_int]),
    "psb200_decouple_covmat": (C.c_int, [C.c_int, DP, C.c_long, DP, C.c_long, DP, C.c_long, DP, C.c_long]),
    "psb200_decouple_covmat_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                                             C.c_void_p]),
    "psb200_zonal_alm": (C.c_int, [C.c_int, C.c_int, DP, DP, DP, C.c_long, C.c_int, DP, C.c_long]),
    "psb200_quickpol_edges": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "psb200_map2alm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, DPP, C.c_double, DP]),
    "psb200_map2alm_many": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, DPP, C.c_int, C.POINTER(C.c_int), DP, DPP, C.c_int]),
    "psb200_alm2map": (C.c_int, [C.c_int, C.c_int, DP, DP]),
    "psb200_alm2cl": (C.c_int, [C.c_int, DP, DP, DP]),
    "psb200_map2alm_dev": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psb200_alm2map_dev": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psb200_alm2cl_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "psb200_sht_stats": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "psb200_sht_release": (C.c_int, []),
    "psb200_host_alloc": (C.c_void_p, [C.c_size_t, C.c_int]),
    "psb200_host_free": (C.c_int, [C.c_void_p]),
    "psb200_host_placement": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int]),
    "psb200_host_numa_nodes": (C.c_int, []),
    "psb200_selftest_delivery": (C.c_int, [C.c_int] * 11 + [DPP, C.c_long]),
}


class PSB200Error(RuntimeError):
    """Non-zero return of a libpsb200 call (codes 2..5 of include/psb200.h)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libpsb200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libpsb200.so (built in-tree by csrc/build.sh / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with powerspectra.jl_b200/csrc/build.sh "
                "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    """Map a return code to the reference's error behaviour: 1 -> ValueError (the Python
    counterpart of Julia's ArgumentError / failed @assert, src/modecoupling.jl:80,225)."""
    if rc == 0:
        return
    msg = lib().psb200_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)
    if rc == 6:
        import numpy as np
        raise np.linalg.LinAlgError(msg)       # Julia: LinearAlgebra.SingularException
    raise PSB200Error(rc, msg)


class HostMatrix:
    """An N x N Float64 result array in page-locked host memory of the library (psb200_host_alloc), column-major like
    parent(SpectralArray).  `interleave=True` spreads its 2 MB pieces over the NUMA nodes of the host, which is what
    a several-GPU host call wants (every GPU writes its own region by DMA).  `.array` is the numpy view; the memory is
    released by `.free()` or when the object dies -- views taken from `.array` must not outlive it."""

    def __init__(self, n: int, interleave: bool = False):
        import numpy as np
        L = lib()
        self._p = L.psb200_host_alloc(int(n) * int(n) * 8, 1 if interleave else 0)
        if not self._p:
            raise PSB200Error(4, L.psb200_last_error().decode("utf-8", "replace"))
        self.array = np.ctypeslib.as_array(C.cast(self._p, DP), shape=(n, n)).T      # Fortran order over the block

    def placement(self, maxnodes: int = 8):
        """Sampled 2 MB pieces per NUMA node, or None where the kernel does not tell."""
        cnt = (C.c_int * maxnodes)()
        seen = lib().psb200_host_placement(self._p, cnt, maxnodes)
        return list(cnt) if seen > 0 else None

    def free(self):
        if self._p:
            self.array = None
            lib().psb200_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
