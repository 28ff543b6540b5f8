"""ctypes front-end of the CPU oracle (oracle/psoracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the CPU
arm of bench.py; never by the product package.  See oracle/psoracle_impl.h for the
reference file:line map of every function.

All matrices are returned as numpy arrays A[l1 - lmin, l2 - lmin] (the memory handed
to C is column-major like a Julia `parent(SpectralArray)`; the results are symmetric
up to the (2l+1) factors so the layout is made explicit with order="F").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpsoracle.so")

MCM_KINDS = {"M00": 0, "M02": 1, "Mpp": 2, "Mmm": 3}
COV_BLOCKS = {"TTTT": 0, "EEEE": 1, "TTTE": 2, "TETE": 3, "TEEE_planck": 4, "TEEE": 5, "TTEE": 6}
COV_NEED = {  # block -> (n spectra, n ratios, n W)
    0: (4, 4, 8), 1: (4, 4, 8), 2: (4, 2, 4), 3: (4, 2, 5), 4: (4, 2, 4), 5: (4, 2, 4), 6: (4, 0, 2),
}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O3 -fopenmp)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("psoracle.c", "psoracle_impl.h", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None
_native = False


def use_native_build() -> str:
    """Timed CPU arm of bench.py only: (re)build the oracle with -O3 -march=native ON THIS MACHINE
    (`make native`) and load that build instead of the portable one.  Must be called
    before the first use of the library in the process."""
    global _native
    if _lib is not None and not _native:
        raise RuntimeError("use_native_build() must come before the first oracle call")
    subprocess.run(["make", "-C", _HERE, "native"], check=True, capture_output=True)
    _native = True
    return os.path.join(_HERE, "_build", "libpsoracle_native.so")


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_build", "libpsoracle_native.so") if _native else _LIB_PATH
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        dpp = C.POINTER(dp)
        ip = C.POINTER(C.c_int)
        for suf in ("", "_ld"):
            f = getattr(L, "pso_mcm" + suf)
            f.restype = C.c_longlong
            f.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, C.c_long, C.c_int, C.c_int]
            f = getattr(L, "pso_cov" + suf)
            f.restype = C.c_longlong
            f.argtypes = [C.c_int, C.c_int, C.c_int, dpp, C.c_int, dpp, C.c_int, dpp, C.c_int, C.c_int,
                          dp, C.c_long, C.c_int, C.c_int]
            f = getattr(L, "pso_mcm_rows" + suf)
            f.restype = C.c_longlong
            f.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, C.c_long, ip, C.c_int]
            f = getattr(L, "pso_cov_rows" + suf)
            f.restype = C.c_longlong
            f.argtypes = [C.c_int, C.c_int, C.c_int, dpp, C.c_int, dpp, C.c_int, dpp, C.c_int, C.c_int,
                          dp, C.c_long, ip, C.c_int]
            f = getattr(L, "pso_w3j_family" + suf)
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, ip, ip]
            f = getattr(L, "pso_quickpol_xi" + suf)
            f.restype = C.c_longlong
            f.argtypes = [C.c_int] * 5 + [dp, C.c_int, C.c_int, C.c_int, dp, C.c_long]
        L.pso_set_abs_mode.argtypes = [C.c_int]
        L.pso_max_threads.restype = C.c_int
        L.pso_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _vecs(vs):
    arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in vs]
    ptrs = (C.POINTER(C.c_double) * max(len(arrs), 1))(*[_dp(a) for a in arrs])
    return arrs, ptrs


def w3j_family(l1, l2, m2, m3, ld=False):
    """f(j) = (j l1 l2; -m2-m3 m2 m3), j = nmin..nmax.  Returns (nmin, values)."""
    n = 2 * min(l1, l2) + 1
    out = np.zeros(max(n, 1))
    nmin, nmax = C.c_int(), C.c_int()
    f = lib().pso_w3j_family_ld if ld else lib().pso_w3j_family
    k = f(l1, l2, m2, m3, _dp(out), out.size, C.byref(nmin), C.byref(nmax))
    if k < 0:
        raise ValueError("w3j_family: buffer too small")
    return nmin.value, out[:k].copy()


def _rows(rows):
    a = np.ascontiguousarray(rows, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def mcm(kind, lmin, lmax, V, ld=False, row0=0, rstep=1, threads=None, return_terms=False, rows=None):
    """inner_mcm00!/02!/++!/--! (src/modecoupling.jl:78-159).  kind: 0..3 or a name.
    rows: explicit list of l1 rows to evaluate (default: lmin+row0, +rstep, ...)."""
    kind = MCM_KINDS.get(kind, kind)
    V = np.ascontiguousarray(V, dtype=np.float64)
    N = lmax - lmin + 1
    M = np.zeros((N, N), order="F")
    if threads:
        lib().pso_set_threads(int(threads))
    if rows is not None:
        ra, rp = _rows(rows)
        f = lib().pso_mcm_rows_ld if ld else lib().pso_mcm_rows
        t = f(kind, lmin, lmax, _dp(V), V.size, _dp(M), N, rp, ra.size)
    else:
        f = lib().pso_mcm_ld if ld else lib().pso_mcm
        t = f(kind, lmin, lmax, _dp(V), V.size, _dp(M), N, row0, rstep)
    if t < 0:
        raise ValueError("oracle mcm: bad arguments")
    return (M, t) if return_terms else M


def cov(block, lmin, lmax, spectra, ratios, W, ld=False, row0=0, rstep=1, threads=None, return_terms=False, rows=None):
    """loop_cov*! (src/covariance.jl:92-446); positional order of the reference."""
    block = COV_BLOCKS.get(block, block)
    sa, sp = _vecs(spectra)
    ra, rp = _vecs(ratios)
    wa, wp = _vecs(W)
    lenW = min(w.size for w in wa)
    for a in sa + ra:
        if a.size < lmax + 1:
            raise ValueError("spectrum / ratio shorter than lmax+1")
    N = lmax - lmin + 1
    Cm = np.zeros((N, N), order="F")
    if threads:
        lib().pso_set_threads(int(threads))
    if rows is not None:
        rowa, rowp = _rows(rows)
        f = lib().pso_cov_rows_ld if ld else lib().pso_cov_rows
        t = f(block, lmin, lmax, sp, len(sa), rp, len(ra), wp, len(wa), lenW, _dp(Cm), N, rowp, rowa.size)
    else:
        f = lib().pso_cov_ld if ld else lib().pso_cov
        t = f(block, lmin, lmax, sp, len(sa), rp, len(ra), wp, len(wa), lenW, _dp(Cm), N, row0, rstep)
    if t < 0:
        raise ValueError("oracle cov: bad arguments")
    return (Cm, t) if return_terms else Cm


def band_to_dense(Xb, lmax, band_lo, band_hi):
    """BandedMatrices storage Xb[band_hi + i - j, j] -> dense A[i, j] (zeros off the band)."""
    n = lmax + 1
    A = np.zeros((n, n))
    for j in range(n):
        i0, i1 = max(0, j - band_hi), min(n - 1, j + band_lo)
        A[i0:i1 + 1, j] = Xb[band_hi + i0 - j:band_hi + i1 - j + 1, j]
    return A


def quickpol_xi(nu1, nu2, s1, s2, lmax, W, band_lo, band_hi, ld=False, threads=None, dense=True,
                return_terms=False):
    """quickpolXi! (src/beam.jl:72-101).  Returns Xi[l'', l] (dense, zeros off the band and for
    l'' < 2 or l < 2) or, with dense=False, the (band_lo+band_hi+1) x (lmax+1) band storage."""
    W = np.ascontiguousarray(W, dtype=np.float64)
    nb = band_lo + band_hi + 1
    Xb = np.zeros((nb, lmax + 1), order="F")
    if threads:
        lib().pso_set_threads(int(threads))
    f = lib().pso_quickpol_xi_ld if ld else lib().pso_quickpol_xi
    t = f(nu1, nu2, s1, s2, lmax, _dp(W), W.size, band_lo, band_hi, _dp(Xb), nb)
    if t < 0:
        raise ValueError("oracle quickpol_xi: bad arguments")
    out = band_to_dense(Xb, lmax, band_lo, band_hi) if dense else Xb
    return (out, t) if return_terms else out


class abs_mode:
    """Context manager: inside it mcm()/cov() return the condition sums S_abs (sum of |terms|)."""

    def __enter__(self):
        lib().pso_set_abs_mode(1)

    def __exit__(self, *a):
        lib().pso_set_abs_mode(0)


def max_threads() -> int:
    return lib().pso_max_threads()


def set_threads(n: int) -> int:
    """Force the OpenMP team size (overrides OMP_NUM_THREADS, which torchrun sets to 1)."""
    lib().pso_set_threads(int(n))
    return max_threads()


def host_cores() -> int:
    """CPUs this process may run on (what the reference's Julia threads would be started with)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
