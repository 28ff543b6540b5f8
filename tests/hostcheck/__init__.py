"""TEST INFRASTRUCTURE ONLY: g++ build of the __host__ __device__ pair function of the QuickPol
CUDA kernel (csrc/psb200_quickpol.cuh), so the CPU suite can check its arithmetic against the
oracle without a GPU.  The product never loads this."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "qp_host.cpp")
_HDR = os.path.join(_HERE, "..", "..", "powerspectra.jl_b200", "csrc", "psb200_quickpol.cuh")
_OUT = os.path.join(_HERE, "_build", "libqphost.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        newest = max(os.path.getmtime(_SRC), os.path.getmtime(_HDR))
        if not os.path.exists(_OUT) or os.path.getmtime(_OUT) < newest:
            os.makedirs(os.path.dirname(_OUT), exist_ok=True)
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared",
                            "-o", _OUT, _SRC, "-lm"], check=True, capture_output=True)
        L = C.CDLL(_OUT)
        dp = C.POINTER(C.c_double)
        L.qp_host_xi.argtypes = [C.c_int] * 6 + [dp, C.c_int, C.c_int, C.c_int, dp, C.c_long]
        L.qp_host_pair.argtypes = [C.c_int] * 7 + [dp, C.c_int]
        L.qp_host_pair.restype = C.c_double
        _lib = L
    return _lib


VARIANTS = {"simple": 0, "tab": 1}


def xi_band(nu1, nu2, s1, s2, lmax, W, band_lo, band_hi, variant="tab"):
    W = np.ascontiguousarray(W, dtype=np.float64)
    nb = band_lo + band_hi + 1
    Xb = np.zeros((nb, lmax + 1), order="F")
    dp = C.POINTER(C.c_double)
    lib().qp_host_xi(VARIANTS[variant], nu1, nu2, s1, s2, lmax, W.ctypes.data_as(dp), W.size, band_lo, band_hi,
                     Xb.ctypes.data_as(dp), nb)
    return Xb


def pair(l, lpp, nu1, nu2, s1, s2, W, variant="tab"):
    W = np.ascontiguousarray(W, dtype=np.float64)
    return lib().qp_host_pair(VARIANTS[variant], l, lpp, nu1, nu2, s1, s2, W.ctypes.data_as(C.POINTER(C.c_double)), W.size)
