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


# ---- host build of the device arithmetic of the spin-0 HEALPix transforms (csrc/psb200_sht.cuh) ----------------
_SHT_SRC = os.path.join(_HERE, "sht_host.cpp")
_SHT_HDR = os.path.join(_HERE, "..", "..", "powerspectra.jl_b200", "csrc", "psb200_sht.cuh")
_SHT_OUT = os.path.join(_HERE, "_build", "libshthost.so")
_sht = None


def sht_lib():
    global _sht
    if _sht is None:
        newest = max(os.path.getmtime(_SHT_SRC), os.path.getmtime(_SHT_HDR))
        if not os.path.exists(_SHT_OUT) or os.path.getmtime(_SHT_OUT) < newest:
            os.makedirs(os.path.dirname(_SHT_OUT), exist_ok=True)
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC",
                            "-shared", "-I/usr/local/cuda/include", "-o", _SHT_OUT, _SHT_SRC, "-lm"],
                           check=True, capture_output=True)
        L = C.CDLL(_SHT_OUT)
        dp = C.POINTER(C.c_double)
        L.sht_host_map2alm.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp]
        L.sht_host_alm2map.argtypes = [C.c_int, C.c_int, dp, dp]
        L.sht_host_lambda.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.sht_host_ring_analyse.argtypes = [dp, C.c_int, C.c_int, C.c_double, C.c_int, dp]
        L.sht_host_ring_synthesise.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp]
        L.sht_host_ring.argtypes = [C.c_int, C.c_int, dp]
        L.sht_host_ring.restype = None
        _sht = L
    return _sht


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def sht_map2alm(f, nside, lmax, niter=3):
    f = np.ascontiguousarray(f, dtype=np.float64)
    alm = np.zeros((lmax + 1) * (lmax + 2) // 2, dtype=np.complex128)
    sht_lib().sht_host_map2alm(nside, lmax, niter, _dp(f), alm.ctypes.data_as(C.POINTER(C.c_double)))
    return alm


def sht_alm2map(alm, nside, lmax):
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    f = np.zeros(12 * nside * nside)
    sht_lib().sht_host_alm2map(nside, lmax, alm.ctypes.data_as(C.POINTER(C.c_double)), _dp(f))
    return f


def sht_lambda(nside, lmax, m, p):
    lam = np.zeros(lmax - m + 1)
    alive = sht_lib().sht_host_lambda(nside, lmax, m, p, _dp(lam))
    return lam, alive


def sht_ring_analyse(f, shifted, scale, mmax):
    f = np.ascontiguousarray(f, dtype=np.float64)
    out = np.zeros(mmax + 1, dtype=np.complex128)
    sht_lib().sht_host_ring_analyse(_dp(f), f.size, int(shifted), scale, mmax, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def sht_ring_synthesise(F, n, shifted):
    F = np.ascontiguousarray(F, dtype=np.complex128)
    f = np.zeros(n)
    sht_lib().sht_host_ring_synthesise(F.ctypes.data_as(C.POINTER(C.c_double)), n, int(shifted), F.size - 1, _dp(f))
    return f


def sht_ring(nside, p):
    out = np.zeros(6)
    sht_lib().sht_host_ring(nside, p, _dp(out))
    return out
