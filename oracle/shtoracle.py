"""CPU statement of the spin-0 HEALPix transforms behind W-spectrum production.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the CPU arm of bench.py;
never by the product package.

What it restates.  The reference builds every window spectrum from
    effective_weight_alm!   /root/reference/src/workspace.jl:141-171   map2alm(mask_i .* mask_j [.* sigma^2 .* Omega_pix]; lmax)
    window_function_W!      /root/reference/src/workspace.jl:174-213   mean over (wX, wY) of alm2cl(w_X, w_Y)[0:lmax]
and `map2alm` / `alm2cl` live in Healpix.jl (Project.toml:14,34, compat "3, 4", no Manifest => no exact pin, source
not under /root/reference).  Published algorithm restated here (Gorski et al. 2005, ApJ 622, 759, section 5 and the
HEALPix C/Fortran `pix2ang_ring`, `map2alm_iterative`):
  * RING pixelisation: npix = 12 nside^2, rings r = 1..4 nside-1 with n_phi = 4 min(r, nside, 4 nside - r) pixels,
    z = 1 - r^2/(3 nside^2) (caps), z = 4/3 - 2r/(3 nside) (belt), first pixel at phi = pi/n_phi on the caps and on
    belt rings with r - nside even, at phi = 0 on the other belt rings;
  * analysis with uniform pixel weights 4 pi/npix,  a_lm = (4 pi/npix) sum_p f_p conj(Y_lm(p)),
    followed by `niter` Jacobi iterations  a <- a + A(f - S a)  (Healpix.jl `map2alm(map; lmax, mmax, niter = 3)`);
  * synthesis  f_p = sum_lm a_lm Y_lm(p);   Y_lm = lambda_lm(theta) e^{i m phi} with the Condon-Shortley phase;
  * alm2cl:  C_l = [a_l0 b_l0 + 2 sum_{m>0} Re(a_lm conj b_lm)]/(2l+1);  Alm storage m-major,
    index(l, m) = m (2 lmax + 1 - m)/2 + l.
The sums are evaluated DIRECTLY (every pixel with its own phi, no FFT, no aliasing argument, no symmetry) in long double,
so that the oracle shares no shortcut with the CUDA path (ring FFTs, north/south folding, scaled recurrences).

PARITY UNPINNED at the fixture level: no healpy / Healpix.jl / libsharp exists in this image and the reference tree holds
no map or alm fixture (.MISSING_LARGE_BLOBS).  What pins this file instead (tests/test_sht.py): lambda_lm against
scipy.special.sph_harm_y, the pixel centres against the closed-form low-nside values of the HEALPix paper, exact
recovery of band-limited fields, and agreement with the Gauss-Legendre zonal statement (synthetic.ZonalSky).
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble


def npix(nside: int) -> int:
    return 12 * nside * nside


def ring_table(nside: int):
    """Per ring r = 1..4 nside - 1: (n_phi, first pixel, z = cos theta, phi of the first pixel)."""
    N = int(nside)
    r = np.arange(1, 4 * N)
    rr = np.minimum(r, 4 * N - r)                      # distance from the nearer pole in rings
    cap = rr < N
    nphi = np.where(cap, 4 * rr, 4 * N)
    ncap = 2 * N * (N - 1)
    start = np.where(r < N, 2 * r * (r - 1),
                     np.where(r <= 3 * N, ncap + (r - N) * 4 * N, npix(N) - 2 * rr * (rr + 1)))
    z = np.where(cap, 1.0 - rr.astype(LD) ** 2 / (3 * LD(N) ** 2), LD(4) / 3 - 2 * r.astype(LD) / (3 * LD(N)))
    z = np.where((r > 3 * N), -z, z)
    shifted = cap | (((r - N) % 2) == 0)
    phi0 = np.where(shifted, LD(np.pi) * 0 + np.pi * LD(1) / nphi, LD(0))
    return nphi.astype(np.int64), start.astype(np.int64), z.astype(LD), phi0.astype(LD)


def pix2ang_ring(nside: int):
    """(theta, phi) of every pixel centre straight from the per-pixel formulas of the HEALPix paper / `pix2ang_ring`
    (independent of ring_table, which the tests compare it with)."""
    N = int(nside)
    n = npix(N)
    ncap = 2 * N * (N - 1)
    p = np.arange(n)
    z = np.zeros(n, dtype=LD)
    phi = np.zeros(n, dtype=LD)
    pi = LD(np.pi)
    north = p < ncap
    ph = (p[north] + 1) / 2.0
    i = np.floor(np.sqrt(ph - np.sqrt(np.floor(ph)))).astype(np.int64) + 1
    j = p[north] + 1 - 2 * i * (i - 1)
    z[north] = 1 - i.astype(LD) ** 2 / (3 * LD(N) ** 2)
    phi[north] = (j.astype(LD) - LD(0.5)) * pi / (2 * i.astype(LD))
    belt = (p >= ncap) & (p < n - ncap)
    ip = p[belt] - ncap
    i = ip // (4 * N) + N
    j = ip % (4 * N) + 1
    fodd = np.where(((i + N) & 1) == 1, LD(1), LD(0.5))
    z[belt] = (2 * N - i).astype(LD) * 2 / (3 * LD(N))
    phi[belt] = (j.astype(LD) - fodd) * pi / (2 * LD(N))
    south = p >= n - ncap
    ip = n - p[south]
    hip = ip / 2.0
    i = np.floor(np.sqrt(hip - np.sqrt(np.floor(hip)))).astype(np.int64) + 1
    j = 4 * i + 1 - (ip - 2 * i * (i - 1))
    z[south] = -(1 - i.astype(LD) ** 2 / (3 * LD(N) ** 2))
    phi[south] = (j.astype(LD) - LD(0.5)) * pi / (2 * i.astype(LD))
    return np.arccos(z), phi


def lam_rows(lmax: int, m: int, x):
    """lambda_lm(x) for l = m..lmax (rows) at every x = cos(theta) (columns), long double, plain upward recurrence
    lambda_l = (x lambda_{l-1} - A_{l-1} lambda_{l-2})/A_l,  A_l = sqrt((l^2 - m^2)/(4 l^2 - 1)),
    from lambda_mm = (-1)^m sqrt((2m+1)/(4 pi) prod_{k<=m} (2k-1)/(2k)) sin^m(theta)."""
    x = np.asarray(x, dtype=LD)
    s = np.sqrt((1 - x) * (1 + x))
    out = np.zeros((lmax - m + 1, x.size), dtype=LD)
    k = np.arange(1, m + 1, dtype=LD)
    c = np.sqrt((2 * LD(m) + 1) / (4 * LD(np.pi)) * np.prod((2 * k - 1) / (2 * k))) if m else np.sqrt(1 / (4 * LD(np.pi)))
    out[0] = (-1) ** m * c * s ** m
    if lmax > m:
        out[1] = x * np.sqrt(LD(2 * m + 3)) * out[0]
    for l in range(m + 2, lmax + 1):
        A = np.sqrt(LD(l * l - m * m) / LD(4 * l * l - 1))
        A1 = np.sqrt(LD((l - 1) ** 2 - m * m) / LD(4 * (l - 1) ** 2 - 1))
        out[l - m] = (x * out[l - m - 1] - A1 * out[l - m - 2]) / A
    return out


def alm_index(lmax: int, l, m):
    return m * (2 * lmax + 1 - m) // 2 + l


def alm_size(lmax: int) -> int:
    return (lmax + 1) * (lmax + 2) // 2


def _ring_phases(nside):
    """Per ring the long-double pixel longitudes (list of arrays)."""
    nphi, start, z, phi0 = ring_table(nside)
    return nphi, start, z, [phi0[r] + 2 * LD(np.pi) * np.arange(nphi[r], dtype=LD) / nphi[r] for r in range(nphi.size)]


def analysis(nside: int, lmax: int, f):
    """A(f): a_lm = (4 pi/npix) sum_p f_p lambda_lm(theta_p) e^{-i m phi_p}, direct sums, long double."""
    f = np.asarray(f, dtype=LD)
    nphi, start, z, phis = _ring_phases(nside)
    w = 4 * LD(np.pi) / npix(nside)
    are = np.zeros(alm_size(lmax), dtype=LD)
    aim = np.zeros(alm_size(lmax), dtype=LD)
    for m in range(lmax + 1):
        gre = np.array([np.sum(f[start[r]:start[r] + nphi[r]] * np.cos(m * phis[r])) for r in range(nphi.size)], dtype=LD)
        gim = np.array([-np.sum(f[start[r]:start[r] + nphi[r]] * np.sin(m * phis[r])) for r in range(nphi.size)], dtype=LD)
        lam = lam_rows(lmax, m, z)
        i0 = alm_index(lmax, m, m)
        are[i0:i0 + lmax - m + 1] = w * (lam @ gre)
        aim[i0:i0 + lmax - m + 1] = w * (lam @ gim)
    return are, aim


def synthesis(nside: int, lmax: int, are, aim):
    """S(a): f_p = sum_l [a_l0 lambda_l0 + 2 sum_{m>0} Re(a_lm e^{i m phi_p}) lambda_lm], direct sums, long double."""
    nphi, start, z, phis = _ring_phases(nside)
    f = np.zeros(npix(nside), dtype=LD)
    for m in range(lmax + 1):
        lam = lam_rows(lmax, m, z)
        i0 = alm_index(lmax, m, m)
        fre = lam.T @ np.asarray(are[i0:i0 + lmax - m + 1], dtype=LD)
        fim = lam.T @ np.asarray(aim[i0:i0 + lmax - m + 1], dtype=LD)
        fac = 1 if m == 0 else 2
        for r in range(nphi.size):
            f[start[r]:start[r] + nphi[r]] += fac * (fre[r] * np.cos(m * phis[r]) - fim[r] * np.sin(m * phis[r]))
    return f


def map2alm(f, nside: int, lmax: int, niter: int = 3, ld: bool = False):
    """Healpix.jl `map2alm(map; lmax, niter)`: pixel-weighted analysis plus `niter` Jacobi iterations.
    Returns complex128 (or a pair of long-double arrays with ld=True) in Alm order."""
    f = np.asarray(f, dtype=LD)
    are, aim = analysis(nside, lmax, f)
    for _ in range(niter):
        res = f - synthesis(nside, lmax, are, aim)
        dre, dim = analysis(nside, lmax, res)
        are, aim = are + dre, aim + dim
    if ld:
        return are, aim
    return are.astype(np.float64) + 1j * aim.astype(np.float64)


def alm2map(alm, nside: int, lmax: int, ld: bool = False):
    alm = np.asarray(alm)
    f = synthesis(nside, lmax, alm.real.astype(LD), alm.imag.astype(LD))
    return f if ld else f.astype(np.float64)


def alm2cl(a, b, lmax: int):
    """Healpix `alm2cl`: C_l = [a_l0 b_l0 + 2 sum_{m>0} Re(a_lm conj b_lm)]/(2l+1)."""
    a = np.asarray(a)
    b = np.asarray(b)
    cl = np.zeros(lmax + 1, dtype=LD)
    for m in range(lmax + 1):
        i0 = alm_index(lmax, m, m)
        sl = slice(i0, i0 + lmax - m + 1)
        pr = a.real[sl].astype(LD) * b.real[sl].astype(LD) + a.imag[sl].astype(LD) * b.imag[sl].astype(LD)
        cl[m:] += pr if m == 0 else 2 * pr
    return (cl / (2 * np.arange(lmax + 1, dtype=LD) + 1)).astype(np.float64)


def effective_weight_alm(nside, lmax, mask_i, mask_j, sigma2=None, niter: int = 3):
    """effective_weight_alm! (src/workspace.jl:141-171): map2alm of mask_i .* mask_j, times sigma^2 .* Omega_pix for the
    noise-weighted kinds (II, QQ, UU)."""
    f = np.asarray(mask_i, dtype=LD) * np.asarray(mask_j, dtype=LD)
    if sigma2 is not None:
        f = f * (4 * LD(np.pi) / npix(nside))          # parent(map_buffer) .*= parent(weight) .* Omega_p  (:161-162)
        f = f * np.asarray(sigma2, dtype=LD)
    return map2alm(f, nside, lmax, niter)


# ---- ring-based CPU restatement (oracle/shtcpu.c: C + OpenMP, the shape of the reference's libsharp path) ---------
_fast = None
_fast_native = False


def fast_lib(native: bool = False):
    """libshtcpu.so (portable build, tests) or, with native=True, the -O3 -march=native build made on this machine
    (bench.py's timed CPU arm)."""
    global _fast, _fast_native
    import ctypes as C
    import os
    import subprocess
    if _fast is None or native != _fast_native:
        here = os.path.dirname(os.path.abspath(__file__))
        subprocess.run(["make", "-C", here] + (["native"] if native else []), check=True, capture_output=True)
        L = C.CDLL(os.path.join(here, "_build", "libshtcpu_native.so" if native else "libshtcpu.so"))
        dp = C.POINTER(C.c_double)
        L.shtcpu_map2alm.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp]
        L.shtcpu_alm2map.argtypes = [C.c_int, C.c_int, dp, dp]
        L.shtcpu_threads.argtypes = [C.c_int]
        _fast, _fast_native = L, native
    return _fast


def fast_map2alm(f, nside: int, lmax: int, niter: int = 3, native: bool = False, threads: int = 0):
    import ctypes as C
    L = fast_lib(native)
    if threads:
        L.shtcpu_threads(threads)
    f = np.ascontiguousarray(f, dtype=np.float64)
    alm = np.zeros(alm_size(lmax), dtype=np.complex128)
    dp = C.POINTER(C.c_double)
    if L.shtcpu_map2alm(nside, lmax, niter, f.ctypes.data_as(dp), alm.ctypes.data_as(dp)):
        raise MemoryError("shtcpu_map2alm")
    return alm


def fast_alm2map(alm, nside: int, lmax: int, native: bool = False):
    import ctypes as C
    L = fast_lib(native)
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    f = np.zeros(npix(nside))
    dp = C.POINTER(C.c_double)
    if L.shtcpu_alm2map(nside, lmax, alm.ctypes.data_as(dp), f.ctypes.data_as(dp)):
        raise MemoryError("shtcpu_alm2map")
    return f
