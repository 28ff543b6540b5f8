"""QuickPol beam-matrix utilities, mirroring /root/reference/src/beam.jl.

`quickpolW` (:43-56) and `kᵤ` (:115-124) are host arithmetic exactly as in the reference; the pair
loop of `quickpolΞ!` (:72-101) -- two general-spin Wigner-3j families and one l' reduction per
stored (l'', l) -- runs on the GPU through `psb200_quickpol_xi` (include/psb200.h).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .modecoupling import Alm
from .spectral import SpectralVector


class BandedSpectralMatrix:
    """The reference's `SpectralArray(BandedMatrix(...))` over 0:lmax x 0:lmax
    (docs/src/beams.md): `band_lo` sub- and `band_hi` super-diagonals, stored exactly like
    BandedMatrices.jl stores them -- `data[band_hi + i - j, j] = A[i, j]`, column-major,
    (band_lo+band_hi+1) x (lmax+1) -- so `data` is what the C ABI reads and writes."""

    def __init__(self, lmax, band_lo, band_hi, data=None):
        self.lmax, self.band_lo, self.band_hi = int(lmax), int(band_lo), int(band_hi)
        if self.lmax < 0 or self.band_lo < 0 or self.band_hi < 0:
            raise ValueError("lmax and band widths must be non-negative")
        shape = (self.band_lo + self.band_hi + 1, self.lmax + 1)
        if data is None:
            data = np.zeros(shape, order="F")
        data = np.asarray(data, dtype=np.float64)
        if data.shape != shape:
            raise ValueError(f"band storage must have shape {shape}")
        self.data = np.asfortranarray(data)

    @property
    def shape(self):
        return (self.lmax + 1, self.lmax + 1)

    def inband(self, i, j):
        return 0 <= i <= self.lmax and 0 <= j <= self.lmax and -self.band_hi <= i - j <= self.band_lo

    def __getitem__(self, key):
        i, j = (int(k) for k in key)
        if not (0 <= i <= self.lmax and 0 <= j <= self.lmax):
            raise IndexError(f"multipoles ({i}, {j}) outside 0:{self.lmax}")
        return float(self.data[self.band_hi + i - j, j]) if self.inband(i, j) else 0.0

    def __setitem__(self, key, value):
        i, j = (int(k) for k in key)
        if not self.inband(i, j):
            raise IndexError(f"({i}, {j}) is outside the band")
        self.data[self.band_hi + i - j, j] = value

    def rowrange(self, r):
        """specrowrange (src/beam.jl:59-63): stored columns of row r, never below multipole 2."""
        return range(max(2, r - self.band_lo), min(self.lmax, r + self.band_hi) + 1)

    def todense(self):
        n = self.lmax + 1
        A = np.zeros((n, n), order="F")
        for j in range(n):
            i0, i1 = max(0, j - self.band_hi), min(n - 1, j + self.band_lo)
            A[i0:i1 + 1, j] = self.data[self.band_hi + i0 - j:self.band_hi + i1 - j + 1, j]
        return A

    def matvec(self, x):
        """A @ x for a 0-indexed vector (the `B * C_l` product of docs/src/beams.md)."""
        x = np.asarray(getattr(x, "parent", x), dtype=np.float64)
        if x.size != self.lmax + 1:
            raise ValueError("vector length must be lmax+1")
        y = np.zeros(self.lmax + 1)
        for k in range(-self.band_hi, self.band_lo + 1):          # diagonal i - j = k
            j0, j1 = max(0, -k), min(self.lmax, self.lmax - k)
            if j1 >= j0:
                js = np.arange(j0, j1 + 1)
                y[js + k] += self.data[self.band_hi + k, js] * x[js]
        return y


def quickpolW(alm1: Alm, alm2: Alm):
    """Scaled spectrum of the scan pattern, W_l' = sum_m a_{l'm} conj(b_{l'm}) over -l' <= m <= l'
    (src/beam.jl:43-56; no 1/(2l+1))."""
    lmax = min(alm1.lmax, alm2.lmax)
    mmax = min(alm1.mmax, alm2.mmax)
    cl = np.zeros(lmax + 1)
    for m in range(mmax + 1):
        i1 = m * (2 * alm1.lmax + 1 - m) // 2
        i2 = m * (2 * alm2.lmax + 1 - m) // 2
        ls = np.arange(m, lmax + 1)
        prod = (alm1.alm[i1 + ls] * np.conj(alm2.alm[i2 + ls])).real
        cl[ls] += prod if m == 0 else 2.0 * prod
    return SpectralVector(cl)


def quickpolXi(Xi: BandedSpectralMatrix, nu1, nu2, s1, s2, omega1, omega2=None, *, ngpus=1):
    """quickpolΞ!(𝚵, ν₁, ν₂, s₁, s₂, ω₁, ω₂) (src/beam.jl:72-101), in place.

    omega1 / omega2: effective scan-weight Alm's of spin s1+nu1 and s2+nu2; or pass the spectrum
    W = quickpolW(omega1, omega2) as `omega1` with omega2=None.
    Entries the reference loop does not visit keep their value times the final sign, as
    `𝚵 .*= sgn` (:98-99) does."""
    if not isinstance(Xi, BandedSpectralMatrix):
        raise ValueError("𝚵 must be a BandedSpectralMatrix (square, 0:lmax)")          # :78
    W = quickpolW(omega1, omega2) if omega2 is not None else omega1
    w = np.ascontiguousarray(getattr(W, "parent", W), dtype=np.float64)
    nu1, nu2, s1, s2 = int(nu1), int(nu2), int(s1), int(s2)
    sgn = -1.0 if (s1 + s2 + nu1 + nu2) % 2 else 1.0
    if sgn < 0:
        # the library overwrites every entry the loop visits (sign included), so scaling the whole band storage
        # first leaves exactly the unvisited entries (rows / columns below multipole 2) multiplied by sgn
        Xi.data *= sgn
    try:
        rc = _lib.lib().psb200_quickpol_xi(nu1, nu2, s1, s2, Xi.lmax, w.ctypes.data_as(_lib.DP), w.size,
                                           Xi.band_lo, Xi.band_hi, Xi.data.ctypes.data_as(_lib.DP),
                                           Xi.data.shape[0], ngpus)
        _lib.check(rc)
    except Exception:
        if sgn < 0:
            Xi.data *= sgn                      # a failed call leaves 𝚵 as it was
        raise
    return Xi


def k_u(u):
    """kᵤ (src/beam.jl:115-124): 1 for u = 0, 1/2 for |u| = 2."""
    if u == 0:
        return 1.0
    if abs(u) == 2:
        return 0.5
    raise ValueError("Defined only for u ∈ {-2, 0, 2}.")
