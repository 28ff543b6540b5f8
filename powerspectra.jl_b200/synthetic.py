"""Synthetic inputs of the benchmark configurations.

No SHT library exists in the build image, so the masks are *zonal* (azimuthally symmetric):
their a_lm reduce to a_l0, obtained by Gauss-Legendre quadrature of w(theta) P_l(cos theta).
Each mask removes a "galactic" band around the equator plus a one-sided polar cap (an
equator-symmetric mask would have a_l0 = 0 for odd l), C2-apodised like the reference's own
test masks (/root/reference/test/data/generate_mask.jl:24,43).  Host-side numpy only.
"""
from __future__ import annotations

import numpy as np

from .covariance import NULL, CovarianceWorkspace
from .spectral import SpectralVector


def gauss_legendre(n: int):
    """Nodes and weights of n-point Gauss-Legendre quadrature on [-1, 1] (Newton on P_n)."""
    k = np.arange(1, n + 1)
    x = np.cos(np.pi * (k - 0.25) / (n + 0.5))
    for _ in range(6):
        p0, p1 = np.ones_like(x), x.copy()
        for l in range(2, n + 1):
            p0, p1 = p1, ((2 * l - 1) * x * p1 - (l - 1) * p0) / l
        dp = n * (x * p1 - p0) / (x * x - 1.0)
        dx = p1 / dp
        x -= dx
        if np.max(np.abs(dx)) < 1e-15:
            break
    p0, p1 = np.ones_like(x), x.copy()
    for l in range(2, n + 1):
        p0, p1 = p1, ((2 * l - 1) * x * p1 - (l - 1) * p0) / l
    dp = n * (x * p1 - p0) / (x * x - 1.0)
    w = 2.0 / ((1.0 - x * x) * dp * dp)
    return x[::-1].copy(), w[::-1].copy()


def _c2(x):
    x = np.clip(x, 0.0, 1.0)
    return x - np.sin(2.0 * np.pi * x) / (2.0 * np.pi)


def mask_profile(theta, seed: int):
    """w(theta) in [0,1]: equatorial band of half-width b and polar cap theta < c removed,
    C2 transitions of width a; (b, c, a) drawn from numpy.random.default_rng(seed)."""
    rng = np.random.default_rng(seed)
    b = rng.uniform(0.15, 0.35)
    c = rng.uniform(0.3, 0.6)
    a = np.deg2rad(rng.uniform(2.0, 5.0))
    return _c2((np.abs(theta - np.pi / 2) - b) / a) * _c2((theta - c) / a)


class ZonalSky:
    """Quadrature grid + Legendre transform of zonal fields up to lmax."""

    def __init__(self, lmax: int, nnodes: int | None = None):
        self.lmax = lmax
        n = nnodes or (2 * lmax + 64)
        self.x, self.w = gauss_legendre(n)
        self.theta = np.arccos(self.x)

    def al0(self, fields):
        """a_l0 = sqrt(pi (2l+1)) int f(x) P_l(x) dx for each row of `fields`."""
        F = np.atleast_2d(np.asarray(fields, dtype=np.float64)) * self.w
        out = np.empty((F.shape[0], self.lmax + 1))
        p0, p1 = np.ones_like(self.x), self.x.copy()
        out[:, 0] = F @ p0
        if self.lmax >= 1:
            out[:, 1] = F @ p1
        for l in range(2, self.lmax + 1):
            p0, p1 = p1, ((2 * l - 1) * self.x * p1 - (l - 1) * p0) / l
            out[:, l] = F @ p1
        ls = np.arange(self.lmax + 1)
        return out * np.sqrt(np.pi * (2 * ls + 1))

    def al0_device(self, fields):
        """The same transform on the GPU (psb200_zonal_alm: W-spectrum production, first slice -- SURVEY.md 8f-4)."""
        from . import _lib
        F = np.ascontiguousarray(np.atleast_2d(np.asarray(fields, dtype=np.float64)))
        out = np.zeros((F.shape[0], self.lmax + 1))
        dp = lambda a: a.ctypes.data_as(_lib.DP)
        x, w = np.ascontiguousarray(self.x), np.ascontiguousarray(self.w)
        _lib.check(_lib.lib().psb200_zonal_alm(F.shape[0], F.shape[1], dp(x), dp(w), dp(F), F.shape[1], self.lmax, dp(out),
                                               self.lmax + 1))
        return out

    def cross(self, a, b):
        """alm2cl of two zonal fields: a_l0 b_l0 / (2l+1)."""
        return a * b / (2.0 * np.arange(self.lmax + 1) + 1.0)


def mask_spectra(lmax: int, seeds=(1001, 1002)):
    """V^{kl}_l for the masks `seeds`: dict (k, l) -> ndarray(lmax+1), k <= l as positions."""
    sky = ZonalSky(lmax)
    al = sky.al0([mask_profile(sky.theta, s) for s in seeds])
    return {(i, j): sky.cross(al[i], al[j]) for i in range(len(seeds)) for j in range(i, len(seeds))}


def theory_spectra(lmax: int):
    """Analytic CMB-like signal spectra (TT, EE > 0; TE changes sign) and noise ratios."""
    l = np.arange(lmax + 1, dtype=np.float64)
    le = np.maximum(l, 1.0)
    tt = 6000.0 * 2.0 * np.pi / (le * (le + 1.0)) * np.exp(-((l / 1500.0) ** 2)) + 1e-5
    ee = 0.02 * tt
    te = 0.1 * tt * np.cos(l / 90.0)
    r_auto = np.sqrt(1.0 + (l / 2000.0) ** 2)
    return {"TT": tt, "EE": ee, "TE": te, "r_auto": r_auto, "r_cross": np.ones_like(l)}


def covariance_inputs(lmax: int, seeds=(1001, 1002, 1003, 1004), names=("A", "B")):
    """Workspace + spectra + noise ratios for two fields A, B with a T and a P mask each,
    used as CovarianceWorkspace(A, B, A, B) like /root/reference/test/test_covmat.jl:80-82 so
    that every window spectrum is populated (a noise-weighted term is non-zero only when the
    two field names agree, src/workspace.jl:158-170).

    Returns (workspace, spectra, noiseratios); the workspace computes window spectra lazily
    with the zonal quadrature (the host 'SHT' of this synthetic sky).
    """
    sky = ZonalSky(lmax)
    th = sky.theta
    prof = {
        (names[0], "TT"): mask_profile(th, seeds[0]), (names[0], "PP"): mask_profile(th, seeds[1]),
        (names[1], "TT"): mask_profile(th, seeds[2]), (names[1], "PP"): mask_profile(th, seeds[3]),
    }
    nside = max((lmax + 1) // 3, 1)
    omega_pix = 4.0 * np.pi / (12.0 * nside * nside)
    c2 = np.cos(th) ** 2
    sigma2 = {"II": 1.0 + 0.5 * c2, "QQ": 2.0 + 0.8 * c2, "UU": 2.0 + 1.1 * c2}
    alm_cache = {}

    def eff_alm(A, i, j, alpha):
        # effective_weight_alm! (src/workspace.jl:141-171)
        key = (A, i, j, alpha)
        if key not in alm_cache:
            X, Y = alpha[0] * 2, alpha[1] * 2
            f = prof[i, X] * prof[j, Y]
            if A == NULL:
                alm_cache[key] = sky.al0(f)[0]
            elif i == j:
                alm_cache[key] = sky.al0(f * sigma2[A] * omega_pix)[0]
            else:
                alm_cache[key] = np.zeros(lmax + 1)
        return alm_cache[key]

    terms = {"TT": ("II",), "PP": ("QQ", "UU"), NULL: (NULL,)}

    def provider(X, Y, i, j, alpha, p, q, beta):
        # window_function_W! (src/workspace.jl:174-213)
        res = np.zeros(lmax + 1)
        for wx in terms[X]:
            for wy in terms[Y]:
                res += sky.cross(eff_alm(wx, i, j, alpha), eff_alm(wy, p, q, beta))
        return SpectralVector(res / (len(terms[X]) * len(terms[Y])))

    ws = CovarianceWorkspace((names[0], names[1], names[0], names[1]), lmax, provider=provider)
    th_sp = theory_spectra(lmax)
    spectra, ratios = {}, {}
    for a in names:
        for b in names:
            for s in ("TT", "EE", "TE"):
                spectra[s, a, b] = SpectralVector(th_sp[s])
            for s in ("TT", "EE"):
                ratios[s, a, b] = SpectralVector(th_sp["r_auto"] if a == b else th_sp["r_cross"])
    return ws, spectra, ratios
