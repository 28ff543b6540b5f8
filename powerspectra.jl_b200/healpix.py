"""W-spectrum production on the GPU: the Healpix.jl calls behind
`effective_weight_alm!` / `window_function_W!` (/root/reference/src/workspace.jl:141-213) and the `map2alm(mask)`
in front of `mcm` (src/modecoupling.jl:250-256).

`HealpixMap`, `map2alm`, `alm2map`, `alm2cl` mirror the Healpix.jl names the reference imports
(src/PowerSpectra.jl:7-9); `CovField` and the map-based `CovarianceWorkspace` constructor mirror
src/workspace.jl:20-60, 77-135.  All transforms run in libpsb200.so (csrc/psb200_sht.cuh); nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .modecoupling import Alm


def nside2npix(nside: int) -> int:
    return 12 * int(nside) ** 2


def nside2lmax(nside: int) -> int:
    """`getlmax` heuristic of the reference (src/workspace.jl:95): 3 nside - 1."""
    return 3 * int(nside) - 1


class HealpixMap:
    """RING-ordered Float64 map (Healpix.HealpixMap{Float64,RingOrder}): `.pixels` is `parent(map)`."""

    def __init__(self, pixels):
        self.pixels = np.ascontiguousarray(pixels, dtype=np.float64)
        nside = int(round(np.sqrt(self.pixels.size / 12.0)))
        if 12 * nside * nside != self.pixels.size or nside < 1 or nside & (nside - 1):
            raise ValueError(f"{self.pixels.size} pixels is not a HEALPix map (12 nside^2, nside a power of two)")
        self.nside = nside

    @classmethod
    def zeros(cls, nside):
        return cls(np.zeros(nside2npix(nside)))

    def __len__(self):
        return self.pixels.size


class PolarizedHealpixMap:
    """(i, q, u) triple of maps, the shape of the reference's sigma^2 argument (src/workspace.jl:20-31)."""

    def __init__(self, i, q, u):
        self.i, self.q, self.u = (m if isinstance(m, HealpixMap) else HealpixMap(m) for m in (i, q, u))


def _dp(a):
    return a.ctypes.data_as(_lib.DP)


def _map2alm_product(factors, scale, lmax, niter):
    maps = [f.pixels if isinstance(f, HealpixMap) else np.ascontiguousarray(f, dtype=np.float64) for f in factors]
    nside = HealpixMap(maps[0]).nside
    if any(m.size != maps[0].size for m in maps):
        raise ValueError("maps of different resolution")
    lmax = nside2lmax(nside) if lmax is None else int(lmax)
    alm = np.zeros((lmax + 1) * (lmax + 2) // 2, dtype=np.complex128)
    ptrs = (_lib.DP * len(maps))(*[_dp(m) for m in maps])
    _lib.check(_lib.lib().psb200_map2alm(nside, lmax, int(niter), len(maps), ptrs, float(scale),
                                         alm.ctypes.data_as(_lib.DP)))
    return Alm(lmax, lmax, alm)


def map2alm(m, *, lmax=None, niter: int = 3) -> Alm:
    """Healpix.jl `map2alm(map; lmax, niter = 3)`: pixel-weighted analysis + `niter` Jacobi iterations."""
    return _map2alm_product([m], 1.0, lmax, niter)


def alm2map(alm: Alm, nside: int) -> HealpixMap:
    if alm.mmax != alm.lmax:
        raise ValueError("alm2map needs mmax == lmax")
    a = np.ascontiguousarray(alm.alm, dtype=np.complex128)
    out = np.zeros(nside2npix(nside))
    _lib.check(_lib.lib().psb200_alm2map(int(nside), alm.lmax, a.ctypes.data_as(_lib.DP), _dp(out)))
    return HealpixMap(out)


def alm2cl_device(a: Alm, b: Alm | None = None) -> np.ndarray:
    """Healpix.jl `alm2cl(a, b)` on the device (full alm, mmax == lmax)."""
    b = a if b is None else b
    if a.lmax != b.lmax or a.mmax != a.lmax or b.mmax != b.lmax:
        raise ValueError("alm2cl_device needs two full alm of the same lmax")
    x = np.ascontiguousarray(a.alm, dtype=np.complex128)
    y = x if b is a else np.ascontiguousarray(b.alm, dtype=np.complex128)
    cl = np.zeros(a.lmax + 1)
    _lib.check(_lib.lib().psb200_alm2cl(a.lmax, x.ctypes.data_as(_lib.DP), y.ctypes.data_as(_lib.DP), _dp(cl)))
    return cl


def release_transform_buffers():
    """Free the per-device tables and work buffers the transforms keep between calls (psb200_sht_release)."""
    _lib.check(_lib.lib().psb200_sht_release())


class CovField:
    """src/workspace.jl:20-60: name, temperature and polarisation masks, pixel variances (i, q, u)."""

    def __init__(self, name, maskT, maskP, sigma2=None):
        self.name = str(name)
        self.maskT = maskT if isinstance(maskT, HealpixMap) else HealpixMap(maskT)
        self.maskP = maskP if isinstance(maskP, HealpixMap) else HealpixMap(maskP)
        if sigma2 is None:                                   # CovField(name, maskT, maskP): zero variance (:49-55)
            z = HealpixMap.zeros(self.maskT.nside)
            sigma2 = PolarizedHealpixMap(z, z, z)
        self.sigma2 = sigma2


def split_maptype(XY: str):
    """"TP" -> ("TT", "PP")  (src/workspace.jl:64-67)."""
    a, b = XY
    return a + a, b + b


def effective_weight_alm(workspace, A, i, j, alpha, *, niter: int = 3) -> Alm:
    """effective_weight_alm! (src/workspace.jl:141-171): map2alm of mask_i^X .* mask_j^Y, times sigma^2_A .* Omega_pix
    for A in (II, QQ, UU) when i == j; zero alm otherwise."""
    from .covariance import NULL
    key = (A, i, j, alpha)
    if key in workspace.effective_weights:
        return workspace.effective_weights[key]
    X, Y = split_maptype(alpha)
    m_iX, m_jY = workspace.mask_p[i, X], workspace.mask_p[j, Y]
    lmax = workspace.lmax
    if A == NULL:
        w = _map2alm_product([m_iX, m_jY], 1.0, lmax, niter)
    elif A in ("II", "QQ", "UU") and i == j:
        omega_p = 4.0 * np.pi / len(m_iX)
        w = _map2alm_product([m_iX, m_jY, workspace.weight_p[i, A]], omega_p, lmax, niter)
    else:
        return Alm(lmax, lmax, np.zeros((lmax + 1) * (lmax + 2) // 2, dtype=np.complex128))      # not cached (:170)
    workspace.effective_weights[key] = w
    return w


def _weight_terms(X):
    return ("II",) if X == "TT" else ("QQ", "UU") if X == "PP" else (X,)


def weights_needed(W_keys):
    """The (A, i, j, alpha) effective weights behind a list of window-spectrum keys (X, Y, i, j, alpha, p, q, beta)."""
    out = []
    for X, Y, i, j, alpha, p, q, beta in W_keys:
        for wx in _weight_terms(X):
            out.append((wx, i, j, alpha))
        for wy in _weight_terms(Y):
            out.append((wy, p, q, beta))
    return list(dict.fromkeys(out))


def precompute_effective_weights(workspace, keys, *, niter: int = 3, ngpus: int = 1):
    """All the effective weights `keys` = [(A, i, j, alpha), ...] not yet in the workspace's cache from ONE library call
    (psb200_map2alm_many): every mask / variance map goes to the device once instead of once per product, and with
    ngpus > 1 the products are dealt to several GPUs.  Same results as effective_weight_alm, which then finds them cached."""
    from .covariance import NULL
    lmax = workspace.lmax
    todo, maps, index = [], [], {}

    def slot(m):
        if id(m) not in index:
            index[id(m)] = len(maps)
            maps.append(m.pixels)
        return index[id(m)]

    for key in dict.fromkeys(keys):
        A, i, j, alpha = key
        if key in workspace.effective_weights:
            continue
        X, Y = split_maptype(alpha)
        m_iX, m_jY = workspace.mask_p[i, X], workspace.mask_p[j, Y]
        if A == NULL:
            todo.append((key, [slot(m_iX), slot(m_jY), -1], 1.0))
        elif A in ("II", "QQ", "UU") and i == j:
            todo.append((key, [slot(m_iX), slot(m_jY), slot(workspace.weight_p[i, A])], 4.0 * np.pi / len(m_iX)))
    if not todo:
        return 0
    nside = HealpixMap(maps[0]).nside
    nalm = (lmax + 1) * (lmax + 2) // 2
    out = [np.zeros(nalm, dtype=np.complex128) for _ in todo]
    idx = (C.c_int * (3 * len(todo)))(*[v for _, ix, _ in todo for v in ix])
    scale = np.array([sc for _, _, sc in todo], dtype=np.float64)
    mp = (_lib.DP * len(maps))(*[_dp(m) for m in maps])
    op = (_lib.DP * len(out))(*[o.ctypes.data_as(_lib.DP) for o in out])
    _lib.check(_lib.lib().psb200_map2alm_many(nside, lmax, int(niter), len(maps), mp, len(todo), idx, _dp(scale), op, int(ngpus)))
    for (key, _, _), o in zip(todo, out):
        workspace.effective_weights[key] = Alm(lmax, lmax, o)
    return len(todo)


def window_spectrum(workspace, X, Y, i, j, alpha, p, q, beta, *, niter: int = 3) -> np.ndarray:
    """The arithmetic of window_function_W! (src/workspace.jl:181-208): TT -> (II,), PP -> (QQ, UU), mean of alm2cl."""
    tx, ty = _weight_terms(X), _weight_terms(Y)
    out = np.zeros(workspace.lmax + 1)
    for wx in tx:
        for wy in ty:
            a = effective_weight_alm(workspace, wx, i, j, alpha, niter=niter)
            b = effective_weight_alm(workspace, wy, p, q, beta, niter=niter)
            out += alm2cl_device(a, b)[:workspace.lmax + 1]
    return out * (1.0 / (len(tx) * len(ty)))
