"""`mcm` and the inner mode-coupling loops, mirroring /root/reference/src/modecoupling.jl:78-256.

Only the callees (`inner_mcm00!` ...) run on the GPU; the dispatcher keeps the
reference's behaviour: same spec names, `lmin` / `lmax` keywords, ValueError (Julia:
ArgumentError, src/modecoupling.jl:225) on an unknown spec, 2x2 block assembly for
`EE_BB` / `EB_BE` (src/modecoupling.jl:210-223, 235-244).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .spectral import BlockSpectralMatrix, SpectralArray, SpectralVector, spectralzeros

# spec -> kind of include/psb200.h.  Julia symbols with superscripts are accepted as strings.
_M00 = ("TT", "M00", "M⁰⁰")
_M02 = ("TE", "ET", "TB", "BT", "M02", "M20", "M⁰²", "M²⁰")
_MPP = ("M++", "Mpp", "M⁺⁺")
_MMM = ("M--", "Mmm", "M⁻⁻")


class Alm:
    """Minimal stand-in for Healpix.Alm: complex a_lm, m-major (healpy) ordering,
    index(l, m) = m (2 lmax + 1 - m) / 2 + l, 0 <= m <= mmax <= lmax."""

    def __init__(self, lmax, mmax, alm):
        self.lmax, self.mmax = int(lmax), int(mmax)
        self.alm = np.asarray(alm, dtype=np.complex128)
        n = (self.mmax + 1) * (self.lmax + 1) - self.mmax * (self.mmax + 1) // 2
        if self.alm.size != n:
            raise ValueError(f"expected {n} coefficients for lmax={lmax}, mmax={mmax}")

    @classmethod
    def zonal(cls, al0):
        """Alm of an azimuthally symmetric field: only a_l0 non-zero."""
        al0 = np.asarray(al0, dtype=np.float64)
        return cls(al0.size - 1, 0, al0.astype(np.complex128))


def alm2cl(a: Alm, b: Alm) -> np.ndarray:
    """Cross-spectrum C_l = [a_l0 b_l0 + 2 sum_{m>0} Re(a_lm conj b_lm)] / (2l+1) (Healpix.alm2cl)."""
    lmax = min(a.lmax, b.lmax)
    mmax = min(a.mmax, b.mmax)
    cl = np.zeros(lmax + 1)
    for m in range(mmax + 1):
        ia = m * (2 * a.lmax + 1 - m) // 2
        ib = m * (2 * b.lmax + 1 - m) // 2
        ls = np.arange(m, lmax + 1)
        prod = (a.alm[ia + ls] * np.conj(b.alm[ib + ls])).real
        cl[ls] += prod if m == 0 else 2.0 * prod
    return cl / (2.0 * np.arange(lmax + 1) + 1.0)


def _dp(a):
    return a.ctypes.data_as(_lib.DP)


def _inner(kind, M: SpectralArray, V: SpectralArray, M2: SpectralArray | None = None, ngpus=1):
    if M.ndim != 2 or M.axes(0) != M.axes(1):
        raise ValueError("mode-coupling matrix must have identical row and column multipoles")  # @assert :80
    if V.ndim != 1 or V.offsets[0] != 0:
        raise ValueError("mask spectrum must be a 0-indexed SpectralVector")
    lmin, lmax = M.firstindex(0), M.lastindex(0)
    v = np.ascontiguousarray(V.parent, dtype=np.float64)
    out2 = _dp(M2.parent) if M2 is not None else None
    rc = _lib.lib().psb200_mcm(kind, lmin, lmax, _dp(v), v.size, _dp(M.parent), M.parent.shape[0], out2, ngpus)
    _lib.check(rc)
    return M


def inner_mcm00(M, V, ngpus=1):
    """inner_mcm⁰⁰! (src/modecoupling.jl:78-95)."""
    return _inner(0, M, V, ngpus=ngpus)


def inner_mcm02(M, V, ngpus=1):
    """inner_mcm⁰²! (src/modecoupling.jl:99-119)."""
    return _inner(1, M, V, ngpus=ngpus)


def inner_mcmpp(M, V, ngpus=1):
    """inner_mcm⁺⁺! (src/modecoupling.jl:123-139)."""
    return _inner(2, M, V, ngpus=ngpus)


def inner_mcmmm(M, V, ngpus=1):
    """inner_mcm⁻⁻! (src/modecoupling.jl:143-159)."""
    return _inner(3, M, V, ngpus=ngpus)


def inner_mcmpp_mcmmm(Mpp, Mmm, V, ngpus=1):
    """Both spin-2 blocks from one evaluation of the (0,-2,2) family (the reference calls
    inner_mcm⁺⁺! and inner_mcm⁻⁻! back to back, src/modecoupling.jl:213-214)."""
    _inner(4, Mpp, V, M2=Mmm, ngpus=ngpus)
    return Mpp, Mmm


def mcm_master(maskT1, maskP1, maskT2, maskP2, *, lmin=0, lmax=None, ngpus=1):
    """All mode-coupling matrices `master` needs (src/modecoupling.jl:339-377) in one fused GPU pass:
    returns dict TT, TE (= TB), ET (= BT) -> SpectralArray and EE_BB, EB_BE -> BlockSpectralMatrix.
    Arguments are the mask Alm's of the two maps (T and P mask each)."""
    if lmax is None:
        lmax = min(a.lmax for a in (maskT1, maskP1, maskT2, maskP2))
    V = [np.ascontiguousarray(alm2cl(a, b)[: lmax + 1]) for a, b in
         ((maskT1, maskT2), (maskT1, maskP2), (maskP1, maskT2), (maskP1, maskP2))]
    r = range(lmin, lmax + 1)
    out = [spectralzeros(r, r) for _ in range(5)]
    N = lmax - lmin + 1
    rc = _lib.lib().psb200_mcm_master(lmin, lmax, _dp(V[0]), _dp(V[1]), _dp(V[2]), _dp(V[3]), lmax + 1,
                                      *[_dp(o.parent) for o in out], N, ngpus)
    _lib.check(rc)
    M00, M02tp, M02pt, Mpp, Mmm = out
    neg = SpectralArray(-Mmm.parent, Mmm.offsets)
    return {"TT": M00, "TE": M02tp, "TB": M02tp, "ET": M02pt, "BT": M02pt,
            "EE_BB": BlockSpectralMatrix([[Mpp, Mmm], [Mmm, Mpp]]),
            "EB_BE": BlockSpectralMatrix([[Mpp, neg], [neg, Mpp]])}


def maskedalm2spectra(maskedmap1, maskT1, maskP1, maskedmap2, maskT2, maskP2, *, lmin=0, ngpus=1):
    """Decoupled TT, TE, ET, TB, BT, EE, BB, EB, BE spectra from the alms of the masked maps
    (T, E, B triples) and of the masks (src/modecoupling.jl:339-377).  The mode-coupling
    matrices come from the fused GPU pass; alm2cl and the LU solves stay on the host."""
    M = mcm_master(maskT1, maskP1, maskT2, maskP2, lmin=lmin, ngpus=ngpus)
    a1 = dict(zip("TEB", maskedmap1))
    a2 = dict(zip("TEB", maskedmap2))
    lmax = M["TT"].lastindex(0)

    def pcl(x, y):
        return SpectralVector(alm2cl(a1[x], a2[y])[lmin: lmax + 1], lmin)
    spectra = {}
    for x, y in (("T", "T"), ("T", "E"), ("E", "T"), ("T", "B"), ("B", "T")):
        spectra[x + y] = M[x + y].solve(pcl(x, y))
    spectra["EE"], spectra["BB"] = M["EE_BB"].solve([pcl("E", "E"), pcl("B", "B")])
    spectra["EB"], spectra["BE"] = M["EB_BE"].solve([pcl("E", "B"), pcl("B", "E")])
    return spectra


def maskedalm2spectra_device(maskedmap1, maskT1, maskP1, maskedmap2, maskT2, maskP2, *, lmin=0, ngpus=1):
    """maskedalm2spectra (src/modecoupling.jl:339-377) with the decoupling on the device: one fused pass builds the five
    mode-coupling matrices on `ngpus` GPUs, they are assembled and LU-solved on one of them (psb200_master_solve) and
    only the nine decoupled spectra come back -- no matrix crosses PCIe.  alm2cl of the maps stays on the host."""
    lmax = min(a.lmax for a in (maskT1, maskP1, maskT2, maskP2))
    V = [np.ascontiguousarray(alm2cl(a, b)[: lmax + 1]) for a, b in
         ((maskT1, maskT2), (maskT1, maskP2), (maskP1, maskT2), (maskP1, maskP2))]
    a1 = dict(zip("TEB", maskedmap1))
    a2 = dict(zip("TEB", maskedmap2))
    names = ("TT", "TE", "ET", "TB", "BT", "EE", "BB", "EB", "BE")
    N = lmax - lmin + 1
    pcl = np.asfortranarray(np.stack([alm2cl(a1[x], a2[y])[lmin: lmax + 1] for x, y in names], axis=1))
    cl = np.zeros_like(pcl, order="F")
    rc = _lib.lib().psb200_master_solve(lmin, lmax, _dp(V[0]), _dp(V[1]), _dp(V[2]), _dp(V[3]), lmax + 1,
                                        _dp(pcl), N, _dp(cl), N, ngpus)
    _lib.check(rc)
    return {n: SpectralVector(np.ascontiguousarray(cl[:, k]), lmin) for k, n in enumerate(names)}


_SYSTEMS = {"TT": 0, "M00": 0, "TE": 1, "ET": 1, "TB": 1, "BT": 1, "M02": 1, "M20": 1, "M++": 2, "M--": 3,
            "EE_BB": 4, "EB_BE": 5}


def mcm_solve(spec, alm1, alm2, pcl, *, lmin=0, lmax=None, ngpus=1):
    """`mcm(spec, alm1, alm2; lmin) \\ pCl` (src/modecoupling.jl:359-379, src/blockspectralmatrix.jl:89-129) without
    bringing the matrix to the host (psb200_mcm_solve).  pcl: a SpectralVector / array over lmin:lmax, a list of two for
    the block systems "EE_BB" / "EB_BE" (stacked like `[pCl_EE; pCl_BB]`), or a 2-d array of several right-hand sides
    as columns.  Returns SpectralVector(s) over lmin:lmax (a 2-d SpectralArray for several right-hand sides)."""
    if spec not in _SYSTEMS:
        raise ValueError(f"{spec} not a valid spectrum.")
    system = _SYSTEMS[spec]
    V, lmax = _mask_spectrum(alm1, alm2, lmax)
    N = lmax - lmin + 1
    nb = 2 if system >= 4 else 1
    get = lambda v: v.parent if isinstance(v, SpectralArray) else np.asarray(v, dtype=np.float64)
    if nb == 2 and isinstance(pcl, (list, tuple)):
        rhs = np.concatenate([get(v) for v in pcl])
    else:
        rhs = get(pcl)
    many = rhs.ndim == 2
    rhs = np.asfortranarray(rhs.reshape(rhs.shape[0], -1))
    if rhs.shape[0] != nb * N:
        raise ValueError(f"right-hand side has {rhs.shape[0]} rows, the system has {nb * N}")
    out = np.zeros_like(rhs, order="F")
    rc = _lib.lib().psb200_mcm_solve(system, lmin, lmax, _dp(V.parent), V.parent.size, _dp(rhs), nb * N, rhs.shape[1],
                                     _dp(out), nb * N, ngpus)
    _lib.check(rc)
    if many:
        return SpectralArray(out, (lmin, 0))
    if nb == 2:
        return SpectralVector(out[:N, 0].copy(), lmin), SpectralVector(out[N:, 0].copy(), lmin)
    return SpectralVector(out[:, 0].copy(), lmin)


def _mask_spectrum(alm1, alm2, lmax):
    if isinstance(alm1, SpectralArray) and alm2 is None:      # already a cross-spectrum V
        if lmax is None:
            lmax = alm1.lastindex(0)
        return SpectralVector(alm1.zero_based(lmax)), lmax
    if lmax is None:
        lmax = min(alm1.lmax, alm2.lmax)                       # src/modecoupling.jl:194-196
    return SpectralVector(alm2cl(alm1, alm2)[: lmax + 1]), lmax  # :197


def mcm(spec, alm1, alm2=None, *, lmin=0, lmax=None, ngpus=1):
    """Mode-coupling matrix (src/modecoupling.jl:192-246).

    spec: "TT" | "TE"/"ET"/"TB"/"BT" | "M++" | "M--" | "EE_BB" | "EB_BE" | ("EE_BB", "EB_BE").
    alm1, alm2: mask Alm's (or one 0-indexed SpectralVector holding their cross-spectrum).
    Returns a SpectralArray over lmin:lmax, or BlockSpectralMatrix(es) for the 2x2 specs.
    """
    V, lmax = _mask_spectrum(alm1, alm2, lmax)
    r = range(lmin, lmax + 1)
    if isinstance(spec, tuple):
        if spec == ("EE_BB", "EB_BE"):
            Mpp, Mmm = inner_mcmpp_mcmmm(spectralzeros(r, r), spectralzeros(r, r), V, ngpus)
            neg = SpectralArray(-Mmm.parent, Mmm.offsets)
            return (BlockSpectralMatrix([[Mpp, Mmm], [Mmm, Mpp]]),
                    BlockSpectralMatrix([[Mpp, neg], [neg, Mpp]]))
        raise ValueError(f"{spec} not a valid spectrum.")
    if spec in _M00:
        return inner_mcm00(spectralzeros(r, r), V, ngpus)
    if spec in _M02:
        return inner_mcm02(spectralzeros(r, r), V, ngpus)
    if spec in _MPP:
        return inner_mcmpp(spectralzeros(r, r), V, ngpus)
    if spec in _MMM:
        return inner_mcmmm(spectralzeros(r, r), V, ngpus)
    if spec in ("EE_BB", "EB_BE"):
        Mpp, Mmm = inner_mcmpp_mcmmm(spectralzeros(r, r), spectralzeros(r, r), V, ngpus)
        if spec == "EE_BB":
            return BlockSpectralMatrix([[Mpp, Mmm], [Mmm, Mpp]])
        neg = SpectralArray(-Mmm.parent, Mmm.offsets)
        return BlockSpectralMatrix([[Mpp, neg], [neg, Mpp]])
    raise ValueError(f"{spec} not a valid spectrum.")
