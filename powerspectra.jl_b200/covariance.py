"""`coupledcov`, `CovarianceWorkspace` and the covariance inner loops, mirroring
/root/reference/src/covariance.jl:34-446 and src/workspace.jl:77-213.

The window spectra W come from spherical-harmonic transforms of mask products
(`effective_weight_alm!` / `window_function_W!`, src/workspace.jl:141-213).  A CovarianceWorkspace
here is either built from four `CovField`s like the reference's (`CovarianceWorkspace.from_fields`:
the transforms then run on the GPU, healpix.py / csrc/psb200_sht.cuh), or is just the *cache* of the
reference's workspace -- field names, lmax and the dictionary of W spectra keyed exactly like
`workspace.W_spectra` (src/workspace.jl:73,83), filled by the caller or lazily by a
`provider(X, Y, i, j, alpha, p, q, beta)` callable.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .spectral import SpectralArray, SpectralVector, spectralones, spectralzeros

NULL = "∅∅"   # the reference's :∅∅ symbol


class ConstantDict(dict):
    """A dictionary that always returns one thing (src/util.jl:51-54)."""

    def __init__(self, c):
        super().__init__()
        self.c = c

    def __getitem__(self, key):
        return self.c

    def __len__(self):
        return 1


class CovarianceWorkspace:
    """field_names = (i, j, p, q); lmax; W_spectra[(X, Y, i, j, alpha, p, q, beta)] -> SpectralVector
    of length lmax+1 (src/workspace.jl:77-84, 197)."""

    def __init__(self, field_names, lmax, W_spectra=None, provider=None):
        if len(field_names) != 4:
            raise ValueError("a covariance relates four fields (i, j, p, q)")
        self.field_names = tuple(field_names)
        self.lmax = int(lmax)
        self.W_spectra = dict(W_spectra or {})
        self.provider = provider
        self.mask_p = None                 # (name, "TT" | "PP") -> HealpixMap           (src/workspace.jl:80)
        self.weight_p = None               # (name, "II" | "QQ" | "UU") -> HealpixMap    (:81)
        self.effective_weights = {}        # (A, i, j, alpha) -> Alm                     (:82)

    @classmethod
    def from_fields(cls, m_i, m_j, m_p, m_q, lmax: int = 0):
        """CovarianceWorkspace(m_i, m_j, m_p, m_q; lmax = 0) of the reference (src/workspace.jl:110-135): four
        CovFields; lmax = 3 nside - 1 when not given (:95, :113)."""
        from .healpix import nside2lmax
        fields = (m_i, m_j, m_p, m_q)
        ws = cls(tuple(f.name for f in fields), lmax if lmax else nside2lmax(m_i.maskT.nside))
        ws.mask_p, ws.weight_p = {}, {}
        for f in fields:
            ws.mask_p[f.name, "TT"] = f.maskT
            ws.mask_p[f.name, "PP"] = f.maskP
            ws.weight_p[f.name, "II"] = f.sigma2.i
            ws.weight_p[f.name, "QQ"] = f.sigma2.q
            ws.weight_p[f.name, "UU"] = f.sigma2.u
        return ws


def window_function_W(workspace, X, Y, i, j, alpha, p, q, beta):
    """Cached lookup of W^{XY, ij alpha, pq beta} (window_function_W!, src/workspace.jl:174-213)."""
    key = (X, Y, i, j, alpha, p, q, beta)
    if key in workspace.W_spectra:
        return workspace.W_spectra[key]
    if workspace.mask_p is not None:
        from .healpix import window_spectrum
        w = SpectralVector(window_spectrum(workspace, X, Y, i, j, alpha, p, q, beta))
        workspace.W_spectra[key] = w
        return w
    if workspace.provider is None:
        raise KeyError(f"window spectrum {key} not in the workspace and no provider given "
                       "(build the workspace with CovarianceWorkspace.from_fields to have it computed)")
    w = workspace.provider(*key)
    if not isinstance(w, SpectralArray):
        w = SpectralVector(w)
    if len(w) != workspace.lmax + 1:
        raise ValueError("window spectra have length workspace.lmax + 1")
    workspace.W_spectra[key] = w
    return w


def _dp(a):
    return a.ctypes.data_as(_lib.DP)


def _loop(block, Cm: SpectralArray, spectra, ratios, Ws, ngpus=1):
    if Cm.ndim != 2 or Cm.axes(0) != Cm.axes(1):
        raise ValueError("covariance matrix must have identical row and column multipoles")
    lmin, lmax = Cm.firstindex(0), Cm.lastindex(0)
    sv = [s.zero_based(lmax) if isinstance(s, SpectralArray) else np.ascontiguousarray(s, dtype=np.float64) for s in spectra]
    rv = [r.zero_based(lmax) if isinstance(r, SpectralArray) else np.ascontiguousarray(r, dtype=np.float64) for r in ratios]
    wv = []
    for w in Ws:
        if isinstance(w, SpectralArray):
            if w.offsets[0] != 0:
                raise ValueError("window spectra must be 0-indexed")
            w = w.parent
        wv.append(np.ascontiguousarray(w, dtype=np.float64))
    lenW = min(w.size for w in wv)
    for a in sv + rv:
        if a.size < lmax + 1:
            raise ValueError("spectra and noise ratios must reach lmax")

    def ptrs(arrs):
        return (_lib.DP * max(len(arrs), 1))(*[_dp(a) for a in arrs])

    rc = _lib.lib().psb200_cov(block, lmin, lmax, ptrs(sv), len(sv), ptrs(rv), len(rv), ptrs(wv), len(wv),
                               lenW, _dp(Cm.parent), Cm.parent.shape[0], ngpus)
    _lib.check(rc)
    return Cm


def loop_covTTTT(Cm, TTip, TTjq, TTiq, TTjp, r_ip, r_jq, r_iq, r_jp, W1, W2, W3, W4, W5, W6, W7, W8, ngpus=1):
    """loop_covTTTT! (src/covariance.jl:92-122)."""
    return _loop(0, Cm, (TTip, TTjq, TTiq, TTjp), (r_ip, r_jq, r_iq, r_jp), (W1, W2, W3, W4, W5, W6, W7, W8), ngpus)


def loop_covEEEE(Cm, EEip, EEjq, EEiq, EEjp, r_ip, r_jq, r_iq, r_jp, W1, W2, W3, W4, W5, W6, W7, W8, ngpus=1):
    """loop_covEEEE! (src/covariance.jl:153-183)."""
    return _loop(1, Cm, (EEip, EEjq, EEiq, EEjp), (r_ip, r_jq, r_iq, r_jp), (W1, W2, W3, W4, W5, W6, W7, W8), ngpus)


def loop_covTTTE(Cm, TTip, TTjp, TEiq, TEjq, r_ip, r_jp, W1, W2, W3, W4, ngpus=1):
    """loop_covTTTE! (src/covariance.jl:208-235)."""
    return _loop(2, Cm, (TTip, TTjp, TEiq, TEjq), (r_ip, r_jp), (W1, W2, W3, W4), ngpus)


def loop_covTETE(Cm, TTip, EEjq, TEiq, TEjp, r_TT_ip, r_PP_jq, W1, W2, W3, W4, W5, ngpus=1):
    """loop_covTETE! (src/covariance.jl:261-302)."""
    return _loop(3, Cm, (TTip, EEjq, TEiq, TEjp), (r_TT_ip, r_PP_jq), (W1, W2, W3, W4, W5), ngpus)


def loop_covTEEE_planck(Cm, EEjq, EEjp, TEip, TEiq, r_EE_jq, r_EE_jp, W1, W2, W3, W4, ngpus=1):
    """loop_covTEEE_planck! (src/covariance.jl:376-402)."""
    return _loop(4, Cm, (EEjq, EEjp, TEip, TEiq), (r_EE_jq, r_EE_jp), (W1, W2, W3, W4), ngpus)


def loop_covTEEE(Cm, EEjq, EEjp, TEip, TEiq, r_EE_jq, r_EE_jp, W1, W2, W3, W4, ngpus=1):
    """loop_covTEEE! (src/covariance.jl:337-372)."""
    return _loop(5, Cm, (EEjq, EEjp, TEip, TEiq), (r_EE_jq, r_EE_jp), (W1, W2, W3, W4), ngpus)


def loop_covTTEE(Cm, TEip, TEiq, TEjq, TEjp, W1, W2, ngpus=1):
    """loop_covTTEE! (src/covariance.jl:422-446)."""
    return _loop(6, Cm, (TEip, TEiq, TEjq, TEjp), (), (W1, W2), ngpus)


# ---- wrappers that pick spectra / ratios / W by key (src/covariance.jl:64-88 etc.) -----------

def coupledcovTTTT(Cm, ws, spectra, noiseratios, ngpus=1):
    i, j, p, q = ws.field_names
    W = lambda *k: window_function_W(ws, *k)
    return loop_covTTTT(
        Cm, spectra["TT", i, p], spectra["TT", j, q], spectra["TT", i, q], spectra["TT", j, p],
        noiseratios["TT", i, p], noiseratios["TT", j, q], noiseratios["TT", i, q], noiseratios["TT", j, p],
        W(NULL, NULL, i, p, "TT", j, q, "TT"), W(NULL, NULL, i, q, "TT", j, p, "TT"),
        W(NULL, "TT", i, p, "TT", j, q, "TT"), W(NULL, "TT", j, q, "TT", i, p, "TT"),
        W(NULL, "TT", i, q, "TT", j, p, "TT"), W(NULL, "TT", j, p, "TT", i, q, "TT"),
        W("TT", "TT", i, p, "TT", j, q, "TT"), W("TT", "TT", i, q, "TT", j, p, "TT"), ngpus=ngpus)


def coupledcovEEEE(Cm, ws, spectra, noiseratios, ngpus=1):
    i, j, p, q = ws.field_names
    W = lambda *k: window_function_W(ws, *k)
    return loop_covEEEE(
        Cm, spectra["EE", i, p], spectra["EE", j, q], spectra["EE", i, q], spectra["EE", j, p],
        noiseratios["EE", i, p], noiseratios["EE", j, q], noiseratios["EE", i, q], noiseratios["EE", j, p],
        W(NULL, NULL, i, p, "PP", j, q, "PP"), W(NULL, NULL, i, q, "PP", j, p, "PP"),
        W(NULL, "PP", i, p, "PP", j, q, "PP"), W(NULL, "PP", j, q, "PP", i, p, "PP"),
        W(NULL, "PP", i, q, "PP", j, p, "PP"), W(NULL, "PP", j, p, "PP", i, q, "PP"),
        W("PP", "PP", i, p, "PP", j, q, "PP"), W("PP", "PP", i, q, "PP", j, p, "PP"), ngpus=ngpus)


def coupledcovTTTE(Cm, ws, spectra, noiseratios, ngpus=1):
    i, j, p, q = ws.field_names
    W = lambda *k: window_function_W(ws, *k)
    return loop_covTTTE(
        Cm, spectra["TT", i, p], spectra["TT", j, p], spectra["TE", i, q], spectra["TE", j, q],
        noiseratios["TT", i, p], noiseratios["TT", j, p],
        W(NULL, NULL, i, p, "TT", j, q, "TP"), W(NULL, NULL, i, q, "TP", j, p, "TT"),
        W(NULL, "TT", j, q, "TP", i, p, "TT"), W(NULL, "TT", i, q, "TP", j, p, "TT"), ngpus=ngpus)


def coupledcovTETE(Cm, ws, spectra, noiseratios, ngpus=1):
    i, j, p, q = ws.field_names
    W = lambda *k: window_function_W(ws, *k)
    return loop_covTETE(
        Cm, spectra["TT", i, p], spectra["EE", j, q], spectra["TE", i, q], spectra["TE", j, p],
        noiseratios["TT", i, p], noiseratios["EE", j, q],
        W(NULL, NULL, i, p, "TT", j, q, "PP"), W(NULL, NULL, i, q, "TP", j, p, "PT"),
        W(NULL, "PP", i, p, "TT", j, q, "PP"), W(NULL, "TT", j, q, "PP", i, p, "TT"),
        W("TT", "PP", i, p, "TT", j, q, "PP"), ngpus=ngpus)


def coupledcovTEEE(Cm, ws, spectra, noiseratios, planck=True, ngpus=1):
    i, j, p, q = ws.field_names
    W = lambda *k: window_function_W(ws, *k)
    loop = loop_covTEEE_planck if planck else loop_covTEEE
    return loop(
        Cm, spectra["EE", j, q], spectra["EE", j, p], spectra["TE", i, p], spectra["TE", i, q],
        noiseratios["EE", j, q], noiseratios["EE", j, p],
        W(NULL, NULL, i, p, "TP", j, q, "PP"), W(NULL, NULL, i, q, "TP", j, p, "PP"),
        W(NULL, "PP", i, p, "TP", j, q, "PP"), W(NULL, "PP", i, q, "TP", j, p, "PP"), ngpus=ngpus)


def coupledcovTTEE(Cm, ws, spectra, noiseratios, ngpus=1):
    i, j, p, q = ws.field_names
    W = lambda *k: window_function_W(ws, *k)
    return loop_covTTEE(
        Cm, spectra["TE", i, p], spectra["TE", i, q], spectra["TE", j, q], spectra["TE", j, p],
        W(NULL, NULL, i, p, "TP", j, q, "TP"), W(NULL, NULL, i, q, "TP", j, p, "TP"), ngpus=ngpus)


def coupledcov(ch1, ch2, workspace, spectra, noiseratios=None, *, lmin=0, lmax=None, ngpus=1):
    """Coupled covariance block (src/covariance.jl:34-61).  Returns a 0-indexed-by-default
    SpectralArray over lmin:lmax; an unsupported channel pair prints a message and returns
    None, as the reference does (:60)."""
    lmax = workspace.lmax if lmax is None else lmax
    r = range(lmin, lmax + 1)
    Cm = spectralzeros(r, r)
    if not noiseratios:                                   # :41-45
        noiseratios = ConstantDict(spectralones(range(0, lmax + 1)))
    table = {("TT", "TT"): coupledcovTTTT, ("EE", "EE"): coupledcovEEEE, ("TE", "TE"): coupledcovTETE,
             ("TT", "TE"): coupledcovTTTE, ("TT", "EE"): coupledcovTTEE, ("TE", "EE"): coupledcovTEEE}
    f = table.get((ch1, ch2))
    if f is None:
        print(f"{ch1},{ch2} not implemented")
        return None
    return f(Cm, workspace, spectra, noiseratios, ngpus=ngpus)
