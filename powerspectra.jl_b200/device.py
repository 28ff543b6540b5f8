"""Device-resident driver of the hot path: buffers live in HBM (torch tensors are used only
as device memory + streams + torch.distributed plumbing), the kernels are libpsb200's.

One process per GPU.  Each rank computes the Xi slab of its work-balanced l1 band
(psb200_mcm_dev / psb200_cov_dev), slabs are gathered to rank 0 with NCCL send/recv over
NVLink (a band of rows is one contiguous slab of the column-major result, see
include/psb200.h), and rank 0 runs the finish kernel that fills both triangles.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

MCM_KINDS = {"M00": 0, "M02": 1, "Mpp": 2, "Mmm": 3, "Mpp_Mmm": 4}
COV_BLOCKS = {"TTTT": 0, "EEEE": 1, "TTTE": 2, "TETE": 3, "TEEE_planck": 4, "TEEE": 5, "TTEE": 6}
REF_FAMILIES = {"master": 7, "M00": 1, "M02": 2, "Mpp": 1, "Mmm": 1, "Mpp_Mmm": 2,
                "TTTT": 1, "EEEE": 1, "TTTE": 1, "TETE": 2, "TEEE_planck": 1, "TEEE": 2, "TTEE": 1}


def band_edges(lmin: int, lmax: int, nbands: int, lenW: int | None = None):
    """Work-balanced l1 bands; lenW = window length of the call (default lmax+1, what `mcm` passes);
    lenW = 0 balances the reference's full-family term count instead."""
    e = (C.c_int * (nbands + 1))()
    _lib.check(_lib.lib().psb200_band_edges(lmin, lmax, lmax + 1 if lenW is None else lenW, nbands, e))
    return list(e)


def terms(name: str, lmax: int, row_lo: int, row_hi: int) -> int:
    """3j terms, counted as the reference evaluates them (full families)."""
    return int(_lib.lib().psb200_terms(REF_FAMILIES[name], lmax, row_lo, row_hi))


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
        raise ValueError("device API needs contiguous float64 CUDA tensors")


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _vptrs(ts):
    arr = (C.c_void_p * max(len(ts), 1))()
    for k, t in enumerate(ts):
        _require_cuda(t)
        arr[k] = t.data_ptr()
    return arr


def mcm_slab(kind, lmin, lmax, V: torch.Tensor, X: torch.Tensor, X2: torch.Tensor | None = None,
             row_lo=None, row_hi=None, bands=None):
    """Stage 1 for one band: X is the FULL (N, N) buffer (X[l1-lmin, l2-lmin], torch row-major =
    the column-major result's transpose) or any view whose data_ptr is row `lmin` of it.
    bands = [(lo, hi), ...] (up to 4, disjoint): several bands in ONE launch (psb200_mcm_dev_bands; folded split)."""
    kind = MCM_KINDS.get(kind, kind)
    N = lmax - lmin + 1
    row_lo = lmin if row_lo is None else row_lo
    row_hi = lmax + 1 if row_hi is None else row_hi
    _require_cuda(V)
    _require_cuda(X)
    x2 = None
    if X2 is not None:
        _require_cuda(X2)
        x2 = C.c_void_p(X2.data_ptr())
    if bands is not None:
        rc = _lib.lib().psb200_mcm_dev_bands(kind, lmin, lmax, C.c_void_p(V.data_ptr()), V.numel(),
                                             C.c_void_p(X.data_ptr()), N, x2, _bands(bands), len(bands), _stream_ptr())
    else:
        rc = _lib.lib().psb200_mcm_dev(kind, lmin, lmax, C.c_void_p(V.data_ptr()), V.numel(),
                                       C.c_void_p(X.data_ptr()), N, x2, row_lo, row_hi, _stream_ptr())
    _lib.check(rc)


def master_slab(lmin, lmax, V_TT, V_TP, V_PT, V_PP, Xs, row_lo=None, row_hi=None):
    """Fused stage 1 of psb200_mcm_master_dev: Xs = five (N, N) buffers M00, M02_TP, M02_PT, Mpp, Mmm."""
    N = lmax - lmin + 1
    row_lo = lmin if row_lo is None else row_lo
    row_hi = lmax + 1 if row_hi is None else row_hi
    for t in (V_TT, V_TP, V_PT, V_PP, *Xs):
        _require_cuda(t)
    rc = _lib.lib().psb200_mcm_master_dev(lmin, lmax, C.c_void_p(V_TT.data_ptr()), C.c_void_p(V_TP.data_ptr()),
                                          C.c_void_p(V_PT.data_ptr()), C.c_void_p(V_PP.data_ptr()), V_TT.numel(),
                                          _vptrs(Xs), N, row_lo, row_hi, _stream_ptr())
    _lib.check(rc)


def cov_slab(block, lmin, lmax, spectra, ratios, W, X: torch.Tensor, row_lo=None, row_hi=None, bands=None):
    block = COV_BLOCKS.get(block, block)
    N = lmax - lmin + 1
    row_lo = lmin if row_lo is None else row_lo
    row_hi = lmax + 1 if row_hi is None else row_hi
    _require_cuda(X)
    lenW = min(int(w.numel()) for w in W)
    for t in list(spectra) + list(ratios):      # the C entry point takes no lengths for these: l = 0..lmax is read
        if int(t.numel()) < lmax + 1:
            raise ValueError(f"spectrum / ratio vector of {int(t.numel())} entries, need lmax+1 = {lmax + 1}")
    if bands is not None:
        rc = _lib.lib().psb200_cov_dev_bands(block, lmin, lmax, _vptrs(spectra), len(spectra), _vptrs(ratios), len(ratios),
                                             _vptrs(W), len(W), lenW, C.c_void_p(X.data_ptr()), N, _bands(bands), len(bands),
                                             _stream_ptr())
    else:
        rc = _lib.lib().psb200_cov_dev(block, lmin, lmax, _vptrs(spectra), len(spectra), _vptrs(ratios), len(ratios),
                                       _vptrs(W), len(W), lenW, C.c_void_p(X.data_ptr()), N, row_lo, row_hi,
                                       _stream_ptr())
    _lib.check(rc)


def finish(X: torch.Tensor, lmin, lmax, scale: bool):
    """Stage 2 in place: both triangles, (2l+1) factors when scale (MCM), plain mirror otherwise."""
    _require_cuda(X)
    N = lmax - lmin + 1
    _lib.check(_lib.lib().psb200_finish_dev(C.c_void_p(X.data_ptr()), N, lmin, lmax, 1 if scale else 0,
                                            _stream_ptr()))


def _bands(bands):
    flat = [int(v) for b in bands for v in b]
    return (C.c_int * len(flat))(*flat)


def folded_bands(lmin: int, lmax: int, world: int, lenW: int | None = None):
    """Folded split for `world` ranks: the rows are cut into 2*world pieces of equal kernel cost and rank r owns pieces
    r and 2*world-1-r -- a low band (many short tiles) and a high band (few long ones) -- launched as ONE tile list, so
    every rank packs its last waves with short tiles and errors of the cost model average out.  Returns
    [[(lo, hi), (lo, hi)] for each rank] (one band per rank when world == 1)."""
    if world == 1:
        return [[(lmin, lmax + 1)]]
    e = band_edges(lmin, lmax, 2 * world, lenW)
    return [[(e[r], e[r + 1]), (e[2 * world - 1 - r], e[2 * world - r])] for r in range(world)]


def gather_slabs(X: torch.Tensor, owners, lmin, rank, group=None):
    """Send every band slab X[lo-lmin : hi-lmin, :] of `owners` = [[(lo, hi), ...] per rank] to rank 0 (grouped
    send/recv).  Works for NCCL (cuda) and gloo (cpu)."""
    import torch.distributed as dist
    if len(owners) == 1:
        return
    ops = []
    for r, bands in enumerate(owners):
        if r == 0:
            continue
        for lo, hi in bands:
            if hi <= lo:
                continue
            if rank == 0:
                ops.append(dist.P2POp(dist.irecv, X[lo - lmin:hi - lmin], r, group))
            elif rank == r:
                ops.append(dist.P2POp(dist.isend, X[lo - lmin:hi - lmin], 0, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def gather_bands(X: torch.Tensor, edges, lmin, rank, world, group=None):
    """Send each rank's contiguous band slab X[e_r-lmin : e_{r+1}-lmin, :] to rank 0 (variable
    sizes => grouped send/recv, there is no gatherv).  Works for NCCL (cuda) and gloo (cpu)."""
    import torch.distributed as dist
    if world == 1:
        return
    ops = []
    if rank == 0:
        for r in range(1, world):
            a, b = edges[r] - lmin, edges[r + 1] - lmin
            if b > a:
                ops.append(dist.P2POp(dist.irecv, X[a:b], r, group))
    else:
        a, b = edges[rank] - lmin, edges[rank + 1] - lmin
        if b > a:
            ops.append(dist.P2POp(dist.isend, X[a:b], 0, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def quickpol_edges(lmax: int, band_lo: int, band_hi: int, nbands: int):
    """Cost-balanced column bands of the QuickPol Xi matrix (psb200_quickpol_edges)."""
    e = (C.c_int * (nbands + 1))()
    _lib.check(_lib.lib().psb200_quickpol_edges(lmax, band_lo, band_hi, nbands, e))
    return list(e)


def quickpol_slab(nu1, nu2, s1, s2, lmax, W: torch.Tensor, Xb: torch.Tensor, band_lo, band_hi,
                  col_lo=None, col_hi=None):
    """Columns [col_lo, col_hi) of the Xi band storage on this rank's GPU (psb200_quickpol_xi_dev).
    Xb is the FULL (lmax+1, ldb) buffer, row l = column l of the column-major band storage
    (Xb[l, band_hi + l'' - l]); a band of columns is one contiguous slab, so `gather_bands(Xb, edges, 0, ...)`
    collects the ranks' slabs on rank 0 exactly as it does for the mode-coupling matrices."""
    _require_cuda(W)
    _require_cuda(Xb)
    if Xb.shape[0] != lmax + 1 or Xb.shape[1] < band_lo + band_hi + 1:
        raise ValueError("Xb must be (lmax+1, >= band_lo+band_hi+1)")
    col_lo = 0 if col_lo is None else col_lo
    col_hi = lmax + 1 if col_hi is None else col_hi
    rc = _lib.lib().psb200_quickpol_xi_dev(nu1, nu2, s1, s2, lmax, C.c_void_p(W.data_ptr()), W.numel(), band_lo,
                                           band_hi, C.c_void_p(Xb.data_ptr()), Xb.shape[1], col_lo, col_hi,
                                           _stream_ptr())
    _lib.check(rc)


def dfma_peak(iters: int = 4096) -> float:
    """Measured FP64 DFMA throughput (FLOP/s) of the current device."""
    v = float(_lib.lib().psb200_dfma_peak(iters))
    if v < 0:
        _lib.check(2)
    return v


# ---- W-spectrum production on device buffers (psb200_sht.cuh) -------------------------------------
def alm_size(lmax: int) -> int:
    return (lmax + 1) * (lmax + 2) // 2


def map2alm_dev(nside: int, lmax: int, dmap: torch.Tensor, dalm: torch.Tensor, niter: int = 3):
    """dalm (complex128, Healpix.jl Alm order) = map2alm(dmap; lmax, niter) on the current stream; dmap is not modified."""
    _require_cuda(dmap)
    if dalm.dtype != torch.complex128 or not dalm.is_cuda or not dalm.is_contiguous() or dalm.numel() != alm_size(lmax):
        raise ValueError("dalm must be a contiguous complex128 CUDA tensor of (lmax+1)(lmax+2)/2 entries")
    if dmap.numel() != 12 * nside * nside:
        raise ValueError("dmap must hold 12 nside^2 pixels")
    _lib.check(_lib.lib().psb200_map2alm_dev(nside, lmax, niter, C.c_void_p(dmap.data_ptr()), C.c_void_p(dalm.data_ptr()),
                                             _stream_ptr()))


def alm2map_dev(nside: int, lmax: int, dalm: torch.Tensor, dmap: torch.Tensor):
    _require_cuda(dmap)
    if dalm.dtype != torch.complex128 or not dalm.is_cuda or not dalm.is_contiguous() or dalm.numel() != alm_size(lmax):
        raise ValueError("dalm must be a contiguous complex128 CUDA tensor of (lmax+1)(lmax+2)/2 entries")
    if dmap.numel() != 12 * nside * nside:
        raise ValueError("dmap must hold 12 nside^2 pixels")
    _lib.check(_lib.lib().psb200_alm2map_dev(nside, lmax, C.c_void_p(dalm.data_ptr()), C.c_void_p(dmap.data_ptr()), _stream_ptr()))


def alm2cl_dev(lmax: int, a: torch.Tensor, b: torch.Tensor, cl: torch.Tensor):
    _require_cuda(cl)
    if cl.numel() < lmax + 1 or a.numel() != alm_size(lmax) or b.numel() != alm_size(lmax):
        raise ValueError("alm2cl_dev: sizes do not match lmax")
    _lib.check(_lib.lib().psb200_alm2cl_dev(lmax, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(cl.data_ptr()),
                                            _stream_ptr()))


def sht_stats(nside: int, lmax: int):
    st = (C.c_longlong * 6)()
    _lib.check(_lib.lib().psb200_sht_stats(nside, lmax, st))
    return {"exec_steps": int(st[0]), "live_steps": int(st[1]), "warps": int(st[2]), "R": int(st[3]), "chunks": int(st[4]),
            "C": int(st[5])}
