"""world_size-2 test of the multi-GPU host logic on CPU (gloo): work-balanced l1 bands, the
variable-size slab gather to rank 0 (device.gather_bands, the same code the NCCL path runs) and
the band bookkeeping.  The per-band compute is stood in for by the CPU oracle -- allowed here
because this is a test; the product computes bands with psb200_mcm_dev on each rank's GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lmin, lmax, out_path, folded=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import psoracle as po
    from powerspectra_jl_b200 import device as dev
    from powerspectra_jl_b200 import synthetic as syn

    V = syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)]
    N = lmax - lmin + 1
    edges = dev.band_edges(lmin, lmax, world)
    owners = dev.folded_bands(lmin, lmax, world) if folded else [[(edges[r], edges[r + 1])] for r in range(world)]
    # X[l1-lmin, l2-lmin] row-major == the column-major result transposed; stage 1 fills rows of my band(s)
    X = torch.full((N, N), float("nan"), dtype=torch.float64)
    full = po.mcm(0, lmin, lmax, V)                       # oracle: M[l1,l2] = (2 l2+1) Xi
    xi = full / (2.0 * np.arange(lmin, lmax + 1) + 1.0)[None, :]
    for lo, hi in owners[rank]:
        for l1 in range(lo, hi):
            X[l1 - lmin, l1 - lmin:] = torch.from_numpy(xi[l1 - lmin, l1 - lmin:])
    if folded:
        dev.gather_slabs(X, owners, lmin, rank)
        edges = [lmin, owners[0][0][1]]                   # rank 0's low band ends here
    else:
        dev.gather_bands(X, edges, lmin, rank, world)
    if rank == 0:
        # stage 2 (what psb200_finish_dev does on the GPU), in numpy
        Xn = X.numpy()
        iu = np.triu_indices(N)
        assert not np.isnan(Xn[iu]).any(), "a band did not arrive"
        M = np.zeros((N, N))
        sc = 2.0 * np.arange(lmin, lmax + 1) + 1.0
        M[iu] = Xn[iu] * sc[iu[1]]
        M.T[iu] = Xn[iu] * sc[iu[0]]
        np.save(out_path, np.array([np.max(np.abs(M - full)), float(edges[1])]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lmin,lmax", [(0, 95), (2, 130)])
def test_band_gather_two_ranks(tmp_path, lmin, lmax):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), lmin, lmax, out), nprocs=2, join=True)
    err, edge = np.load(out)
    assert err < 1e-15          # (x / s) * s rounding only: every band arrived in the right place
    assert lmin < edge <= lmax          # both ranks own rows


@pytest.mark.parametrize("lmin,lmax", [(0, 95), (3, 140)])
def test_folded_band_gather_two_ranks(tmp_path, lmin, lmax):
    """The folded split of the N-GPU driver: each rank owns a low and a high band (device.folded_bands), both slabs of
    every rank reach rank 0 (device.gather_slabs)."""
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), lmin, lmax, out, True), nprocs=2, join=True)
    err, edge = np.load(out)
    assert err < 1e-15
    assert lmin < edge < lmax


def _qp_worker(rank, world, port, lmax, bl, bh, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import psoracle as po
    from powerspectra_jl_b200 import device as dev

    rng = np.random.default_rng(3)
    W = rng.normal(size=2 * lmax + 1)
    full = po.quickpol_xi(2, 0, 1, -1, lmax, W, bl, bh, dense=False)       # (nb, lmax+1) band storage
    edges = dev.quickpol_edges(lmax, bl, bh, world)
    lo, hi = edges[rank], edges[rank + 1]
    X = torch.full((lmax + 1, bl + bh + 1), float("nan"), dtype=torch.float64)   # row l = column l of the band storage
    X[lo:hi] = torch.from_numpy(np.ascontiguousarray(full.T[lo:hi]))          # stand-in for quickpol_slab on my GPU
    dev.gather_bands(X, edges, 0, rank, world)
    if rank == 0:
        Xn = X.numpy()
        assert not np.isnan(Xn).any(), "a column band did not arrive"
        np.save(out_path, np.array([float(np.max(np.abs(Xn.T - full))), float(edges[1])]))
    dist.barrier()
    dist.destroy_process_group()


def test_quickpol_column_gather_two_ranks(tmp_path):
    out = str(tmp_path / "res.npy")
    lmax = 90
    mp.spawn(_qp_worker, args=(2, _free_port(), lmax, 12, 7, out), nprocs=2, join=True)
    err, edge = np.load(out)
    assert err == 0.0
    assert 0.5 * lmax < edge < 0.8 * lmax        # column cost grows with l (plus a fixed per-pair part): split above the middle
