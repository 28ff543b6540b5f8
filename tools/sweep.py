"""BASELINE configs[4]: MCM scaling sweep lmax 767 -> 12287 (kinds TT and fused M++/M--), device-resident,
one process per GPU (run under torchrun for N > 1).  Prints one JSON line per (lmax, kind)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from powerspectra_jl_b200 import device as dev
from powerspectra_jl_b200 import synthetic as syn

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lmaxes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [767, 1535, 3071, 6143, 12287]
reps = 5
for lmax in lmaxes:
    N = lmax + 1
    V = torch.tensor(syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)], device="cuda")
    X = torch.empty((N, N), dtype=torch.float64, device="cuda")
    X2 = torch.empty_like(X)
    edges = dev.band_edges(0, lmax, world)
    lo, hi = edges[rank], edges[rank + 1]
    for kind, name in ((0, "M00"), (4, "Mpp_Mmm")):
        def step():
            dev.mcm_slab(kind, 0, lmax, V, X, X2 if kind == 4 else None, lo, hi)
            for Xo in ((X, X2) if kind == 4 else (X,)):
                dev.gather_bands(Xo, edges, 0, rank, world)
                if rank == 0:
                    dev.finish(Xo, 0, lmax, True)
        for _ in range(2):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            terms = dev.terms(name, lmax, 0, N)
            print(json.dumps({"lmax": lmax, "kind": name, "n_gpus": world, "ms": round(float(ms), 4),
                              "terms_per_s": terms / (float(ms) * 1e-3), "ref_terms": terms}), flush=True)
    del X, X2
if world > 1:
    dist.destroy_process_group()
