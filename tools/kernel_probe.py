"""Times the pair kernels of the headline step (device-resident, one GPU) for the library PSB200_LIB points at.

    PSB200_LIB=tools/_build/libpsb200_<variant>.so python tools/kernel_probe.py <tag> [lmax] >> gpurun_out/probe.jsonl

One JSON line: per-job mean ms over 3 launches after one warm-up, and a checksum of each output (variants of one
source must agree to rounding).  Development tool for A/B-ing tiling parameters; not part of the product."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from powerspectra_jl_b200 import device as dev

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 6143
N = lmax + 1
inp = bench.make_inputs(lmax)
t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
X = [torch.zeros((N, N), dtype=torch.float64, device="cuda") for _ in range(2)]
out = {"tag": tag, "lmax": lmax, "lib": os.environ.get("PSB200_LIB", "default"), "ms": {}, "sum": {}}
total = 0.0
for name, api, code, fam, _ in bench.JOBS + [("M02", "mcm", 1, 2, None), ("Mmm", "mcm", 3, 1, None)]:
    a = inp[name if name in inp else "Mpp_Mmm"]
    if api == "mcm":
        V = t(a["V"])
        fn = lambda: dev.mcm_slab(code, 0, lmax, V, X[0], X[1] if code == 4 else None)
    else:
        sp, rt, W = [t(x) for x in a["sp"]], [t(x) for x in a["rt"]], [t(x) for x in a["W"]]
        fn = lambda: dev.cov_slab(code, 0, lmax, sp, rt, W, X[0])
    fn()
    torch.cuda.synchronize()
    if os.environ.get("PROBE_ONCE"):          # one launch per job: what an ncu capture wants
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out["ms"][name] = round(ms, 3)
    out["sum"][name] = float(torch.triu(X[0]).sum().item())
    if name in {j[0] for j in bench.JOBS}:
        total += ms
out["ms"]["step"] = round(total, 3)
print(json.dumps(out))
