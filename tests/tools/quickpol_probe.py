"""Times the QuickPol Xi kernel (psb200_quickpol_xi[_dev]) on one GPU: both instantiations, device-resident
and end to end through the host-level C call, next to the CPU oracle on a sample of rows.

  python tests/tools/quickpol_probe.py [lmax] [band] [out.json]

A "term" is one 3j family value as the reference evaluates it (src/beam.jl:86-93: two full families per
stored pair).  Writes one JSON line (stdout and, if given, out.json)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import powerspectra_jl_b200 as ps                      # noqa: E402
from oracle import psoracle as po                      # noqa: E402  (CPU baseline leg only)

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6143
band = int(sys.argv[2]) if len(sys.argv) > 2 else 128
out = sys.argv[3] if len(sys.argv) > 3 else None
case = (2, -2, 2, 2)                                    # nu1, nu2, s1, s2: m1 = 4 and 0
nu1, nu2, s1, s2 = case
nb = 2 * band + 1
rng = np.random.default_rng(7)
l = np.arange(2 * lmax + 1)
W = rng.normal(size=l.size) / (1.0 + l / 40.0) ** 2 + 1.0 / (1.0 + l) ** 1.5

# reference term count: both families, full length, every stored pair
terms = 0
for lpp in range(2, lmax + 1):
    ls = np.arange(max(2, lpp - band), min(lmax, lpp + band) + 1)
    d = np.abs(ls - lpp)
    for m1 in (s1 + nu1, s2 + nu2):
        lo = np.maximum(d, abs(m1))
        terms += int(np.sum(np.maximum(ls + lpp - lo + 1, 0)))

L = ps.lib()
dW = torch.tensor(W, device="cuda")
dX = torch.zeros((lmax + 1, nb), device="cuda", dtype=torch.float64)
res = {"what": "quickpol Xi", "lmax": lmax, "band": band, "case": case, "terms": terms, "pairs": int((lmax - 1) * nb)}
ref = None
fast = bool(os.environ.get("QP_PROBE_FAST"))            # under ncu: device launches only
for variant in os.environ.get("QP_PROBE_VARIANTS", "tab,simple").split(","):
    os.environ["PSB200_QP"] = variant

    def launch():
        rc = L.psb200_quickpol_xi_dev(nu1, nu2, s1, s2, lmax, dW.data_ptr(), W.size, band, band, dX.data_ptr(), nb,
                                      0, lmax + 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, L.psb200_last_error()
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    x = dX.cpu().numpy()
    if ref is None:
        ref = x.copy()
    if fast:
        res[variant] = {"ms": ms, "lib": os.environ.get("PSB200_LIB", "default"), "abs_sum": float(np.abs(x).sum()),
                        "sum": float(x.sum())}
        continue
    Xb = np.zeros((nb, lmax + 1), order="F")
    wp, xp = W.ctypes.data_as(ps._lib.DP), Xb.ctypes.data_as(ps._lib.DP)
    L.psb200_quickpol_xi(nu1, nu2, s1, s2, lmax, wp, W.size, band, band, xp, nb, 1)
    t0 = time.perf_counter()
    for _ in range(3):
        rc = L.psb200_quickpol_xi(nu1, nu2, s1, s2, lmax, wp, W.size, band, band, xp, nb, 1)
        assert rc == 0
    e2e = (time.perf_counter() - t0) / 3 * 1e3
    res[variant] = {"ms": ms, "terms_per_s": terms / (ms * 1e-3), "e2e_ms": e2e,
                    "max_abs_diff_vs_tab": float(np.max(np.abs(x - ref))),
                    "host_call_equals_device_call": bool(np.array_equal(Xb.T, x))}

# CPU oracle on every k-th row (band storage makes a row sample awkward: time a smaller lmax slice of the same
# band and scale by the exact term ratio is NOT done -- the sample is the low-l part, whose families are shorter,
# so the oracle's terms/s is reported on its own term count)
if fast:
    print(json.dumps(res))
    sys.exit(0)
lm_cpu = min(lmax, 3071)
t0 = time.perf_counter()
_, t_cpu = po.quickpol_xi(nu1, nu2, s1, s2, lm_cpu, W[: 2 * lm_cpu + 1], band, band, dense=False, return_terms=True)
dt = time.perf_counter() - t0
res["cpu_oracle"] = {"lmax": lm_cpu, "terms": int(t_cpu), "s": dt, "terms_per_s": t_cpu / dt, "threads": po.max_threads(),
                     "kind": "port (C/OpenMP restatement of src/beam.jl:72-101, not Julia)"}
res["peak_dfma_tflops"] = L.psb200_dfma_peak(4000) / 1e12
line = json.dumps(res)
print(line)
if out:
    with open(out, "w") as f:
        f.write(line + "\n")
