"""One benchmark step through the host calls (psb200_mcm / psb200_cov, page-locked result arrays of the library) with the
standard delivery and with the mirror delivery (PSB200_MIRROR=1), in a process of its own: bench.py runs it as a
subprocess at several GPUs so that the optional mirror path can never cost the bench line.  No torch.
    python tools/e2e_step_probe.py [ngpus] [steps] [lmax]      -> one JSON line"""
import json, os, sys, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import powerspectra_jl_b200 as ps

ngpus = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lmax = int(sys.argv[3]) if len(sys.argv) > 3 else 6143
N = lmax + 1
L, DP = ps.lib(), ps._lib.DP
inp = bench.make_inputs(lmax)
ptrs = lambda arrs: (DP * max(len(arrs), 1))(*[a.ctypes.data_as(DP) for a in arrs])


def alloc():
    return {name: [ps._lib.HostMatrix(N) for _ in range(2 if (api == "mcm" and code == 4) else 1)]
            for name, api, code, fam, _ in bench.JOBS}


def host_calls(dst):
    for name, api, code, fam, _ in bench.JOBS:
        a, O = inp[name], [h.array for h in dst[name]]
        if api == "mcm":
            rc = L.psb200_mcm(code, 0, lmax, a["V"].ctypes.data_as(DP), a["V"].size, O[0].ctypes.data_as(DP), N,
                              O[1].ctypes.data_as(DP) if len(O) > 1 else None, ngpus)
        else:
            rc = L.psb200_cov(code, 0, lmax, ptrs(a["sp"]), len(a["sp"]), ptrs(a["rt"]), len(a["rt"]),
                              ptrs(a["W"]), len(a["W"]), a["W"][0].size, O[0].ctypes.data_as(DP), N, ngpus)
        ps._lib.check(rc)


def timed(dst):
    host_calls(dst)
    t0 = time.perf_counter()
    for _ in range(steps):
        host_calls(dst)
    return (time.perf_counter() - t0) * 1e3 / steps


std, mir = alloc(), alloc()
os.environ.pop("PSB200_MIRROR", None)
t_std = timed(std)
os.environ["PSB200_MIRROR"] = "1"
t_mir = timed(mir)
equal = all(np.array_equal(a.array, b.array) for name in std for a, b in zip(std[name], mir[name]))
print(json.dumps({"ngpus": ngpus, "lmax": lmax, "steps": steps, "standard_ms_per_step": t_std, "mirror_ms_per_step": t_mir,
                  "equals_standard_delivery": bool(equal)}), flush=True)
