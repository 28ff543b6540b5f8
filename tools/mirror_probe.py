"""Mirror delivery (PSB200_MIRROR=1: block columns only over PCIe, the symmetric side written by the host scatter threads)
against the standard delivery: bit-identical results and ms per call.  No torch.  python tools/mirror_probe.py [lmax] [ngpus]"""
import json, os, sys, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import powerspectra_jl_b200 as ps
import highl_inputs

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6143
NG = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = lmax + 1
L, DP = ps.lib(), ps._lib.DP
V = np.ascontiguousarray(np.load(os.path.join(ROOT, "tests", "golden", "mcm_entries_mp.npz"))["V_6143"][:N])
sp, rt, W = highl_inputs.cov_inputs(lmax)["TTTT"]
ptrs = lambda arrs: (DP * len(arrs))(*[a.ctypes.data_as(DP) for a in arrs])


def run(what, X, Y, reps):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        if what == "TTTT":
            rc = L.psb200_cov(0, 0, lmax, ptrs(sp), 4, ptrs(rt), 4, ptrs(W), 8, W[0].size, X.ctypes.data_as(DP), N, NG)
        else:
            kind = 0 if what == "TT" else 4
            rc = L.psb200_mcm(kind, 0, lmax, V.ctypes.data_as(DP), V.size, X.ctypes.data_as(DP), N,
                              Y.ctypes.data_as(DP) if kind == 4 else None, NG)
        ps._lib.check(rc)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


HA, HB = ps._lib.HostMatrix(N), ps._lib.HostMatrix(N)
PA, PB = np.zeros((N, N), order="F") + 0.0, np.zeros((N, N), order="F") + 0.0
for what in ("TT", "EE_BB", "TTTT"):
    os.environ.pop("PSB200_MIRROR", None)
    run(what, HA.array, HB.array, 1)
    row = {"lmax": lmax, "ngpus": NG, "call": what, "standard_page_locked_ms": run(what, HA.array, HB.array, 3)}
    refA, refB = HA.array.copy(), HB.array.copy()
    row["standard_pageable_ms"] = run(what, PA, PB, 2)
    os.environ["PSB200_MIRROR"] = "1"
    HA.array[:] = 0.0; HB.array[:] = 0.0; PA[:] = 0.0; PB[:] = 0.0
    row["mirror_page_locked_ms"] = run(what, HA.array, HB.array, 3)
    row["mirror_pageable_ms"] = run(what, PA, PB, 2)
    row["bit_identical"] = bool(np.array_equal(HA.array, refA) and np.array_equal(PA, refA)
                                and (what != "EE_BB" or (np.array_equal(HB.array, refB) and np.array_equal(PB, refB))))
    print(json.dumps(row), flush=True)
