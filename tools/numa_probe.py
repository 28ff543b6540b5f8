"""Result arrays on one NUMA node against arrays interleaved over the nodes (psb200_host_alloc policy 0 / 1) for the host
calls on every visible GPU; no torch.  python tools/numa_probe.py [lmax]"""
import ctypes, json, os, sys, time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import powerspectra_jl_b200 as ps
from powerspectra_jl_b200 import synthetic as syn

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6143
N = lmax + 1
L, DP = ps.lib(), ps._lib.DP
ng = L.psb200_device_count()
V = np.ascontiguousarray(syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)])
print(json.dumps({"gpus": ng, "numa_nodes": L.psb200_host_numa_nodes(), "cpus": os.cpu_count(),
                  "nodes_online": open("/sys/devices/system/node/online").read().strip()}), flush=True)


def call(kind, X, Y, ngpus, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ps._lib.check(L.psb200_mcm(kind, 0, lmax, V.ctypes.data_as(DP), V.size, X.ctypes.data_as(DP), N,
                                   Y.ctypes.data_as(DP) if kind == 4 else None, ngpus))
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), float(np.median(ts))


ref = {}
for policy, nombind in ((0, ""), (1, ""), (1, "1")):
    if nombind:
        os.environ["PSB200_NO_MBIND"] = "1"          # first touch by node-bound threads instead of mbind
    t0 = time.perf_counter()
    HA, HB = ps._lib.HostMatrix(N, bool(policy)), ps._lib.HostMatrix(N, bool(policy))
    row = {"policy": policy, "first_touch_only": bool(nombind), "alloc_ms": (time.perf_counter() - t0) * 1e3,
           "placement": HA.placement()}
    for kind in (0, 4):
        for g in sorted({1, ng}):
            call(kind, HA.array, HB.array, g, 1)
            row[f"kind{kind}_ngpus{g}_ms_min_median"] = call(kind, HA.array, HB.array, g)
            key = (kind,)
            if key not in ref:
                ref[key] = (HA.array.copy(), HB.array.copy())
            assert np.array_equal(HA.array, ref[key][0]) and (kind != 4 or np.array_equal(HB.array, ref[key][1]))
    HA.free(), HB.free()
    os.environ.pop("PSB200_NO_MBIND", None)
    print(json.dumps(row), flush=True)
