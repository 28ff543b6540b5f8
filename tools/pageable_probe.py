"""Host calls into PAGEABLE result arrays: the library's staged delivery against the CUDA runtime's bounce copies, over
worker counts and chunk sizes (no torch: starts in seconds).  python tools/pageable_probe.py [lmax] [ngpus] [quick]"""
import ctypes, json, os, sys, time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import powerspectra_jl_b200 as ps
from powerspectra_jl_b200 import synthetic as syn

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6143
NG = int(sys.argv[2]) if len(sys.argv) > 2 else 1
QUICK = len(sys.argv) > 3
N = lmax + 1
L, DP = ps.lib(), ps._lib.DP
V = np.ascontiguousarray(syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)])
A, B = np.zeros((N, N), order="F") + 0.0, np.zeros((N, N), order="F") + 0.0
HA, HB = ps._lib.HostMatrix(N), ps._lib.HostMatrix(N)


def call(kind, X, Y, reps=4):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        ps._lib.check(L.psb200_mcm(kind, 0, lmax, V.ctypes.data_as(DP), V.size, X.ctypes.data_as(DP), N,
                                   Y.ctypes.data_as(DP) if kind == 4 else None, NG))
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


for kind in (0, 4):
    call(kind, HA.array, HB.array, 1)
    row = {"lmax": lmax, "ngpus": NG, "kind": kind, "page_locked_ms": call(kind, HA.array, HB.array)}
    refA, refB = HA.array.copy(), HB.array.copy()
    os.environ["PSB200_STAGED"] = "0"
    row["runtime_bounce_ms"] = call(kind, A, B)
    os.environ["PSB200_STAGED"] = "1"
    for nt in ((1,) if QUICK else (0, 1)):
        for thr in ((4, 8) if QUICK else (2, 4, 8, 12, 15)):
            for mb in (4, 8, 16) if thr == 12 else (8,):
                os.environ["PSB200_STAGE_THREADS"], os.environ["PSB200_STAGE_CHUNK_MB"] = str(thr), str(mb)
                os.environ["PSB200_STAGE_NT"] = str(nt)
                A[:] = 0.0
                row[f"staged_nt{nt}_t{thr}_c{mb}_ms"] = call(kind, A, B)
                assert np.array_equal(A, refA) and (kind != 4 or np.array_equal(B, refB))
    del os.environ["PSB200_STAGE_THREADS"], os.environ["PSB200_STAGE_CHUNK_MB"], os.environ["PSB200_STAGE_NT"]
    row["staged_default_ms"] = call(kind, A, B)
    print(json.dumps(row), flush=True)
