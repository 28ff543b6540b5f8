"""Times the host-level C-ABI call phase by phase (PSB200_TRACE=1) for pageable and pinned outputs."""
import os, sys, time
os.environ["PSB200_TRACE"] = "1"
sys.path.insert(0, ".")
import numpy as np, torch
import powerspectra_jl_b200 as ps
from powerspectra_jl_b200 import synthetic as syn
lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6143
ngpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N = lmax + 1
V = syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)]
L, DP = ps.lib(), ps._lib.DP
outs = {"pageable": np.zeros((N, N), order="F"), "pinned": torch.empty((N, N), dtype=torch.float64).pin_memory().numpy()}
M2 = torch.empty((N, N), dtype=torch.float64).pin_memory().numpy()
for name, M in outs.items():
    for rep in range(3):
        t = time.perf_counter()
        rc = L.psb200_mcm(kind, 0, lmax, V.ctypes.data_as(DP), V.size, M.ctypes.data_as(DP), N, M2.ctypes.data_as(DP) if kind == 4 else None, ngpus)
        print(f"== {name} rep {rep}: rc={rc} total {1e3*(time.perf_counter()-t):.1f} ms", flush=True)
