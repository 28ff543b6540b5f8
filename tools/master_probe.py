"""Times the fused psb200_mcm_master_dev pass against the five separate stage-1 calls it replaces."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from powerspectra_jl_b200 import device as dev, synthetic as syn
lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6143
N = lmax + 1
Vs = syn.mask_spectra(lmax, seeds=(1001, 1002, 1003, 1004))
V = [torch.tensor(Vs[k], device="cuda") for k in ((0, 2), (0, 3), (1, 2), (1, 3))]
X = [torch.empty((N, N), dtype=torch.float64, device="cuda") for _ in range(5)]
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fused = t(lambda: dev.master_slab(0, lmax, V[0], V[1], V[2], V[3], X))
def sep():
    dev.mcm_slab(0, 0, lmax, V[0], X[0]); dev.mcm_slab(1, 0, lmax, V[1], X[1]); dev.mcm_slab(1, 0, lmax, V[2], X[2])
    dev.mcm_slab(4, 0, lmax, V[3], X[3], X[4])
separate = t(sep)
print(f"lmax={lmax}: fused master pass {fused:.2f} ms, separate TT+TE+ET+EE/BB {separate:.2f} ms, ratio {separate / fused:.2f}")
