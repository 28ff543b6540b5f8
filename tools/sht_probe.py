"""Times the spin-0 HEALPix transforms (psb200_sht.cuh) device-resident on one GPU.

    [PSB200_SHT_R=2|4|8] python tools/sht_probe.py [nside] [lmax] >> gpurun_out/sht_probe.jsonl

One JSON line: ms of one analysis (map2alm niter 0), one synthesis (alm2map), map2alm with niter 3, the executed-work
fraction of the DFMA peak of the two Legendre passes, and a band-limited round-trip error.  Development tool."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import powerspectra_jl_b200 as ps
from powerspectra_jl_b200 import device as dev

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 3 * nside - 1
once = bool(os.environ.get("PROBE_ONCE"))
L = ps.lib()
g = torch.Generator(device="cuda").manual_seed(1)
n = dev.alm_size(lmax)
alm = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)
ls = torch.tensor(np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)]), device="cuda")
alm *= torch.exp(-0.5 * (ls / (0.25 * lmax)) ** 2)
alm[:lmax + 1] = alm[:lmax + 1].real.to(torch.complex128)
f = torch.empty(12 * nside * nside, dtype=torch.float64, device="cuda")
back = torch.empty_like(alm)


def dtime(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    if once:
        return 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_syn = dtime(lambda: dev.alm2map_dev(nside, lmax, alm, f))
t_ana = dtime(lambda: dev.map2alm_dev(nside, lmax, f, back, 0))
err0 = float((back - alm).abs().max() / alm.abs().max())
t_m2a = dtime(lambda: dev.map2alm_dev(nside, lmax, f, back, 3), reps=2)
err3 = float((back - alm).abs().max() / alm.abs().max())
st = dev.sht_stats(nside, lmax)
rec = {"nside": nside, "lmax": lmax, "R": st["R"], "ms_alm2map": t_syn, "ms_analysis": t_ana, "ms_map2alm_niter3": t_m2a,
       "roundtrip_err_niter0": err0, "roundtrip_err_niter3": err3, "stats": st}
if not once:
    peak = L.psb200_dfma_peak(20000) / 2.0           # FP64 lane-instructions / s
    rec["dfma_peak_tflops"] = 2 * peak / 1e12
    rec["frac_synthesis_pass"] = 4.0 * st["exec_steps"] / (t_syn * 1e-3) / peak          # includes the ring stage in the time
    rec["frac_analysis_pass"] = 4.0 * st["exec_steps"] / (t_ana * 1e-3) / peak
    rec["frac_map2alm"] = 7 * 4.0 * st["exec_steps"] / (t_m2a * 1e-3) / peak
    rec["useful_over_exec"] = st["live_steps"] / st["exec_steps"]
print(json.dumps(rec))
