"""Small run of every job (for compute-sanitizer): MCM kinds, fused kinds, all covariance blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import powerspectra_jl_b200 as ps
from powerspectra_jl_b200 import synthetic as syn
lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 150
V = ps.SpectralVector(syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)])
for spec in ("TT", "TE", "M++", "M--", "EE_BB"):
    ps.mcm(spec, V, lmin=2)
ws, sp, rt = syn.covariance_inputs(lmax)
for a, b in (("TT", "TT"), ("EE", "EE"), ("TE", "TE"), ("TT", "TE"), ("TT", "EE"), ("TE", "EE")):
    ps.coupledcov(a, b, ws, sp, rt)
import powerspectra_jl_b200.covariance as cv
cv.coupledcovTEEE(ps.spectralzeros(range(0, lmax + 1), range(0, lmax + 1)), ws, sp, rt, planck=False)
sky = syn.ZonalSky(lmax)
al = [ps.Alm.zonal(a) for a in sky.al0([syn.mask_profile(sky.theta, s) for s in (1001, 1002, 1003, 1004)])]
ps.mcm_master(*al)
print("done")
