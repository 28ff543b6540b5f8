"""Sampled-row parity of the GPU MCMs against the long-double oracle at large lmax (default 12287)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import powerspectra_jl_b200 as ps
from oracle import psoracle as po
from powerspectra_jl_b200 import synthetic as syn
from conftest import parity_worst
lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 12287
rstep = int(sys.argv[2]) if len(sys.argv) > 2 else 768
V = syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)]
rows = np.arange(5, lmax + 1, rstep)
sel = np.zeros(lmax + 1, bool); sel[rows] = True
t0 = time.time()
M00 = ps.mcm("TT", ps.SpectralVector(V)).parent
both = ps.mcm("EE_BB", ps.SpectralVector(V))
print(f"GPU TT + EE_BB at lmax={lmax}: {time.time()-t0:.2f} s (host call incl. allocation)")
for name, G, k in (("TT", M00, 0), ("M++", both.getblock(0, 0).parent, 2), ("M--", both.getblock(0, 1).parent, 3)):
    R = po.mcm(k, 0, lmax, V, row0=5, rstep=rstep, ld=True)
    with po.abs_mode():
        S = po.mcm(k, 0, lmax, V, row0=5, rstep=rstep)
    Gu, Ru, Su = np.triu(G)[sel][:, 2:], np.triu(R)[sel][:, 2:], np.triu(S)[sel][:, 2:]
    m = np.abs(Ru) > 1e-30 * np.abs(Ru).max(axis=1, keepdims=True)
    rel = np.abs(Gu[m] - Ru[m]) / np.abs(Ru[m])
    well = m & (np.abs(Su) <= 1e3 * np.abs(Ru))
    relw = np.abs(Gu[well] - Ru[well]) / np.abs(Ru[well])
    print(f"{name:4s} rows={rows.size} entries={m.sum()}  err/bound={parity_worst(Gu, Ru, Su):.4f}  strict max={rel.max():.2e} "
          f"frac>1e-10={np.mean(rel > 1e-10):.2e}  strict max on well-conditioned entries={relw.max():.2e}", flush=True)
