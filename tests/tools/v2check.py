"""Parity report: GPU (tuned kernel) and the reference-shaped Float64 oracle, both against the
long-double oracle, on every entry above 1e-30 of its row maximum (two distinct masks => the
cross-spectrum changes sign => cancelling sums).  Columns:
  worst/bound : max |x - ref| / (1e-10 |ref| + 1e-13 S_abs)      (the test criterion, <= 1 passes)
  strict max  : max |x - ref| / |ref|                            (north-star form)
  frac>1e-10  : fraction of entries whose strict relative error exceeds 1e-10
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import powerspectra_jl_b200 as ps
from oracle import psoracle as po
from powerspectra_jl_b200 import synthetic as syn
from conftest import parity_worst

def stats(X, R, S, lo):
    X, R, S = X[lo:, lo:], R[lo:, lo:], S[lo:, lo:]
    sel = np.abs(R) > 1e-30 * np.abs(R).max(axis=1, keepdims=True)
    rel = np.abs(X[sel] - R[sel]) / np.abs(R[sel])
    return parity_worst(X, R, S), rel.max(), float((rel > 1e-10).mean())

lmaxes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [40, 300, 767]
print(f"{'lmax':>5} {'kind':>4} | {'GPU worst/bound':>15} {'strict max':>11} {'frac>1e-10':>11} | {'F64 oracle w/b':>15} {'strict max':>11} {'frac>1e-10':>11}")
for lmax in lmaxes:
    V = syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)]
    for spec, kind in (("TT", 0), ("TE", 1), ("M++", 2), ("M--", 3)):
        M = ps.mcm(spec, ps.SpectralVector(V)).parent
        R = po.mcm(kind, 0, lmax, V, ld=True)
        D = po.mcm(kind, 0, lmax, V)
        with po.abs_mode():
            S = po.mcm(kind, 0, lmax, V)
        lo = 2 if kind else 0
        g, d = stats(M, R, S, lo), stats(D, R, S, lo)
        print(f"{lmax:5d} {spec:>4} | {g[0]:15.4f} {g[1]:11.2e} {g[2]:11.2e} | {d[0]:15.4f} {d[1]:11.2e} {d[2]:11.2e}", flush=True)
