import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import powerspectra_jl_b200 as ps
from oracle import psoracle as po
from powerspectra_jl_b200 import synthetic as syn
from conftest import parity_worst, parity_error
for lmax in (40, 300, 767):
    V = syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)]
    for spec, kind in (("TT", 0), ("TE", 1), ("M++", 2), ("M--", 3)):
        M = ps.mcm(spec, ps.SpectralVector(V)).parent
        R = po.mcm(kind, 0, lmax, V, ld=True)
        with po.abs_mode(): S = po.mcm(kind, 0, lmax, V)
        lo = 2 if kind else 0
        print(lmax, spec, "worst/bound %.3g" % parity_worst(M[lo:,lo:], R[lo:,lo:], S[lo:,lo:]), "strict rel %.3g" % parity_error(M[lo:,lo:], R[lo:,lo:]), flush=True)
