"""Generates the committed golden fixtures of tests/golden/.  Run HERE (the build container),
where /root/reference exists; the fixtures then travel with the repo.

  python tests/golden/make_golden.py

1. w3j_exact.npz -- exact Wigner 3j values from sympy (rational arithmetic, evaluated to 30
   digits) for the two families of the hot path, f00 = (j l1 l2; 0 0 0), f22 = (j l1 l2; 0 -2 2):
   every pair 0 <= l1 <= l2 <= 24 (f22 only where |m| <= l) plus a few pairs up to l = 80.
2. namaster_diag.npz -- the three NaMaster golden diagonals the reference's own tests hold
   (/root/reference/test/data/mcm_TT_diag.txt, mcm_EE_diag.txt, mcm_TE_diag.txt; used at
   test/test_mcm.jl:12-50; 765 values each, l = 2..766, nside 256, one mask), stored verbatim,
   plus V_even: the even-l mask power spectrum (l3 = 0, 2, ..., 766) recovered from them.
   On the diagonal (l1 = l2) only even l3 contribute, so the 3 x 765 goldens are 2295 linear
   equations in 384 unknowns; the least-squares solution under the ORACLE's kernels fits all
   of them to ~2e-14 relative, which pins the oracle's f00^2, f22^2 and f00*f22 (even parity,
   diagonal pairs) to NaMaster.  tests/test_oracle.py re-derives the diagonals from V_even.
3. theory_noise_767.npz -- cltt, clte, clee, nltt, nlee (l = 0..767) of test/data/theory.csv and
   noise.csv, the spectra of the reference's covariance test (test/test_covmat.jl:38-45).
"""
import csv
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference/test/data/"


def make_w3j():
    from sympy import N as SN
    from sympy.physics.wigner import wigner_3j
    pairs = [(a, b) for a in range(0, 25) for b in range(a, 25)]
    pairs += [(30, 45), (40, 40), (2, 60), (17, 80), (64, 64), (3, 77), (80, 80)]
    rows = []
    vals = []
    for fam, (m2, m3) in enumerate([(0, 0), (-2, 2)]):
        for (l1, l2) in pairs:
            if fam == 1 and l1 < 2:
                continue
            js = range(abs(l1 - l2), l1 + l2 + 1)
            v = [float(SN(wigner_3j(j, l1, l2, -m2 - m3, m2, m3), 30)) for j in js]
            rows.append((fam, l1, l2, len(vals), len(v)))
            vals.extend(v)
    np.savez_compressed(os.path.join(HERE, "w3j_exact.npz"), index=np.array(rows, dtype=np.int64),
                        values=np.array(vals))
    print("w3j_exact:", len(rows), "families,", len(vals), "values")


def make_namaster():
    from oracle import psoracle as po
    tt = np.loadtxt(REF + "mcm_TT_diag.txt")
    ee = np.loadtxt(REF + "mcm_EE_diag.txt")
    te = np.loadtxt(REF + "mcm_TE_diag.txt")
    lmax = 767
    ells = np.arange(2, 767)
    K = np.zeros((3, ells.size, 384))
    for i, l in enumerate(ells):
        _, f0 = po.w3j_family(int(l), int(l), 0, 0, ld=True)
        _, f2 = po.w3j_family(int(l), int(l), -2, 2, ld=True)
        j = np.arange(0, 2 * l + 1)
        sel = (j % 2 == 0) & (j <= lmax)
        jj = j[sel]
        pref = (2 * l + 1) / (4 * np.pi) * (2 * jj + 1)
        K[0, i, jj // 2] = pref * f0[sel] ** 2
        K[1, i, jj // 2] = pref * f2[sel] ** 2
        K[2, i, jj // 2] = pref * f0[sel] * f2[sel]
    A = np.vstack([K[0] / tt[:, None], K[1] / ee[:, None], K[2] / te[:, None]])
    V, _, rank, sv = np.linalg.lstsq(A, np.ones(A.shape[0]), rcond=None)
    r = A @ V - 1.0
    print("namaster_diag: rank", rank, "cond %.3g" % (sv[0] / sv[-1]), "max rel residual %.3g" % np.abs(r).max())
    np.savez_compressed(os.path.join(HERE, "namaster_diag.npz"), tt=tt, ee=ee, te=te, V_even=V)


def make_theory():
    def col(fn):
        with open(REF + fn) as f:
            rd = csv.DictReader(f)
            rows = list(rd)
        return {k.strip(): np.array([float(r[k]) for r in rows]) for k in rows[0].keys() if k.strip()}
    th, no = col("theory.csv"), col("noise.csv")
    np.savez_compressed(os.path.join(HERE, "theory_noise_767.npz"), cltt=th["cltt"], clte=th["clte"],
                        clee=th["clee"], nltt=no["nltt"], nlee=no["nlee"])
    print("theory_noise_767:", {k: v.size for k, v in th.items()}, {k: v.size for k, v in no.items()})


if __name__ == "__main__":
    make_theory()
    make_namaster()
    make_w3j()
