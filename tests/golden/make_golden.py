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
4. w3j_general_exact.npz -- exact 3j values (sympy) of the general-spin families the QuickPol
   path evaluates, f(j) = (j l l''; s+nu, -s, -nu) (/root/reference/src/beam.jl:86-93), for a grid of
   (s, nu) including the degenerate ones (B(j) = 0 for every j; |s| = l; one-term families), and
   quickpol_xi_exact: the full Xi matrix of src/beam.jl:72-101 at lmax = 10 from exact 3j symbols
   for five (nu1, nu2, s1, s2) and one seeded W.
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


def make_w3j_general():
    from sympy import N as SN
    from sympy.physics.wigner import wigner_3j
    combos = [(s, nu) for s in (-3, -2, -1, 0, 1, 2, 3, 5) for nu in (-2, 0, 2)]
    pairs = [(2, 2), (2, 3), (3, 2), (4, 4), (5, 7), (7, 5), (6, 6), (9, 12), (12, 12), (16, 11), (20, 20), (5, 30),
             (30, 28)]
    rows, vals = [], []
    for (s, nu) in combos:
        for (l, lpp) in pairs:
            if abs(s) > l or abs(nu) > lpp:
                continue
            m1 = s + nu
            lo = max(abs(l - lpp), abs(m1))
            v = [float(SN(wigner_3j(j, l, lpp, m1, -s, -nu), 30)) for j in range(lo, l + lpp + 1)]
            rows.append((l, lpp, -s, -nu, lo, len(vals), len(v)))
            vals.extend(v)
    # exact Xi matrices at lmax = 10, full band
    lmax = 10
    rng = np.random.default_rng(20240611)
    W = rng.normal(size=2 * lmax + 1)
    cases = [(0, 0, 0, 0), (2, 2, 2, 2), (-2, 2, 0, 1), (0, 2, 3, -1), (2, -2, -2, 2)]
    cache = {}

    def fam(l, lpp, s, nu):
        key = (l, lpp, s, nu)
        if key not in cache:
            m1 = s + nu
            lo = max(abs(l - lpp), abs(m1))
            cache[key] = (lo, np.array([float(SN(wigner_3j(j, l, lpp, m1, -s, -nu), 30))
                                        for j in range(lo, l + lpp + 1)]))
        return cache[key]
    xis = np.zeros((len(cases), lmax + 1, lmax + 1))
    for c, (nu1, nu2, s1, s2) in enumerate(cases):
        sgn = -1.0 if (s1 + s2 + nu1 + nu2) % 2 else 1.0
        for lpp in range(2, lmax + 1):
            for l in range(2, lmax + 1):
                if abs(s1) > l or abs(s2) > l or abs(nu1) > lpp or abs(nu2) > lpp:
                    continue
                lo1, f1 = fam(l, lpp, s1, nu1)
                lo2, f2 = fam(l, lpp, s2, nu2)
                a = max(lo1, lo2)
                js = np.arange(a, l + lpp + 1)
                if js.size:
                    xis[c, lpp, l] = sgn * np.sum(W[js] * f1[js - lo1] * f2[js - lo2])
    np.savez_compressed(os.path.join(HERE, "w3j_general_exact.npz"), index=np.array(rows, dtype=np.int64),
                        values=np.array(vals), xi_lmax=lmax, xi_W=W, xi_cases=np.array(cases, dtype=np.int64),
                        xi=xis)
    print("w3j_general_exact:", len(rows), "families,", len(vals), "values;", len(cases), "Xi matrices")


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
    make_w3j_general()
