"""Fuzzes the arithmetic of the QuickPol CUDA pair function (host build, tests/hostcheck) against the long-double
oracle on random admissible pairs, any spin up to the angular momentum.  CPU only.

  python tests/tools/quickpol_fuzz.py [seed] [count] [lmax]

Round 1: seed 1, 20000 pairs, l < 120 -> worst error / bound 0.0056; seed 2, 3000 pairs, l < 3000 -> 0.011; no failure.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostcheck as hc                                 # noqa: E402
from oracle import psoracle as po                      # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
count = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
lmax = int(sys.argv[3]) if len(sys.argv) > 3 else 120
rng = np.random.default_rng(seed)
bad, worst = 0, 0.0
for _ in range(count):
    l, lpp = int(rng.integers(2, lmax)), int(rng.integers(2, lmax))
    s1, s2 = int(rng.integers(-l, l + 1)), int(rng.integers(-l, l + 1))
    if rng.random() < 0.5:
        s1 = int(np.clip(rng.integers(-6, 7), -l, l))
    if rng.random() < 0.5:
        s2 = int(np.clip(rng.integers(-6, 7), -l, l))
    nu1, nu2 = (int(np.clip(rng.integers(-3, 4), -lpp, lpp)) for _ in range(2))
    W = rng.normal(size=int(rng.integers(1, 2 * lmax + 20)))
    n1, f1 = po.w3j_family(l, lpp, -s1, -nu1, ld=True)
    n2, f2 = po.w3j_family(l, lpp, -s2, -nu2, ld=True)
    a, e = max(n1, n2), min(l + lpp, W.size - 1)
    ref = sab = 0.0
    if e >= a and f1.size and f2.size:
        j = np.arange(a, e + 1)
        t = W[j] * f1[j - n1] * f2[j - n2]
        ref, sab = (-1.0) ** ((s1 + s2 + nu1 + nu2) % 2) * t.sum(), np.abs(t).sum()
    for v in ("tab", "simple"):
        got = hc.pair(l, lpp, nu1, nu2, s1, s2, W, variant=v)
        r = abs(got - ref) / (1e-10 * abs(ref) + 1e-13 * sab + 1e-300)
        worst = max(worst, r)
        if not r <= 1:
            bad += 1
            print("FAIL", v, (l, lpp, s1, nu1, s2, nu2, W.size), got, ref, r)
print(f"seed {seed}, {count} pairs, l < {lmax}: worst error / bound {worst:.4f}, failures {bad}")
