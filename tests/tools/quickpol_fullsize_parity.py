"""Full-size parity evidence for the QuickPol kernel without a GPU in the loop (run in the build container):

1. the long-double oracle at lmax 6143, band +-128, on the probe's inputs (tests/tools/quickpol_probe.py);
2. the host build of the kernel arithmetic (tests/hostcheck, same source as the CUDA kernel) against it;
3. the checksums the B200 run of the probe printed for the same inputs (profiles/r01_quickpol_ab.jsonl)
   against the checksums of (2) -- ties the device output at full size to (1).

  python tests/tools/quickpol_fullsize_parity.py [lmax] [band] > profiles/r01_quickpol_parity_fullsize.txt
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostcheck as hc                                 # noqa: E402
from oracle import psoracle as po                      # noqa: E402

lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 6143
band = int(sys.argv[2]) if len(sys.argv) > 2 else 128
case = (2, -2, 2, 2)
rng = np.random.default_rng(7)
l = np.arange(2 * lmax + 1)
W = rng.normal(size=l.size) / (1.0 + l / 40.0) ** 2 + 1.0 / (1.0 + l) ** 1.5      # as tests/tools/quickpol_probe.py

t = time.time()
ref = po.quickpol_xi(*case, lmax, W, band, band, ld=True, dense=False)
with po.abs_mode():
    sabs = po.quickpol_xi(*case, lmax, W, band, band, ld=True, dense=False)
print(f"long-double oracle, lmax={lmax}, band=+-{band}, case (nu1,nu2,s1,s2)={case}: {time.time() - t:.0f} s")
for variant in ("tab", "simple"):
    x = hc.xi_band(*case, lmax, W, band, band, variant=variant)
    err = np.abs(x - ref)
    ratio = err / (1e-10 * np.abs(ref) + 1e-13 * sabs + 1e-300)
    sel = (np.abs(ref) > 0) & (sabs <= 1e3 * np.abs(ref))
    print(f"host build of the kernel arithmetic [{variant}]: entries {np.count_nonzero(ref)}, "
          f"max |err| {err.max():.3e}, max err/(1e-10|ref|+1e-13 S_abs) {ratio.max():.4f}, "
          f"strict rel err on the {sel.sum()} non-cancelling entries {np.max(err[sel] / np.abs(ref[sel])):.3e}")
    print(f"  checksums: sum |x| = {np.abs(x).sum()!r}, sum x = {x.sum()!r}")
ab = os.path.join(ROOT, "profiles", "r01_quickpol_ab.jsonl")
if os.path.exists(ab) and (lmax, band) == (6143, 128):
    d = json.loads(open(ab).readline())
    x = hc.xi_band(*case, lmax, W, band, band, variant="tab")
    for variant in ("tab", "simple"):
        g = d[variant]
        print(f"B200 run [{variant}] ({g['lib']} build): sum |x| = {g['abs_sum']!r}, sum x = {g['sum']!r}; "
              f"relative difference to the host build: {abs(g['abs_sum'] - np.abs(x).sum()) / np.abs(x).sum():.2e}, "
              f"{abs(g['sum'] - x.sum()) / np.abs(x).sum():.2e}")
