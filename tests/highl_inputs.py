"""Inputs of the full-size covariance known answers (tests/golden/cov_entries_mp.npz), shared by the generator
(tests/golden/make_golden_highl.py) and the tests.  Only +, -, *, /, sqrt and fmod on doubles: every one of them is
correctly rounded by IEEE 754, so the vectors are bit-identical on any machine (the fixture stores their SHA-256)."""
import hashlib

import numpy as np


def cov_inputs(lmax):
    l = np.arange(lmax + 1, dtype=np.float64)
    tt = [(6000.0 + 500.0 * k) / ((l + 12.0) * (l + 13.0 + k)) + 1e-7 * (k + 1) for k in range(4)]      # positive
    ee = [0.02 * t * (1.0 + 0.05 * k) for k, t in enumerate(tt)]
    te = [0.1 * tt[k] * (np.fmod(l + k, 9.0) - 4.0) / 4.0 for k in range(4)]                            # changes sign
    r = [np.sqrt(1.0 + (l / (1500.0 + 200.0 * k)) * (l / (1500.0 + 200.0 * k))) for k in range(4)]
    W = [(1.0 + 0.125 * q) * (1.0 - 0.25 * np.fmod(l + q, 3.0)) / ((1.0 + l / (40.0 + 8.0 * q)) * (1.0 + l / (40.0 + 8.0 * q)))
         for q in range(8)]
    blocks = {
        "TTTT": (tt, r, W),
        "EEEE": (ee, r, W),
        "TETE": ([tt[0], ee[1], te[2], te[3]], [r[0], r[1]], W[:5]),
        "TTTE": ([tt[0], tt[3], te[2], te[1]], [r[0], r[3]], W[:4]),
        "TEEE_planck": ([ee[1], ee[3], te[0], te[2]], [r[1], r[3]], W[:4]),
        "TEEE": ([ee[1], ee[3], te[0], te[2]], [r[1], r[3]], W[:4]),
        "TTEE": ([te[0], te[2], te[1], te[3]], [], W[:2]),
    }
    return {k: tuple([np.ascontiguousarray(x) for x in part] for part in v) for k, v in blocks.items()}


def quickpol_window(lmax):
    """W_l' of the QuickPol known answers (what quickpolW hands to quickpolXi!, src/beam.jl:43-56): changes sign."""
    l = np.arange(lmax + 1, dtype=np.float64)
    return np.ascontiguousarray((np.fmod(l, 5.0) - 1.5) / ((1.0 + l / 200.0) * (1.0 + l / 200.0)))


def digest(inputs):
    h = hashlib.sha256()
    for name in sorted(inputs):
        for part in inputs[name]:
            for x in part:
                h.update(np.ascontiguousarray(x, dtype=np.float64).tobytes())
    return h.hexdigest()
