"""Pins the CPU oracle (oracle/psoracle.c) -- runs on CPU, no GPU needed.

The oracle restates WignerFamilies.jl's family evaluation and the reference loops
(/root/reference/src/modecoupling.jl:3-159, src/covariance.jl:92-446).  No mask-level fixture
of the reference is runnable here, so it is pinned by: exact 3j values
(sympy), the NaMaster golden diagonals held by the reference's tests, mpmath high-precision
recurrences at large l, and the analytic identities of SURVEY.md section 8c.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, parity_error


def test_families_match_exact_3j(oracle):
    g = np.load(os.path.join(GOLDEN, "w3j_exact.npz"))
    worst = 0.0
    for fam, l1, l2, off, n in g["index"]:
        m2, m3 = ((0, 0), (-2, 2))[fam]
        exact = g["values"][off:off + n]
        for ld in (False, True):
            nmin, f = oracle.w3j_family(int(l1), int(l2), m2, m3, ld=ld)
            assert nmin == abs(l1 - l2) and f.size == n
            worst = max(worst, np.max(np.abs(f - exact)))
            # argument order (l1,l2) vs (l2,l1): (j l1 l2; 0 m2 m3) = (-1)^(j+l1+l2) (j l2 l1; 0 m3 m2)
    assert worst < 5e-16, worst


def test_family_normalisation_and_sign(oracle):
    for (l1, l2) in [(5, 9), (100, 100), (300, 767), (2, 6143), (3000, 3071), (6100, 6143)]:
        for (m2, m3) in [(0, 0), (-2, 2)]:
            nmin, f = oracle.w3j_family(l1, l2, m2, m3)
            j = nmin + np.arange(f.size)
            assert abs(np.sum((2 * j + 1) * f * f) - 1.0) < 1e-13
            assert np.sign(f[-1]) == (-1) ** (l1 - l2)          # sgn f(jmax) = (-1)^(j2-j3-m1), m1 = 0
            if (m2, m3) == (0, 0):
                assert np.all(f[1::2] == 0.0)                  # l1+l2+j odd => exactly zero


def test_double_vs_long_double_family(oracle):
    worst = 0.0
    for (l1, l2) in [(300, 767), (767, 767), (3000, 3071), (6100, 6143), (2, 6143), (1500, 4000)]:
        for (m2, m3) in [(0, 0), (-2, 2)]:
            _, a = oracle.w3j_family(l1, l2, m2, m3)
            _, b = oracle.w3j_family(l1, l2, m2, m3, ld=True)
            worst = max(worst, np.max(np.abs(a - b)))
    assert worst < 2e-15, worst


def test_high_l_family_against_mpmath(oracle):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    for (l1, l2, m2, m3) in [(700, 767, -2, 2), (3000, 3071, -2, 2), (3000, 3071, 0, 0), (40, 6143, -2, 2)]:
        d, s = l2 - l1, l1 + l2 + 1
        a = lambda j: mp.sqrt((mp.mpf(j) ** 2 - d * d) * (mp.mpf(s) ** 2 - mp.mpf(j) ** 2))
        n = 2 * min(l1, l2) + 1
        f = [mp.mpf(1)]
        fm = mp.mpf(0)
        for t in range(n - 1):
            j = d + t
            fn = -((2 * j + 1) * (m3 - m2) * f[-1] + a(j) * fm) / a(j + 1)
            fm = f[-1]
            f.append(fn)
        norm = mp.sqrt(sum((2 * (d + t) + 1) * f[t] ** 2 for t in range(n)))
        sgn = 1 if (f[-1] > 0) == ((l1 - l2) % 2 == 0) else -1
        ref = np.array([float(sgn * v / norm) for v in f])
        _, got = oracle.w3j_family(l1, l2, m2, m3)
        assert np.max(np.abs(got - ref)) < 2e-15


def test_namaster_golden_diagonals(oracle):
    """diag(mcm) from the oracle reproduces the three NaMaster goldens the reference's tests
    hold (test/test_mcm.jl:12-50) given the even-l mask spectrum recovered from them
    (tests/golden/make_golden.py): 2295 golden values, 384 fitted numbers."""
    g = np.load(os.path.join(GOLDEN, "namaster_diag.npz"))
    V = np.zeros(768)
    V[0::2] = g["V_even"]            # odd l3 never reach the diagonal (l1+l2+l3 must be even)
    ells = np.arange(2, 767)
    for kind, key in ((0, "tt"), (2, "ee"), (1, "te")):
        M = oracle.mcm(kind, 0, 767, V)
        d = np.diag(M)[ells]
        assert np.max(np.abs(d / g[key] - 1.0)) < 2e-13, key
    # M-- on the diagonal takes odd l3 only => exactly zero for this V
    assert np.all(np.diag(oracle.mcm(3, 0, 767, V)) == 0.0)


def test_full_sky_mask_gives_identity(oracle):
    lmax = 64
    V = np.zeros(lmax + 1)
    V[0] = 4 * np.pi
    I = np.eye(lmax + 1)
    assert np.max(np.abs(oracle.mcm(0, 0, lmax, V) - I)) < 1e-13
    for kind in (1, 2):
        M = oracle.mcm(kind, 0, lmax, V)
        assert np.max(np.abs(M[2:, 2:] - I[2:, 2:])) < 1e-13
    assert np.max(np.abs(oracle.mcm(3, 0, lmax, V)[2:, 2:])) < 1e-13


def test_completeness_relations(oracle):
    lmax = 48
    V = np.ones(2 * lmax + 1)                       # nV >= 2 lmax + 1: no truncation
    l2 = np.arange(lmax + 1)
    expect = np.broadcast_to((2 * l2 + 1) / (4 * np.pi), (lmax + 1, lmax + 1))
    assert np.max(np.abs(oracle.mcm(0, 0, lmax, V) / expect - 1)) < 1e-12
    S = oracle.mcm(2, 0, lmax, V) + oracle.mcm(3, 0, lmax, V)
    assert np.max(np.abs(S[2:, 2:] / expect[2:, 2:] - 1)) < 1e-12


def test_lmin_crops_rows_only(oracle):
    rng = np.random.default_rng(3)
    V = rng.normal(size=80)
    for kind in range(4):
        full = oracle.mcm(kind, 0, 79, V)
        crop = oracle.mcm(kind, 7, 79, V)
        assert np.array_equal(full[7:, 7:], crop)


def test_covariance_reduces_to_mcm(oracle):
    """unit spectra, zero ratios, W1 = V, W2 = 0 => C = Xi(V)."""
    lmax = 60
    rng = np.random.default_rng(5)
    V = rng.normal(size=lmax + 1)
    one, zero = np.ones(lmax + 1), np.zeros(lmax + 1)
    scale = 2 * np.arange(lmax + 1) + 1.0
    C = oracle.cov("TTTT", 0, lmax, [one] * 4, [zero] * 4, [V] + [zero] * 7)
    assert np.max(np.abs(C - oracle.mcm(0, 0, lmax, V) / scale)) < 1e-15
    C = oracle.cov("EEEE", 0, lmax, [one] * 4, [zero] * 4, [V] + [zero] * 7)
    assert np.max(np.abs(C - oracle.mcm(2, 0, lmax, V) / scale)) < 1e-15
    C = oracle.cov("TETE", 0, lmax, [one, one, zero, zero], [zero] * 2, [V] + [zero] * 4)
    assert np.max(np.abs(C - oracle.mcm(1, 0, lmax, V) / scale)) < 1e-15


def test_covariance_symmetric_and_row_sampling(oracle, ps):
    from powerspectra_jl_b200 import synthetic as syn
    lmax = 96
    ws, sp, rt = syn.covariance_inputs(lmax)
    i, j, p, q = ws.field_names
    W = [ps.window_function_W(ws, *k).parent for k in [
        (ps.covariance.NULL, ps.covariance.NULL, i, p, "TT", j, q, "PP"),
        (ps.covariance.NULL, ps.covariance.NULL, i, q, "TP", j, p, "PT"),
        (ps.covariance.NULL, "PP", i, p, "TT", j, q, "PP"),
        (ps.covariance.NULL, "TT", j, q, "PP", i, p, "TT"),
        ("TT", "PP", i, p, "TT", j, q, "PP")]]
    S = [sp["TT", i, p].parent, sp["EE", j, q].parent, sp["TE", i, q].parent, sp["TE", j, p].parent]
    R = [rt["TT", i, p].parent, rt["EE", j, q].parent]
    C, terms = oracle.cov("TETE", 0, lmax, S, R, W, return_terms=True)
    assert np.array_equal(C, C.T)
    assert terms == 2 * sum((2 * l + 1) * (lmax - l + 1) for l in range(lmax + 1))
    Cs, ts = oracle.cov("TETE", 0, lmax, S, R, W, row0=3, rstep=8, return_terms=True)
    rows = np.arange(3, lmax + 1, 8)
    for r in rows:
        assert np.array_equal(Cs[r, r:], C[r, r:])
    assert ts == 2 * sum((2 * l + 1) * (lmax - l + 1) for l in rows)
    # long-double instantiation agrees with the Float64 one far inside the parity budget
    Cl = oracle.cov("TETE", 0, lmax, S, R, W, ld=True)
    assert parity_error(C, Cl) < 1e-11
