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


def exact_mcm_from_sympy(kind, lmax, V):
    """Brute-force M[l1,l2] from the exact (sympy) 3j values of tests/golden/w3j_exact.npz, written
    straight from the definitions in docs/src/spectra.md / src/modecoupling.jl:3-66 -- shares no
    code with the oracle's loops."""
    g = np.load(os.path.join(GOLDEN, "w3j_exact.npz"))
    fam = {}
    for f, l1, l2, off, n in g["index"]:
        fam[(int(f), int(l1), int(l2))] = g["values"][off:off + n]
    M = np.zeros((lmax + 1, lmax + 1))
    for l1 in range(lmax + 1):
        for l2 in range(lmax + 1):
            a, b = min(l1, l2), max(l1, l2)
            f0 = fam[(0, a, b)]
            f2 = fam.get((1, a, b))
            if kind != 0 and f2 is None:
                continue                           # l < 2 for a spin-2 kind: don't-care region
            js = np.arange(b - a, a + b + 1)
            sel = js < V.size
            par = (l1 + l2 + js) % 2
            if kind == 0:
                w = f0 ** 2
            elif kind == 1:
                w = f0 * f2
                sel &= par == 0
            else:
                w = f2 ** 2
                sel &= par == (0 if kind == 2 else 1)
            M[l1, l2] = (2 * l2 + 1) / (4 * np.pi) * np.sum((2 * js[sel] + 1) * w[sel] * V[js[sel]])
    return M


@pytest.mark.parametrize("nV", [25, 49, 12])
def test_oracle_mcm_against_exact_3j_bruteforce(oracle, nV):
    lmax = 24
    rng = np.random.default_rng(100 + nV)
    V = rng.normal(size=nV)
    for kind in range(4):
        E = exact_mcm_from_sympy(kind, lmax, V)
        lo = 2 if kind else 0
        for ld in (False, True):
            M = oracle.mcm(kind, 0, lmax, V, ld=ld)
            assert np.max(np.abs(M[lo:, lo:] - E[lo:, lo:])) < 5e-14 * max(1.0, np.abs(E).max()), (kind, ld)


def exact_xi_from_sympy(lmax, W, which):
    """Xi[l1,l2] = (1/4pi) sum (2 l3+1) w(l3) W(l3) from exact 3j; which: "00" f00^2 all parities,
    "22" f22^2 even parity, "02" f00 f22 even parity (src/modecoupling.jl:3-66)."""
    kind = {"00": 0, "22": 2, "02": 1}[which]
    M = exact_mcm_from_sympy(kind, lmax, np.asarray(W))
    return M / (2 * np.arange(lmax + 1) + 1.0)[None, :]


def exact_cov_from_sympy(block, lmax, sp, rt, W):
    """Second, independent transcription of the seven block formulas of /root/reference/src/covariance.jl
    (numpy, vectorised over (l1, l2)), on top of the exact-3j Xi."""
    l = np.arange(lmax + 1)
    A = lambda k: np.asarray(sp[k])[l][:, None]      # value at l1
    B = lambda k: np.asarray(sp[k])[l][None, :]      # value at l2
    Ra = lambda k: np.asarray(rt[k])[l][:, None]
    Rb = lambda k: np.asarray(rt[k])[l][None, :]
    X = lambda k, which: exact_xi_from_sympy(lmax, W[k], which)
    if block in ("TTTT", "EEEE"):
        w = "00" if block == "TTTT" else "22"
        return (np.sqrt(A(0) * B(0) * A(1) * B(1)) * X(0, w) + np.sqrt(A(2) * B(2) * A(3) * B(3)) * X(1, w)
                + np.sqrt(A(0) * B(0)) * X(2, w) * Ra(1) * Rb(1) + np.sqrt(A(1) * B(1)) * X(3, w) * Ra(0) * Rb(0)
                + np.sqrt(A(2) * B(2)) * X(4, w) * Ra(3) * Rb(3) + np.sqrt(A(3) * B(3)) * X(5, w) * Ra(2) * Rb(2)
                + X(6, w) * Ra(0) * Ra(1) * Rb(0) * Rb(1) + X(7, w) * Ra(2) * Ra(3) * Rb(2) * Rb(3))
    if block == "TTTE":
        return (np.sqrt(A(0) * B(0)) * (A(3) + B(3)) * X(0, "00") + np.sqrt(A(1) * B(1)) * (A(2) + B(2)) * X(1, "00")
                + (A(3) + B(3)) * X(2, "00") * Ra(0) * Rb(0) + (A(2) + B(2)) * X(3, "00") * Ra(1) * Rb(1)) / 2
    if block == "TETE":
        return (np.sqrt(A(0) * B(0) * A(1) * B(1)) * X(0, "02") + 0.5 * (A(2) * B(3) + A(3) * B(2)) * X(1, "00")
                + np.sqrt(A(0) * B(0)) * X(2, "02") * Ra(1) * Rb(1) + np.sqrt(A(1) * B(1)) * X(3, "02") * Ra(0) * Rb(0)
                + X(4, "02") * Ra(0) * Rb(0) * Ra(1) * Rb(1))
    if block in ("TEEE_planck", "TEEE"):
        w = "22" if block == "TEEE_planck" else "02"
        return (np.sqrt(A(0) * B(0)) * (A(2) + B(2)) * X(0, w) + np.sqrt(A(1) * B(1)) * (A(3) + B(3)) * X(1, w)
                + (A(2) + B(2)) * X(2, w) * Ra(0) * Rb(0) + (A(3) + B(3)) * X(3, w) * Ra(1) * Rb(1)) / 2
    if block == "TTEE":
        return ((A(0) * B(2) + A(2) * B(0)) * X(0, "00") + (A(1) * B(3) + A(3) * B(1)) * X(1, "00")) / 2
    raise KeyError(block)


COV_SHAPES = {"TTTT": (4, 4, 8), "EEEE": (4, 4, 8), "TTTE": (4, 2, 4), "TETE": (4, 2, 5),
              "TEEE_planck": (4, 2, 4), "TEEE": (4, 2, 4), "TTEE": (4, 0, 2)}


def cov_case_small(block, lmax, seed=0):
    rng = np.random.default_rng(seed + len(block))
    nsp, nrt, nw = COV_SHAPES[block]
    sp = [rng.uniform(0.5, 2.0, size=lmax + 1) for _ in range(nsp)]
    if block in ("TTTE", "TEEE", "TEEE_planck"):
        sp[2], sp[3] = sp[2] - 1.2, sp[3] - 1.2          # TE spectra may be negative
    if block == "TETE":
        sp[2], sp[3] = sp[2] - 1.2, sp[3] - 1.2
    if block == "TTEE":
        sp = [x - 1.2 for x in sp]
    rt = [rng.uniform(0.5, 1.5, size=lmax + 1) for _ in range(nrt)]
    W = [rng.normal(size=lmax + 1) for _ in range(nw)]
    return sp, rt, W


@pytest.mark.parametrize("block", list(COV_SHAPES))
def test_oracle_cov_against_exact_3j_bruteforce(oracle, block):
    lmax = 24
    sp, rt, W = cov_case_small(block, lmax)
    E = exact_cov_from_sympy(block, lmax, sp, rt, W)
    lo = 0 if block in ("TTTT", "TTTE", "TTEE") else 2
    for ld in (False, True):
        Cm = oracle.cov(block, 0, lmax, sp, rt, W, ld=ld)
        assert np.max(np.abs(Cm[lo:, lo:] - E[lo:, lo:])) < 1e-13 * max(1.0, np.abs(E[lo:, lo:]).max()), (block, ld)


def test_even_parity_spin2_identity_exact():
    """The identity the tuned kernel uses for every even-parity-only job (psb200_pair_v2.cuh):
        (j l1 l2; 0 -2 2) = (j l1 l2; 0 0 0) * N(x) / D   for l1+l2+j even,
        x = j(j+1), a = l1(l1+1), b = l2(l2+1), N = (x-a-b)(x-a-b+2)/2 - a b,
        D = sqrt((l1-1) l1 (l1+1)(l1+2)(l2-1) l2 (l2+1)(l2+2)),
    checked on every exact (sympy) family of tests/golden/w3j_exact.npz (l up to 80)."""
    g = np.load(os.path.join(GOLDEN, "w3j_exact.npz"))
    fam = {(int(f), int(a), int(b)): g["values"][off:off + n] for f, a, b, off, n in g["index"]}
    worst, checked = 0.0, 0
    for (f, l1, l2), f22 in fam.items():
        if f != 1:
            continue
        f00 = fam[(0, l1, l2)]
        j = np.arange(l2 - l1, l1 + l2 + 1, dtype=np.float64)
        even = ((l1 + l2 + j.astype(int)) % 2) == 0
        x, a, b = j * (j + 1), l1 * (l1 + 1.0), l2 * (l2 + 1.0)
        N = 0.5 * (x - a - b) * (x - a - b + 2) - a * b
        D = np.sqrt((l1 - 1.0) * l1 * (l1 + 1) * (l1 + 2) * (l2 - 1.0) * l2 * (l2 + 1) * (l2 + 2))
        worst = max(worst, np.max(np.abs(f22[even] - f00[even] * N[even] / D)))
        checked += int(even.sum())
    assert checked > 3000 and worst < 2e-15, (checked, worst)


def test_closed_form_products():
    """The closed forms the default kernel evaluates instead of any recurrence (psb200_pair_v3.cuh), against every exact
    (sympy) family of tests/golden/w3j_exact.npz (l up to 80).  g(n) = binom(2n,n)/4^n, d = l2-l1, L = 2 l1+1, t = j-d, m = j+d:
        even parity:  f00(j)^2 = PT[t] PV[m],  PT[t] = g(t/2) g(l1-t/2),  PV[m] = g(m/2) / (g(m/2+l1) (m+L))
                      2 D f00 f22 = PT PV (u^2 - (2ab+1)),   4 D^2 f22^2 = PT PV (u^2 - (2ab+1))^2,   u = x - (a+b-1)
        odd parity:   4 D^2 f22^2 = (x-a-b+2)^2 QT[t] QV[m],  QT[t] = t (L-t) PT[t-1],  QV[m] = m (m+L) PV[m-1]."""
    from math import comb
    gm = [comb(2 * n, n) / 4.0 ** n for n in range(400)]
    g = np.load(os.path.join(GOLDEN, "w3j_exact.npz"))
    fam = {(int(f), int(a), int(b)): g["values"][off:off + n] for f, a, b, off, n in g["index"]}
    worst = {"f00^2": 0.0, "f00 f22": 0.0, "f22^2 even": 0.0, "f22^2 odd": 0.0}
    checked = 0
    for (f, l1, l2), f22 in fam.items():
        if f != 1 or l1 < 2:
            continue
        f00 = fam[(0, l1, l2)]
        d, L = l2 - l1, 2 * l1 + 1
        a, b = l1 * (l1 + 1.0), l2 * (l2 + 1.0)
        D2 = (l1 - 1.0) * l1 * (l1 + 1) * (l1 + 2) * (l2 - 1.0) * l2 * (l2 + 1) * (l2 + 2)
        for i, j in enumerate(range(d, l1 + l2 + 1)):
            t, m, x = j - d, j + d, j * (j + 1.0)
            if t % 2 == 0:
                PT = gm[t // 2] * gm[l1 - t // 2]
                PV = gm[m // 2] / (gm[m // 2 + l1] * (m + L))
                u = x - (a + b - 1.0)
                nn = u * u - (2.0 * a * b + 1.0)
                scale = max(abs(f00[i]), abs(f22[i]), 1e-300) ** 2
                worst["f00^2"] = max(worst["f00^2"], abs(PT * PV - f00[i] ** 2) / f00[i] ** 2)
                worst["f00 f22"] = max(worst["f00 f22"], abs(PT * PV * nn / (2 * np.sqrt(D2)) - f00[i] * f22[i]) / scale)
                worst["f22^2 even"] = max(worst["f22^2 even"], abs(PT * PV * nn * nn / (4 * D2) - f22[i] ** 2) / scale)
            else:
                assert abs(f00[i]) < 1e-300
                PT = gm[(t - 1) // 2] * gm[l1 - (t - 1) // 2]
                PV = gm[(m - 1) // 2] / (gm[(m - 1) // 2 + l1] * (m - 1 + L))
                uo = x - a - b + 2.0
                h = (t * (L - t) * PT) * (m * (m + L) * PV) * uo * uo / (4 * D2)
                worst["f22^2 odd"] = max(worst["f22^2 odd"], abs(h - f22[i] ** 2) / max(f22[i] ** 2, 1e-300))
            checked += 1
    assert checked > 5000 and max(worst.values()) < 5e-14, (checked, worst)
