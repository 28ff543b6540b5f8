"""Mode-coupling matrix and covariance entries at BASELINE's full sizes (lmax 6143, 12287) against multiprecision known
answers.

tests/golden/mcm_entries_mp.npz (made by tests/golden/make_golden_highl.py) holds, for 96 + 48 (l1, l2) pairs -- near and
far from the diagonal, lowest spin-2 rows, last rows -- the four sums Xi_TT / Xi_TE / Xi_EE / Xi_EB of
/root/reference/src/modecoupling.jl:3-66 evaluated with mpmath at 50 digits by a plain forward recurrence: no code of
oracle/ and none of the CUDA path is involved.  The CPU tests hold the oracle (both instantiations) to them, the GPU
tests the library; both triangles M[l1,l2] = (2 l2 + 1) Xi, M[l2,l1] = (2 l1 + 1) Xi (:90-91).

tests/golden/cov_entries_mp.npz: the same for 32 entries of each of the seven covariance blocks at lmax 6143
(src/covariance.jl:92-446) over the inputs of tests/highl_inputs.py.

Criterion: the north-star 1e-10 relative on every entry whose l3 sum does not cancel by more than 1e3, and the
condition-aware bound 1e-10 |ref| + 1e-13 S_abs of tests/conftest.py on all of them.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, RTOL, TAU

KIND_COL = {0: 0, 1: 1, 2: 2, 3: 3}           # psb200_mcm kind -> column of xi / sabs


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "mcm_entries_mp.npz"))


def _check(get, gold, lmax, kinds, label):
    """get(kind) -> N x N matrix (or a callable (i, j) -> value); every golden entry, both triangles."""
    P, XI, SA = gold[f"pairs_{lmax}"], gold[f"xi_{lmax}"], gold[f"sabs_{lmax}"]
    worst_strict, worst_bound, nstrict = 0.0, 0.0, 0
    for kind in kinds:
        M = get(kind)
        c = KIND_COL[kind]
        for (l1, l2), xi, sa in zip(P, XI[:, c], SA[:, c]):
            for (i, j) in ((l1, l2), (l2, l1)):
                ref, cond = (2 * j + 1) * xi, (2 * j + 1) * sa
                err = abs(M[i, j] - ref)
                worst_bound = max(worst_bound, err / (RTOL * abs(ref) + TAU * cond))
                if cond <= 1e3 * abs(ref):
                    worst_strict = max(worst_strict, err / abs(ref))
                    nstrict += 1
    assert worst_bound <= 1.0, (label, lmax, worst_bound)
    assert worst_strict < RTOL, (label, lmax, worst_strict)
    assert nstrict > len(P)                    # the strict criterion applies to most entries
    return worst_strict, worst_bound


def test_golden_file_is_self_consistent(gold):
    """Completeness of the two families over a flat window would need len(V) > 2 lmax; what the file itself allows:
    |Xi| <= S_abs, Xi_TT's terms are non-negative only if V is (it is not: cross-mask spectrum), Xi_EE + Xi_EB and
    Xi_TT share the bound sum (2 l3 + 1) f^2 |V| / 4 pi <= max|V| / 4 pi."""
    for lmax in (6143, 12287):
        P, XI, SA, V = gold[f"pairs_{lmax}"], gold[f"xi_{lmax}"], gold[f"sabs_{lmax}"], gold[f"V_{lmax}"]
        assert V.size == lmax + 1 and P.shape[1] == 2 and XI.shape == SA.shape == (len(P), 4)
        assert (P[:, 0] <= P[:, 1]).all() and P.min() >= 2 and P.max() == lmax
        assert (np.abs(XI) <= SA * (1 + 1e-12)).all()
        vmax = np.abs(V).max() / (4 * np.pi)
        assert (SA[:, 0] <= vmax * (1 + 1e-12)).all() and (SA[:, 2] + SA[:, 3] <= vmax * (1 + 1e-12)).all()


@pytest.mark.parametrize("ld", [False, True])
def test_oracle_entries_lmax6143(oracle, gold, ld):
    lmax = 6143
    rows = np.unique(gold[f"pairs_{lmax}"][:, 0])
    V = gold[f"V_{lmax}"]
    _check(lambda kind: oracle.mcm(kind, 0, lmax, V, ld=ld, rows=rows), gold, lmax, (0, 1, 2, 3), f"oracle ld={ld}")


def test_oracle_entries_lmax12287(oracle, gold):
    lmax = 12287
    rows = np.unique(gold[f"pairs_{lmax}"][:, 0])
    V = gold[f"V_{lmax}"]
    _check(lambda kind: oracle.mcm(kind, 0, lmax, V, ld=True, rows=rows), gold, lmax, (0, 3), "oracle ld")


@pytest.mark.gpu
def test_gpu_entries_lmax6143(ps, gold):
    lmax = 6143
    V = ps.SpectralVector(gold[f"V_{lmax}"])
    ee_bb = ps.mcm("EE_BB", V)
    mats = {0: ps.mcm("TT", V).parent, 1: ps.mcm("TE", V).parent,
            2: ee_bb.getblock(0, 0).parent, 3: ee_bb.getblock(0, 1).parent}
    ws, wb = _check(mats.__getitem__, gold, lmax, (0, 1, 2, 3), "gpu")
    print(f"gpu vs 50-digit entries, lmax {lmax}: strict max {ws:.2e}, err/bound max {wb:.3f}")


@pytest.mark.gpu
@pytest.mark.parametrize("kinds", [(0,), (2, 3)])
def test_gpu_entries_lmax12287(ps, gold, kinds):
    lmax = 12287
    V = ps.SpectralVector(gold[f"V_{lmax}"])
    if kinds == (0,):
        mats = {0: ps.mcm("TT", V).parent}
    else:
        ee_bb = ps.mcm("EE_BB", V)
        mats = {2: ee_bb.getblock(0, 0).parent, 3: ee_bb.getblock(0, 1).parent}
    ws, wb = _check(mats.__getitem__, gold, lmax, kinds, "gpu")
    print(f"gpu vs 50-digit entries, lmax {lmax}, kinds {kinds}: strict max {ws:.2e}, err/bound max {wb:.3f}")


# ---- covariance blocks of the benchmark step at lmax 6143 (cov_entries_mp.npz) ----------------------------------

@pytest.fixture(scope="module")
def cov_gold():
    import highl_inputs
    g = np.load(os.path.join(GOLDEN, "cov_entries_mp.npz"))
    inputs = highl_inputs.cov_inputs(int(g["lmax"]))
    assert highl_inputs.digest(inputs) == str(g["inputs_sha256"]), "tests/highl_inputs.py no longer makes the vectors of the fixture"
    return g, inputs


BLOCKS = ("TTTT", "EEEE", "TETE", "TTTE", "TEEE_planck", "TEEE", "TTEE")


def _check_cov(get, g, label):
    worst_strict, worst_bound, nstrict = 0.0, 0.0, 0
    for block in BLOCKS:
        Cm = get(block)
        for (l1, l2), ref, cond in zip(g["pairs"], g[block], g[block + "_sabs"]):
            assert Cm[l1, l2] == Cm[l2, l1]                       # symmetric by copy (src/covariance.jl:119)
            err = abs(Cm[l1, l2] - ref)
            worst_bound = max(worst_bound, err / (RTOL * abs(ref) + TAU * cond))
            if cond <= 1e3 * abs(ref):
                worst_strict = max(worst_strict, err / abs(ref))
                nstrict += 1
    assert worst_bound <= 1.0, (label, worst_bound)
    assert worst_strict < RTOL, (label, worst_strict)
    assert nstrict >= 2 * len(g["pairs"])
    return worst_strict, worst_bound


@pytest.mark.parametrize("ld", [False, True])
def test_oracle_cov_entries_lmax6143(oracle, cov_gold, ld):
    g, inputs = cov_gold
    lmax = int(g["lmax"])
    rows = np.unique(g["pairs"][:, 0])
    ws, wb = _check_cov(lambda b: oracle.cov(b, 0, lmax, *inputs[b], ld=ld, rows=rows), g, f"oracle ld={ld}")
    print(f"oracle ld={ld} vs 50-digit covariance entries: strict max {ws:.2e}, err/bound max {wb:.4f}")


@pytest.mark.gpu
def test_gpu_cov_entries_lmax6143(ps, cov_gold):
    g, inputs = cov_gold
    lmax = int(g["lmax"])
    loops = {"TTTT": ps.loop_covTTTT, "EEEE": ps.loop_covEEEE, "TETE": ps.loop_covTETE, "TTTE": ps.loop_covTTTE,
             "TEEE_planck": ps.loop_covTEEE_planck, "TEEE": ps.loop_covTEEE, "TTEE": ps.loop_covTTEE}

    def get(block):
        sp, rt, W = inputs[block]
        Cm = ps.spectralzeros(range(0, lmax + 1), range(0, lmax + 1))
        loops[block](Cm, *[ps.SpectralVector(x) for x in sp], *[ps.SpectralVector(x) for x in rt], *[ps.SpectralVector(x) for x in W])
        return Cm.parent
    ws, wb = _check_cov(get, g, "gpu")
    print(f"gpu vs 50-digit covariance entries, lmax {lmax}: strict max {ws:.2e}, err/bound max {wb:.3f}")


# ---- QuickPol Xi at lmax 6143, band +-128 (quickpol_entries_mp.npz) ----------------------------------------------

@pytest.fixture(scope="module")
def qp_gold():
    import highl_inputs
    g = np.load(os.path.join(GOLDEN, "quickpol_entries_mp.npz"))
    W = highl_inputs.quickpol_window(int(g["lmax"]))
    assert highl_inputs.digest({"W": ([W],)}) == str(g["window_sha256"])
    return g, W


def _check_qp(get, g, label):
    """get(case index) -> callable (l'', l) -> Xi value; every golden entry of every spin case."""
    worst_strict, worst_bound, nstrict = 0.0, 0.0, 0
    for c in range(len(g["cases"])):
        Xi = get(c)
        for (lpp, l), ref, cond in zip(g["entries"], g["xi"][c], g["sabs"][c]):
            err = abs(Xi(int(lpp), int(l)) - ref)
            if cond == 0.0:                                       # projection larger than the angular momentum: exactly 0
                assert err == 0.0
                continue
            worst_bound = max(worst_bound, err / (RTOL * abs(ref) + TAU * cond))
            if cond <= 1e3 * abs(ref):
                worst_strict = max(worst_strict, err / abs(ref))
                nstrict += 1
    assert worst_bound <= 1.0, (label, worst_bound)
    assert worst_strict < RTOL, (label, worst_strict)
    assert nstrict >= 2 * len(g["entries"])
    return worst_strict, worst_bound


@pytest.mark.parametrize("ld", [False, True])
def test_oracle_quickpol_entries(oracle, qp_gold, ld):
    """The oracle's general-spin family routine at high l: Xi entries rebuilt from its families (Xisum, src/beam.jl:17-28)."""
    import math
    g, W = qp_gold

    def get(c):
        nu1, nu2, s1, s2 = (int(x) for x in g["cases"][c])
        sgn = -1.0 if (s1 + s2 + nu1 + nu2) % 2 else 1.0

        def Xi(lpp, l):
            if abs(s1) > l or abs(s2) > l or abs(nu1) > lpp or abs(nu2) > lpp:
                return 0.0
            a1, f1 = oracle.w3j_family(l, lpp, -s1, -nu1, ld=ld)
            a2, f2 = oracle.w3j_family(l, lpp, -s2, -nu2, ld=ld)
            lo, hi = max(a1, a2), min(a1 + f1.size - 1, a2 + f2.size - 1, W.size - 1)
            return sgn * math.fsum(W[lo:hi + 1] * f1[lo - a1:hi - a1 + 1] * f2[lo - a2:hi - a2 + 1])
        return Xi
    ws, wb = _check_qp(get, g, f"oracle ld={ld}")
    print(f"oracle families ld={ld} vs 60-digit QuickPol entries: strict max {ws:.2e}, err/bound max {wb:.4f}")


@pytest.mark.gpu
def test_gpu_quickpol_entries(ps, qp_gold):
    g, W = qp_gold
    lmax, band = int(g["lmax"]), int(g["band"])

    def get(c):
        case = tuple(int(x) for x in g["cases"][c])               # (nu1, nu2, s1, s2)
        Xi = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, band, band), *case, ps.SpectralVector(W))
        return lambda lpp, l: Xi[lpp, l]
    ws, wb = _check_qp(get, g, "gpu")
    print(f"gpu vs 60-digit QuickPol entries, lmax {lmax}, band +-{band}: strict max {ws:.2e}, err/bound max {wb:.3f}")
