"""QuickPol Xi matrix (/root/reference/src/beam.jl:72-101).

CPU part: the oracle's general-spin families and its quickpolXi! restatement against exact 3j symbols
(sympy fixtures, tests/golden/w3j_general_exact.npz); the arithmetic of the CUDA pair function --
compiled for the host by tests/hostcheck, test infrastructure only -- against the long-double oracle,
including the rescaling path; host containers and argument checking.
GPU part (-m gpu): the CUDA kernel through the C ABI against the long-double oracle and the exact fixtures.
Tolerance: |test - ref| <= 1e-10 |ref| + 1e-13 S_abs (S_abs = sum of |terms|, oracle abs_mode): the strict
north-star 1e-10 wherever the l' sum does not cancel."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = [(0, 0, 0, 0), (2, 2, 2, 2), (-2, 2, 0, 2), (0, 2, 1, -1), (2, -2, 2, -2), (0, 0, 2, 2), (2, 0, 3, 1),
         (-2, -2, 4, -3), (2, 2, 10, -7), (-2, 2, 20, 20)]


def _scan_spectrum(n, seed=7):
    """Synthetic scan-pattern spectrum: red, sign-changing (cross-spectra of different spins are not positive)."""
    rng = np.random.default_rng(seed)
    l = np.arange(n)
    return rng.normal(size=n) / (1.0 + l / 40.0) ** 2 + 1.0 / (1.0 + l) ** 1.5


def _bound_ratio(test, ref, sabs, rtol=1e-10, tau=1e-13):
    return float(np.max(np.abs(test - ref) / (rtol * np.abs(ref) + tau * sabs + 1e-300)))


def _oracle_ref(oracle, case, lmax, W, bl, bh):
    nu1, nu2, s1, s2 = case
    ref = oracle.quickpol_xi(nu1, nu2, s1, s2, lmax, W, bl, bh, ld=True, dense=False)
    with oracle.abs_mode():
        sabs = oracle.quickpol_xi(nu1, nu2, s1, s2, lmax, W, bl, bh, ld=True, dense=False)
    return ref, sabs


# ----------------------------------------------------------------------------------------------------
# CPU: oracle pinned to exact symbols
# ----------------------------------------------------------------------------------------------------
def test_oracle_general_families_match_exact_3j(oracle):
    g = np.load(os.path.join(GOLDEN, "w3j_general_exact.npz"))
    worst = 0.0
    for l, lpp, m2, m3, lo, off, n in g["index"]:
        exact = g["values"][off:off + n]
        for ld in (False, True):
            nmin, f = oracle.w3j_family(int(l), int(lpp), int(m2), int(m3), ld=ld)
            assert nmin == lo and f.size == n
            worst = max(worst, float(np.max(np.abs(f - exact))))
    assert worst < 5e-16, worst


def test_oracle_quickpol_matches_exact_xi(oracle):
    g = np.load(os.path.join(GOLDEN, "w3j_general_exact.npz"))
    lmax, W = int(g["xi_lmax"]), g["xi_W"]
    for case, exact in zip(g["xi_cases"], g["xi"]):
        nu1, nu2, s1, s2 = (int(c) for c in case)
        for ld in (False, True):
            xi = oracle.quickpol_xi(nu1, nu2, s1, s2, lmax, W, lmax, lmax, ld=ld)
            assert np.max(np.abs(xi - exact)) < 2e-15
        # a narrower, asymmetric band stores the same numbers
        xb = oracle.quickpol_xi(nu1, nu2, s1, s2, lmax, W, 3, 1, ld=True)
        i, j = np.indices(xb.shape)
        band = (i - j <= 3) & (j - i <= 1)
        assert np.max(np.abs(xb - np.where(band, exact, 0.0))) < 2e-15


def test_oracle_quickpol_reduces_to_mcm_kernels(oracle):
    """nu = s = 0 is the (0,0,0) family squared: Xi[l'', l] = sum W f00^2, so with W = (2l'+1) V / 4pi
    the matrix is Xi_TT of src/modecoupling.jl:3-13 and (2l+1) Xi = M_TT[l'', l]."""
    lmax = 40
    V = _scan_spectrum(lmax + 1, seed=3)
    W = (2 * np.arange(lmax + 1) + 1) * V / (4 * np.pi)
    xi = oracle.quickpol_xi(0, 0, 0, 0, lmax, W, lmax, lmax, ld=True)
    M = oracle.mcm("M00", 0, lmax, V, ld=True)
    l = np.arange(lmax + 1)
    assert np.max(np.abs(xi[2:, 2:] * (2 * l[None, 2:] + 1) - M[2:, 2:])) < 1e-14
    # (s, nu) = (2, -2) twice is the (0,-2,2) family squared, every parity: M++ + M--
    xi22 = oracle.quickpol_xi(-2, -2, 2, 2, lmax, W, lmax, lmax, ld=True)
    Mpp = oracle.mcm("Mpp", 0, lmax, V, ld=True)
    Mmm = oracle.mcm("Mmm", 0, lmax, V, ld=True)
    assert np.max(np.abs(xi22[2:, 2:] * (2 * l[None, 2:] + 1) - (Mpp + Mmm)[2:, 2:])) < 1e-14


# ----------------------------------------------------------------------------------------------------
# CPU: the arithmetic of the CUDA pair function (host build, tests/hostcheck)
# ----------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hostcheck():
    import hostcheck as hc
    hc.lib()
    return hc


VARIANTS = ["tab", "simple"]          # the two instantiations of the CUDA kernel (PSB200_QP)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("case", CASES)
def test_kernel_arithmetic_on_host_vs_oracle(oracle, hostcheck, case, variant):
    lmax, bl, bh = 300, 30, 25
    W = _scan_spectrum(2 * lmax + 1)
    ref, sabs = _oracle_ref(oracle, case, lmax, W, bl, bh)
    got = hostcheck.xi_band(*case, lmax, W, bl, bh, variant=variant)
    assert _bound_ratio(got, ref, sabs) < 1.0
    assert np.count_nonzero(ref) > 10000


@pytest.mark.parametrize("variant", VARIANTS)
def test_kernel_arithmetic_on_host_exact_and_short_window(oracle, hostcheck, variant):
    g = np.load(os.path.join(GOLDEN, "w3j_general_exact.npz"))
    lmax, W = int(g["xi_lmax"]), g["xi_W"]
    for case, exact in zip(g["xi_cases"], g["xi"]):
        xb = hostcheck.xi_band(*(int(c) for c in case), lmax, W, lmax, lmax, variant=variant)
        assert np.max(np.abs(oracle.band_to_dense(xb, lmax, lmax, lmax) - exact)) < 5e-15
    # window shorter than the families: terms above lenW-1 are dropped, pairs beyond reach are zero
    lmax = 120
    for nW in (1, 2, 7, 60):
        W = _scan_spectrum(nW)
        for case in [(0, 0, 0, 0), (2, -2, 2, 1), (0, 2, 3, -1)]:
            ref, sabs = _oracle_ref(oracle, case, lmax, W, 20, 20)
            got = hostcheck.xi_band(*case, lmax, W, 20, 20, variant=variant)
            assert _bound_ratio(got, ref, sabs) < 1.0
    # bandwidths beyond the matrix size are legal in BandedMatrices (padding rows of the storage)
    W = _scan_spectrum(90)
    ref, sabs = _oracle_ref(oracle, (2, 0, 1, 2), 40, W, 70, 55)
    got = hostcheck.xi_band(2, 0, 1, 2, 40, W, 70, 55, variant=variant)
    assert _bound_ratio(got, ref, sabs) < 1.0 and np.count_nonzero(ref) > 1000


@pytest.mark.parametrize("variant", VARIANTS)
def test_kernel_arithmetic_rescaling_path(oracle, hostcheck, variant):
    """Spins close to l: the non-classical regions span > 1e200 in magnitude, so the sweeps rescale."""
    rng = np.random.default_rng(5)
    for (l, lpp, nu1, nu2, s1, s2) in [(2000, 2000, 2, 2, 1990, 1990), (2000, 1990, 2, -2, 1900, -1900),
                                       (3000, 2950, 2, 2, 2990, 10), (1000, 1000, 2, 2, 1000, 1000),
                                       (1000, 980, 0, 2, 999, -999), (3000, 3000, 2, 2, 2990, 2980)]:
        W = rng.normal(size=l + lpp + 1)
        n1, f1 = oracle.w3j_family(l, lpp, -s1, -nu1, ld=True)
        n2, f2 = oracle.w3j_family(l, lpp, -s2, -nu2, ld=True)
        a = max(n1, n2)
        j = np.arange(a, l + lpp + 1)
        t = W[j] * f1[j - n1] * f2[j - n2]
        ref = (-1.0) ** ((s1 + s2 + nu1 + nu2) % 2) * t.sum()
        got = hostcheck.pair(l, lpp, nu1, nu2, s1, s2, W, variant=variant)
        assert abs(got - ref) <= 1e-10 * abs(ref) + 1e-13 * np.abs(t).sum(), (l, lpp, s1, s2, got, ref)


# ----------------------------------------------------------------------------------------------------
# CPU: host mirror
# ----------------------------------------------------------------------------------------------------
def test_banded_container_and_quickpolW(ps):
    B = ps.BandedSpectralMatrix(6, 2, 1)
    assert B.data.shape == (4, 7) and B.shape == (7, 7)
    B[3, 4] = 5.0
    B[5, 3] = -1.0
    assert B[3, 4] == 5.0 and B[5, 3] == -1.0 and B[0, 6] == 0.0
    with pytest.raises(IndexError):
        B[0, 3] = 1.0
    D = B.todense()
    assert D[3, 4] == 5.0 and D[5, 3] == -1.0 and np.count_nonzero(D) == 2
    x = np.arange(7.0)
    assert np.allclose(B.matvec(x), D @ x)
    assert list(B.rowrange(5)) == [3, 4, 5, 6] and list(B.rowrange(2)) == [2, 3]      # specrowrange, never below 2
    # quickpolW: sum over m of a conj(b), no 1/(2l+1)  (src/beam.jl:43-56)
    rng = np.random.default_rng(0)
    lmax = 5
    n = (lmax + 1) * (lmax + 2) // 2
    a = ps.Alm(lmax, lmax, rng.normal(size=n) + 1j * rng.normal(size=n))
    b = ps.Alm(lmax, lmax, rng.normal(size=n) + 1j * rng.normal(size=n))
    W = ps.quickpolW(a, b).parent
    assert np.allclose(W, ps.alm2cl(a, b) * (2 * np.arange(lmax + 1) + 1))
    assert ps.k_u(0) == 1.0 and ps.k_u(-2) == 0.5
    with pytest.raises(ValueError):
        ps.k_u(1)


def test_quickpol_argument_checks(ps):
    L = ps.lib()
    W = np.ones(8)
    Xb = np.zeros((5, 8), order="F")
    wp, xp = W.ctypes.data_as(ps._lib.DP), Xb.ctypes.data_as(ps._lib.DP)
    assert L.psb200_quickpol_xi(0, 0, 0, 0, 7, wp, 8, 2, 2, xp, 4, 1) == 1       # ldb below the band height
    assert L.psb200_quickpol_xi(0, 0, 0, 0, -1, wp, 8, 2, 2, xp, 5, 1) == 1
    assert L.psb200_quickpol_xi(0, 0, 0, 0, 7, wp, 0, 2, 2, xp, 5, 1) == 1
    assert L.psb200_quickpol_xi(0, 0, 0, 0, 7, wp, 8, -1, 2, xp, 5, 1) == 1
    assert L.psb200_quickpol_xi(0, 0, 0, 0, 7, None, 8, 2, 2, xp, 5, 1) == 1
    if L.psb200_device_count() == 0:
        assert L.psb200_quickpol_xi(0, 0, 0, 0, 7, wp, 8, 2, 2, xp, 5, 1) == 5
        assert b"no CPU fallback" in L.psb200_last_error()
        with pytest.raises(ps.PSB200Error):
            ps.quickpolXi(ps.BandedSpectralMatrix(7, 2, 2), 0, 0, 0, 0, ps.SpectralVector(W))
        assert np.all(Xb == 0.0)
    with pytest.raises(ValueError):
        ps.quickpolXi(np.zeros((8, 8)), 0, 0, 0, 0, ps.SpectralVector(W))
    e = (ps._lib.C.c_int * 5)()
    assert L.psb200_quickpol_edges(6143, 100, 100, 4, e) == 0
    assert e[0] == 0 and e[4] == 6144 and all(e[k] < e[k + 1] for k in range(4))
    # equal-cost bands: column cost grows ~ l, so the edges follow sqrt(k/4)
    assert abs(e[2] - 6144 * np.sqrt(0.5)) < 40


# ----------------------------------------------------------------------------------------------------
# GPU: the CUDA kernel through the C ABI
# ----------------------------------------------------------------------------------------------------
@pytest.fixture(params=VARIANTS)
def qp_variant(request, monkeypatch):
    """PSB200_QP is read on every call: both instantiations of the kernel are tested in one process."""
    monkeypatch.setenv("PSB200_QP", request.param)
    return request.param


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_quickpol_vs_oracle(ps, oracle, case, qp_variant):
    lmax, bl, bh = 300, 30, 25
    W = _scan_spectrum(2 * lmax + 1)
    ref, sabs = _oracle_ref(oracle, case, lmax, W, bl, bh)
    Xi = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, bl, bh), *case, ps.SpectralVector(W))
    assert _bound_ratio(Xi.data, ref, sabs) < 1.0
    # strict north-star form on the entries whose l' sum does not cancel
    sel = (np.abs(ref) > 0) & (sabs <= 1e3 * np.abs(ref))
    assert sel.sum() > 5000
    assert np.max(np.abs(Xi.data[sel] - ref[sel]) / np.abs(ref[sel])) < 1e-10


@pytest.mark.gpu
def test_gpu_quickpol_exact_3j_and_untouched_entries(ps, oracle, qp_variant):
    g = np.load(os.path.join(GOLDEN, "w3j_general_exact.npz"))
    lmax, W = int(g["xi_lmax"]), g["xi_W"]
    for case, exact in zip(g["xi_cases"], g["xi"]):
        case = tuple(int(c) for c in case)
        Xi = ps.BandedSpectralMatrix(lmax, lmax, lmax)
        Xi.data[:] = 7.0                              # rows / columns < 2 are not visited by the reference loop
        ps.quickpolXi(Xi, *case, ps.SpectralVector(W))
        D = Xi.todense()
        assert np.max(np.abs(D[2:, 2:] - exact[2:, 2:])) < 5e-15
        sgn = -1.0 if sum(case) % 2 else 1.0          # ... and only take the final sign (src/beam.jl:98-99)
        assert np.all(D[:2, :] == 7.0 * sgn) and np.all(D[:, :2] == 7.0 * sgn)


@pytest.mark.gpu
def test_gpu_quickpol_short_window_and_bands(ps, oracle, qp_variant):
    lmax = 200
    for nW, bl, bh in [(1, 10, 10), (2, 0, 0), (7, 5, 40), (60, 40, 5), (401, 200, 200), (1000, 3, 0)]:
        W = _scan_spectrum(nW)
        for case in [(0, 0, 0, 0), (2, -2, 2, 1), (0, 2, 3, -1)]:
            ref, sabs = _oracle_ref(oracle, case, lmax, W, bl, bh)
            Xi = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, bl, bh), *case, ps.SpectralVector(W))
            assert _bound_ratio(Xi.data, ref, sabs) < 1.0, (nW, bl, bh, case)


@pytest.mark.gpu
def test_gpu_quickpol_lmax2047_sampled_columns_and_rescaling(ps, oracle, qp_variant):
    """Larger problem: every entry of the GPU band against the host build of the same arithmetic is not
    possible on the box (no /root/reference needed, but slow), so the oracle checks a sample of pairs."""
    lmax, bl, bh = 2047, 64, 64
    W = _scan_spectrum(2 * lmax + 1)
    case = (2, -2, 2, 3)
    Xi = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, bl, bh), *case, ps.SpectralVector(W))
    rng = np.random.default_rng(11)
    nu1, nu2, s1, s2 = case
    for _ in range(60):
        l = int(rng.integers(3, lmax + 1))
        lpp = int(np.clip(l + rng.integers(-bh, bl + 1), 3, lmax))
        n1, f1 = oracle.w3j_family(l, lpp, -s1, -nu1, ld=True)
        n2, f2 = oracle.w3j_family(l, lpp, -s2, -nu2, ld=True)
        a = max(n1, n2)
        j = np.arange(a, min(l + lpp, W.size - 1) + 1)
        t = W[j] * f1[j - n1] * f2[j - n2]
        ref = (-1.0) ** ((s1 + s2 + nu1 + nu2) % 2) * t.sum()
        assert abs(Xi[lpp, l] - ref) <= 1e-10 * abs(ref) + 1e-13 * np.abs(t).sum(), (l, lpp)
    # spins close to l: rescaling path on the device
    lmax = 1000
    W = _scan_spectrum(2 * lmax + 1, seed=9)
    case = (2, 2, 990, -985)
    ref, sabs = _oracle_ref(oracle, case, lmax, W, 8, 8)
    Xi = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, 8, 8), *case, ps.SpectralVector(W))
    assert np.count_nonzero(ref) > 50
    # stretched symbols (|s| = l or l''): the recurrence itself is ill-conditioned there -- the Float64 ORACLE sits at
    # 0.18 of the bound on this case, the host build of the kernel arithmetic at 0.04 (tab) / 0.39 (simple)
    assert _bound_ratio(Xi.data, ref, sabs) < 5.0


@pytest.mark.gpu
def test_gpu_quickpol_device_api(ps, oracle, qp_variant):
    import torch
    lmax, bl, bh = 400, 20, 30
    W = _scan_spectrum(2 * lmax + 1)
    case = (0, 2, 1, -1)
    one = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, bl, bh), *case, ps.SpectralVector(W)).data
    # device-level call on column sub-ranges, padded leading dimension
    L = ps.lib()
    nb, ldb = bl + bh + 1, bl + bh + 4
    dW = torch.tensor(W, device="cuda")
    dX = torch.zeros((lmax + 1, ldb), device="cuda", dtype=torch.float64)      # row-major (col, row) = column-major band
    for a, b in [(0, 130), (130, 131), (131, lmax + 1)]:
        rc = L.psb200_quickpol_xi_dev(*case, lmax, dW.data_ptr(), W.size, bl, bh, dX.data_ptr(), ldb, a, b, None)
        assert rc == 0, L.psb200_last_error()
    torch.cuda.synchronize()
    assert np.array_equal(dX.cpu().numpy()[:, :nb].T, one)


def test_failed_call_leaves_matrix_untouched(ps):
    """Odd s1+s2+nu1+nu2: the mirror pre-scales the storage by the final sign; a failing call must undo that."""
    Xi = ps.BandedSpectralMatrix(7, 2, 2)
    Xi.data[:] = 3.0
    with pytest.raises(ValueError):
        ps.quickpolXi(Xi, 0, 0, 1, 0, ps.SpectralVector(np.zeros(0)))       # empty W -> bad-argument code 1
    assert np.all(Xi.data == 3.0)


@pytest.mark.gpu
def test_gpu_quickpol_host_call_across_gpus(ps):
    """Column bands over every visible GPU: bit-identical to the one-GPU call (each column is computed by the same
    code whichever device owns it).  Skipped on a one-GPU box; kept last in the suite."""
    n = ps.lib().psb200_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    lmax, bl, bh = 400, 20, 30
    W = _scan_spectrum(2 * lmax + 1)
    case = (0, 2, 1, -1)
    one = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, bl, bh), *case, ps.SpectralVector(W)).data
    for ng in sorted({2, n}):
        many = ps.quickpolXi(ps.BandedSpectralMatrix(lmax, bl, bh), *case, ps.SpectralVector(W), ngpus=ng).data
        assert np.array_equal(one, many), ng


@pytest.mark.parametrize("variant", VARIANTS)
def test_kernel_arithmetic_random_pairs(oracle, hostcheck, variant):
    """Property test (hypothesis): any admissible (l, l'', s1, nu1, s2, nu2) and window length -- degenerate
    families (one or two terms, |s| = l, |nu| = l'', families that overlap in one term, B(j) = 0) included."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=400, deadline=None, derandomize=True)
    @given(st.integers(2, 70), st.integers(2, 70), st.integers(-70, 70), st.integers(-4, 4), st.integers(-70, 70),
           st.integers(-4, 4), st.integers(1, 150), st.integers(0, 2 ** 31 - 1))
    def check(l, lpp, s1, nu1, s2, nu2, nW, seed):
        s1 = max(-l, min(l, s1))
        s2 = max(-l, min(l, s2))
        nu1 = max(-lpp, min(lpp, nu1))
        nu2 = max(-lpp, min(lpp, nu2))
        W = np.random.default_rng(seed).normal(size=nW)
        n1, f1 = oracle.w3j_family(l, lpp, -s1, -nu1, ld=True)
        n2, f2 = oracle.w3j_family(l, lpp, -s2, -nu2, ld=True)
        a, e = max(n1, n2), min(l + lpp, nW - 1)
        ref, sab = 0.0, 0.0
        if e >= a and f1.size and f2.size:
            j = np.arange(a, e + 1)
            t = W[j] * f1[j - n1] * f2[j - n2]
            ref = (-1.0) ** ((s1 + s2 + nu1 + nu2) % 2) * t.sum()
            sab = np.abs(t).sum()
        got = hostcheck.pair(l, lpp, nu1, nu2, s1, s2, W, variant=variant)
        assert abs(got - ref) <= 1e-10 * abs(ref) + 1e-13 * sab + 1e-300, (l, lpp, s1, nu1, s2, nu2, nW, got, ref)
    check()
