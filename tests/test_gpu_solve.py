"""Decoupling on the device: psb200_mcm_solve, psb200_master_solve, psb200_decouple_covmat against host
LAPACK on the same matrices.  Reference: `M \\ pCl` src/blockspectralmatrix.jl:124-129, the 2N x 2N block systems
src/modecoupling.jl:213-223 + src/blockspectralmatrix.jl:89-122, maskedalm2spectra src/modecoupling.jl:341-377,
decouple_covmat src/covariance.jl:8-14 (its own test: test/test_covmat.jl:11-18)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10          # on the solution, relative to its largest entry


@pytest.fixture(scope="module")
def masks(ps):
    from powerspectra_jl_b200 import synthetic as syn
    lmax = 383
    sky = syn.ZonalSky(lmax)
    al = sky.al0([syn.mask_profile(sky.theta, s) for s in (1001, 1002, 1003, 1004)])
    return lmax, [ps.Alm.zonal(a) for a in al]


def _close(x, ref, rtol=RTOL):
    x, ref = np.asarray(x), np.asarray(ref)
    return float(np.max(np.abs(x - ref)) / np.max(np.abs(ref))) <= rtol


@pytest.mark.parametrize("spec", ["TT", "TE", "M++"])
@pytest.mark.parametrize("lmin", [0, 2])
def test_mcm_solve_single_systems(ps, masks, spec, lmin):
    lmax, (mT1, mP1, mT2, mP2) = masks
    if spec != "TT" and lmin < 2:
        pytest.skip("spin-2 rows l < 2 are the reference's don't-care region (singular there)")
    a, b = {"TT": (mT1, mT2), "TE": (mT1, mP2), "M++": (mP1, mP2)}[spec]
    M = ps.mcm(spec, a, b, lmin=lmin).parent
    rng = np.random.default_rng(3)
    N = lmax + 1 - lmin
    cl = rng.uniform(0.5, 1.5, size=N)
    pcl = M @ cl
    x = ps.mcm_solve(spec, a, b, ps.SpectralVector(pcl, lmin), lmin=lmin)
    assert x.offsets == (lmin,)
    assert _close(x.parent, np.linalg.solve(M, pcl))
    assert _close(x.parent, cl, 1e-9)
    # several right-hand sides at once
    P = np.asfortranarray(rng.normal(size=(N, 3)))
    X = ps.mcm_solve(spec, a, b, P, lmin=lmin)
    assert _close(X.parent, np.linalg.solve(M, P))


@pytest.mark.parametrize("spec", ["EE_BB", "EB_BE"])
def test_mcm_solve_block_systems(ps, masks, spec):
    lmax, (mT1, mP1, mT2, mP2) = masks
    lmin = 2
    B = ps.mcm(spec, mP1, mP2, lmin=lmin)
    rng = np.random.default_rng(4)
    N = lmax + 1 - lmin
    x1, x2 = rng.uniform(0.5, 1.5, size=N), rng.uniform(0.05, 0.15, size=N)
    rhs = B.parent @ np.concatenate([x1, x2])
    ref = np.linalg.solve(B.parent, rhs)
    y1, y2 = ps.mcm_solve(spec, mP1, mP2, [ps.SpectralVector(rhs[:N], lmin), ps.SpectralVector(rhs[N:], lmin)], lmin=lmin)
    assert _close(np.concatenate([y1.parent, y2.parent]), ref)
    assert _close(y1.parent, x1, 1e-9) and _close(y2.parent, x2, 1e-8)
    # the host path of the mirror (reference-shaped: lu of the dense hvcat) gives the same spectra
    h1, h2 = B.solve(rhs)
    assert _close(y1.parent, h1.parent) and _close(y2.parent, h2.parent)


def test_master_solve_matches_host_solves(ps, masks):
    """maskedalm2spectra with the solves on the device == the mirror's host solves on the same matrices."""
    lmax, (mT1, mP1, mT2, mP2) = masks
    lmin = 2
    rng = np.random.default_rng(5)
    maps1 = [ps.Alm.zonal(rng.normal(size=lmax + 1)) for _ in range(3)]
    maps2 = [ps.Alm.zonal(rng.normal(size=lmax + 1)) for _ in range(3)]
    dev = ps.maskedalm2spectra_device(maps1, mT1, mP1, maps2, mT2, mP2, lmin=lmin)
    host = ps.maskedalm2spectra(maps1, mT1, mP1, maps2, mT2, mP2, lmin=lmin)
    assert set(dev) == set(host) == {"TT", "TE", "ET", "TB", "BT", "EE", "BB", "EB", "BE"}
    for k in host:
        assert dev[k].offsets == host[k].offsets == (lmin,)
        assert _close(dev[k].parent, host[k].parent), k


def test_decouple_covmat_device(ps):
    """test/test_covmat.jl:11-18: decouple_covmat(A, B1, B2) == inv(B1) A inv(B2)'."""
    rng = np.random.default_rng(6)
    n = 300
    A = rng.normal(size=(n, n))
    B1 = rng.normal(size=(n, n)) + 10 * np.eye(n)
    B2 = rng.normal(size=(n, n)) + 10 * np.eye(n)
    S = ps.SpectralArray
    out = ps.decouple_covmat_device(S(A), S(B1), S(B2)).parent
    ref = np.linalg.inv(B1) @ A @ np.linalg.inv(B2).T
    assert _close(out, ref)
    assert _close(out, ps.decouple_covmat(S(A), S(B1), S(B2)).parent)
    # non-square leading dimensions through the C ABI
    lib, DP = ps.lib(), ps._lib.DP
    ld = n + 7
    pad = lambda M: np.asfortranarray(np.vstack([M, np.zeros((7, n))]))
    y, b1, b2 = pad(A), pad(B1), pad(B2)
    o = np.zeros((ld, n), order="F")
    rc = lib.psb200_decouple_covmat(n, y.ctypes.data_as(DP), ld, b1.ctypes.data_as(DP), ld, b2.ctypes.data_as(DP), ld,
                                    o.ctypes.data_as(DP), ld)
    assert rc == 0
    assert _close(o[:n], ref)


def test_decouple_coupled_covariance_on_device(ps, oracle):
    """The chain the reference documents: C = coupledcov(TT, TT); decouple_covmat(C, M_TT, M_TT), all three matrices from
    the GPU, decoupled on the GPU, against the host solve."""
    from powerspectra_jl_b200 import synthetic as syn
    lmax, lmin = 255, 2
    ws, sp, rt = syn.covariance_inputs(lmax)
    C = ps.coupledcov("TT", "TT", ws, sp, rt, lmin=lmin)
    V = syn.mask_spectra(lmax, seeds=(1001, 1003))[(0, 1)]
    M = ps.mcm("TT", ps.SpectralVector(V), lmin=lmin)
    D = ps.decouple_covmat_device(C, M, M)
    H = ps.decouple_covmat(C, M, M)
    assert D.offsets == C.offsets
    assert _close(D.parent, H.parent)


def test_solve_error_codes(ps, masks):
    lmax, (mT1, mP1, mT2, mP2) = masks
    lib, DP = ps.lib(), ps._lib.DP
    V = np.ones(16)
    p, c = np.ones(16), np.zeros(16)
    dp = lambda a: a.ctypes.data_as(DP)
    assert lib.psb200_mcm_solve(7, 0, 15, dp(V), 16, dp(p), 16, 1, dp(c), 16, 1) == 1       # unknown system
    assert lib.psb200_mcm_solve(0, 0, 15, dp(V), 16, dp(p), 8, 1, dp(c), 16, 1) == 1        # ldp < N
    assert lib.psb200_mcm_solve(4, 2, 15, dp(V), 16, dp(p), 16, 1, dp(c), 16, 1) == 1       # block system needs 2N rows
    # an exactly singular system (V = 0 => M = 0): code 6, LinAlgError in the mirror (Julia: SingularException)
    Z = np.zeros(16)
    assert lib.psb200_mcm_solve(0, 0, 15, dp(Z), 16, dp(p), 16, 1, dp(c), 16, 1) == 6
    with pytest.raises(np.linalg.LinAlgError):
        ps.mcm_solve("TT", ps.SpectralVector(Z), None, p)


def test_solves_across_gpus(ps, masks):
    """ngpus = 2: the row bands of the other device arrive in the root's matrix through peer stores of its pair kernel
    (or one peer copy); the assembled matrix -- hence the LU and the solution -- must be bit-identical to the 1-GPU call.
    Also with the explicit-copy path forced (PSB200_NO_PEER_STORES)."""
    import os
    if ps.lib().psb200_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    lmax, (mT1, mP1, mT2, mP2) = masks
    lmin = 2
    rng = np.random.default_rng(8)
    N = lmax + 1 - lmin
    p1 = rng.normal(size=N)
    one = ps.mcm_solve("TT", mT1, mT2, ps.SpectralVector(p1, lmin), lmin=lmin, ngpus=1).parent
    two = ps.mcm_solve("TT", mT1, mT2, ps.SpectralVector(p1, lmin), lmin=lmin, ngpus=2).parent
    assert np.array_equal(one, two)
    pb = [ps.SpectralVector(rng.normal(size=N), lmin), ps.SpectralVector(rng.normal(size=N), lmin)]
    b1 = ps.mcm_solve("EE_BB", mP1, mP2, pb, lmin=lmin, ngpus=1)
    b2 = ps.mcm_solve("EE_BB", mP1, mP2, pb, lmin=lmin, ngpus=2)
    assert np.array_equal(b1[0].parent, b2[0].parent) and np.array_equal(b1[1].parent, b2[1].parent)
    maps1 = [ps.Alm.zonal(rng.normal(size=lmax + 1)) for _ in range(3)]
    maps2 = [ps.Alm.zonal(rng.normal(size=lmax + 1)) for _ in range(3)]
    d1 = ps.maskedalm2spectra_device(maps1, mT1, mP1, maps2, mT2, mP2, lmin=lmin, ngpus=1)
    d2 = ps.maskedalm2spectra_device(maps1, mT1, mP1, maps2, mT2, mP2, lmin=lmin, ngpus=2)
    for k in d1:
        assert np.array_equal(d1[k].parent, d2[k].parent), k
    os.environ["PSB200_NO_PEER_STORES"] = "1"
    try:
        # a fresh pair of devices is not needed: the flag is read per call, peer access already enabled is simply not used
        three = ps.mcm_solve("TT", mT1, mT2, ps.SpectralVector(p1, lmin), lmin=lmin, ngpus=2).parent
    finally:
        del os.environ["PSB200_NO_PEER_STORES"]
    assert np.array_equal(one, three)
