"""W-spectrum production: spin-0 HEALPix map2alm / alm2map / alm2cl.

CPU part: pins of the oracle (oracle/shtoracle.py: direct long-double sums) and the arithmetic of the CUDA path
compiled for the host (tests/hostcheck/sht_host.cpp) against it.  GPU part: libpsb200 through the C ABI against
the oracle at sizes it finishes in seconds, against the host build at medium sizes, and through size-independent
properties at nside 1024 / 2048.

Tolerances.  alm: |gpu - ref| <= 1e-10 * max|ref alm| (the north-star 1e-10, relative to the scale of the transform: a
single a_lm of a random map is a sum of 12 nside^2 terms of either sign).  Maps likewise.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import shtoracle as so  # noqa: E402

LD = np.longdouble
TOL = 1e-10


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def random_alm(rng, lmax, lcut=None):
    a = rng.normal(size=so.alm_size(lmax)) + 1j * rng.normal(size=so.alm_size(lmax))
    for l in range(lmax + 1):
        i = so.alm_index(lmax, l, 0)
        a[i] = a[i].real
    if lcut is not None:
        for m in range(lmax + 1):
            for l in range(max(m, lcut + 1), lmax + 1):
                a[so.alm_index(lmax, l, m)] = 0.0
    return a


# ---------------------------------------------------------------------------------------------------------------
# CPU: the oracle itself
# ---------------------------------------------------------------------------------------------------------------
def test_oracle_pixel_centres_closed_form():
    # nside = 1 (HEALPix paper, fig. 4 / eq. 4-9): z = 2/3, 0, -2/3; phi = pi/4 + k pi/2 on the caps, k pi/2 on the equator
    th, ph = so.pix2ang_ring(1)
    assert np.allclose(np.cos(th).astype(float), [2 / 3] * 4 + [0] * 4 + [-2 / 3] * 4, atol=1e-15)
    assert np.allclose((ph / np.pi).astype(float), [0.25, 0.75, 1.25, 1.75, 0, 0.5, 1, 1.5, 0.25, 0.75, 1.25, 1.75], atol=1e-15)
    # nside = 2: first ring z = 1 - 1/12, 4 pixels; second ring z = 2/3, 8 pixels at (j - 1/2) pi/4; third ring z = 1/3 unshifted
    th, ph = so.pix2ang_ring(2)
    assert np.allclose(np.cos(th[:4]).astype(float), 11 / 12) and np.allclose(np.cos(th[4:12]).astype(float), 2 / 3)
    assert np.allclose((ph[4:12] / np.pi).astype(float), (np.arange(8) + 0.5) / 4)
    assert np.allclose(np.cos(th[12:20]).astype(float), 1 / 3) and np.allclose((ph[12:20] / np.pi).astype(float), np.arange(8) / 4)
    assert np.allclose(np.cos(th[20:28]).astype(float), 0, atol=1e-15) and np.allclose((ph[20:28] / np.pi).astype(float), (np.arange(8) + 0.5) / 4)


@pytest.mark.parametrize("nside", [1, 2, 4, 16, 64])
def test_oracle_ring_table_matches_pixel_formulas(nside):
    nphi, start, z, phi0 = so.ring_table(nside)
    th, ph = so.pix2ang_ring(nside)
    assert nphi.sum() == so.npix(nside) and start[0] == 0 and np.all(start[1:] == np.cumsum(nphi)[:-1])
    for r in range(nphi.size):
        sl = slice(start[r], start[r] + nphi[r])
        assert np.allclose(np.cos(th[sl]).astype(float), float(z[r]), atol=2e-16)
        assert np.allclose(ph[sl].astype(float), (phi0[r] + 2 * np.pi * np.arange(nphi[r]) / nphi[r]).astype(float), atol=1e-14)
    # equal-area pixelisation: sum over rings of n_phi * dz-band = 2 (checked through Y_00: exact quadrature of a constant)
    are, aim = so.analysis(nside, 0, np.ones(so.npix(nside)))
    assert abs(float(are[0]) - np.sqrt(4 * np.pi)) < 1e-15


def test_oracle_lambda_vs_scipy():
    from scipy.special import sph_harm_y
    lmax = 48
    x = np.cos(np.linspace(0.03, 3.11, 19))
    for m in range(0, lmax + 1, 4):
        lam = so.lam_rows(lmax, m, x).astype(float)
        for l in range(m, lmax + 1):
            ref = sph_harm_y(l, m, np.arccos(x), 0.0).real
            assert np.max(np.abs(lam[l - m] - ref)) < 2e-13 * max(1.0, np.max(np.abs(ref)))


def test_oracle_synthesis_is_the_sum_of_spherical_harmonics():
    from scipy.special import sph_harm_y
    nside, lmax = 4, 9
    rng = np.random.default_rng(11)
    a = random_alm(rng, lmax)
    th, ph = so.pix2ang_ring(nside)
    f = np.zeros(so.npix(nside))
    for l in range(lmax + 1):
        for m in range(l + 1):
            t = a[so.alm_index(lmax, l, m)] * sph_harm_y(l, m, th.astype(float), ph.astype(float))
            f += t.real if m == 0 else 2 * t.real
    assert rel(so.alm2map(a, nside, lmax), f) < 1e-13


def test_oracle_jacobi_iterations_recover_a_band_limited_field():
    nside, lmax = 8, 10
    a = random_alm(np.random.default_rng(12), lmax)
    f = so.alm2map(a, nside, lmax)
    errs = [rel(so.map2alm(f, nside, lmax, it), a) for it in (0, 1, 3, 8)]
    assert errs[0] > 1e-3 and errs[1] < errs[0] / 5 and errs[2] < errs[1] / 20 and errs[3] < 1e-9


def test_oracle_alm2cl_matches_the_mirror():
    import powerspectra_jl_b200 as ps
    lmax = 17
    rng = np.random.default_rng(13)
    a, b = random_alm(rng, lmax), random_alm(rng, lmax)
    assert np.allclose(so.alm2cl(a, b, lmax), ps.alm2cl(ps.Alm(lmax, lmax, a), ps.Alm(lmax, lmax, b)), rtol=1e-13, atol=1e-15)


def test_oracle_zonal_map_agrees_with_the_gauss_legendre_statement():
    # a smooth zonal field sampled on HEALPix: m != 0 vanishes by symmetry for m not a multiple of 4; the a_l0 converge
    # to the Gauss-Legendre values of synthetic.ZonalSky (the statement behind psb200_zonal_alm) as iterations proceed
    from powerspectra_jl_b200 import synthetic as syn
    nside, lmax = 16, 20
    th, _ = so.pix2ang_ring(nside)
    g = lambda t: 1.0 + 0.5 * np.cos(t) ** 3 - 0.25 * np.cos(t) ** 6        # band limit 6
    alm = so.map2alm(g(th.astype(float)), nside, lmax, niter=8)
    zs = syn.ZonalSky(lmax)
    al0 = zs.al0(g(zs.theta))[0]
    got = np.array([alm[so.alm_index(lmax, l, 0)].real for l in range(lmax + 1)])
    assert np.max(np.abs(got - al0)) < 1e-9
    for m in (1, 2, 3, 5):
        i0 = so.alm_index(lmax, m, m)
        assert np.max(np.abs(alm[i0:i0 + lmax - m + 1])) < 1e-14


@pytest.mark.parametrize("nside,lmax", [(1, 3), (4, 15), (16, 47), (32, 95)])
def test_ring_based_cpu_restatement_vs_direct_sums(nside, lmax):
    """oracle/shtcpu.c (per-ring FFT + scaled recurrences + parity folding in plain C, the shape of the reference's
    libsharp path) against the direct long-double sums: the fast twin used where the direct oracle cannot go."""
    f = np.random.default_rng(40 + nside).normal(size=so.npix(nside))
    for it in (0, 3):
        assert rel(so.fast_map2alm(f, nside, lmax, it), so.map2alm(f, nside, lmax, it)) < 1e-12
    a = so.map2alm(f, nside, lmax, 0)
    assert rel(so.fast_alm2map(a, nside, lmax), so.alm2map(a, nside, lmax)) < 1e-12


# ---------------------------------------------------------------------------------------------------------------
# CPU: the arithmetic of the CUDA path, compiled for the host
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hc():
    import hostcheck
    hostcheck.sht_lib()
    return hostcheck


@pytest.mark.parametrize("n,shifted", [(4, 1), (8, 1), (12, 1), (20, 0), (4 * 37, 1), (256, 0), (4 * 509, 1), (4 * 625, 1), (4096, 0)])
def test_host_ring_transforms_any_length(hc, n, shifted):
    """Mixed-radix Stockham FFT (radices 4, 2, 3, 5, ..., large primes), real-input packing, aliasing of m >= n,
    phase of shifted rings -- against direct sums with exactly reduced angles."""
    rng = np.random.default_rng(n)
    f = rng.normal(size=n)
    mmax = 3 * n // 2 + 5
    k = np.arange(n)
    ref = np.zeros(mmax + 1, dtype=complex)
    F = rng.normal(size=mmax + 1) + 1j * rng.normal(size=mmax + 1)
    fr = np.zeros(n, dtype=LD)
    for m in range(mmax + 1):
        ang = LD(np.pi) * ((m * (2 * k + shifted)) % (2 * n)).astype(LD) / n
        ref[m] = 0.7 * complex(np.sum(f * np.cos(ang)), -np.sum(f * np.sin(ang)))
        t = F[m].real * np.cos(ang) - F[m].imag * np.sin(ang)
        fr += t if m == 0 else 2 * t
    assert np.max(np.abs(hc.sht_ring_analyse(f, shifted, 0.7, mmax) - ref)) < 2e-14 * np.sqrt(n)
    assert np.max(np.abs(hc.sht_ring_synthesise(F, n, shifted) - fr.astype(float))) < 5e-14 * np.sqrt(mmax)


def test_host_ring_transforms_property(hc):
    """Random ring lengths 4r (r up to 2048: every mix of radices the HEALPix caps produce): synthesis of a band-limited
    spectrum followed by analysis returns it (mmax < n/2: no aliasing), for shifted and unshifted rings."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(st.integers(min_value=1, max_value=2048), st.booleans(), st.integers(min_value=0, max_value=2 ** 31 - 1))
    def run(r, shifted, seed):
        n = 4 * r
        mmax = n // 2 - 1
        rng = np.random.default_rng(seed)
        F = rng.normal(size=mmax + 1) + 1j * rng.normal(size=mmax + 1)
        F[0] = F[0].real
        f = hc.sht_ring_synthesise(F, n, shifted)
        back = hc.sht_ring_analyse(f, shifted, 1.0 / n, mmax)
        assert np.max(np.abs(back - F)) < 1e-12 * np.sqrt(n)
    run()


@pytest.mark.parametrize("nside", [1, 2, 8, 64, 2048])
def test_host_ring_geometry(hc, nside):
    nphi, start, z, phi0 = so.ring_table(nside)
    N = nside
    for p in sorted(set(list(range(min(2 * N, 40))) + [N - 2, N - 1, N, 2 * N - 2, 2 * N - 1]) & set(range(2 * N))):
        g = hc.sht_ring(N, p)
        rs = 4 * N - 2 - p
        assert g[0] == nphi[p] and g[1] == start[p] and abs(g[3] - float(z[p])) < 2e-16 and g[5] == (phi0[p] > 0)
        assert abs(g[4] - float(np.sqrt((1 - z[p]) * (1 + z[p])))) < 2e-16
        if p < 2 * N - 1:
            assert g[2] == start[rs] and abs(g[3] + float(z[rs])) < 2e-16 and nphi[rs] == nphi[p] and (phi0[rs] > 0) == (phi0[p] > 0)


def test_host_scaled_recurrence_and_skip_margin(hc):
    """lambda_lm through the scaled recurrence (start from log2 lambda_mm, rescale every 16 steps, alive flag) against the
    plain long-double recurrence at the same Float64 cos(theta); what the ring skips or drops before it is alive is
    below 1e-100 / 1e-30."""
    N, lmax = 256, 767
    _, _, z, _ = so.ring_table(N)
    zd = z.astype(np.float64).astype(LD)
    seen_scaled = seen_skipped = False
    for m in (0, 1, 5, 100, 300, 500, 766, 767):
        ref = so.lam_rows(lmax, m, zd[:2 * N]).astype(float)
        for p in (0, 1, 2, 5, 17, 100, 255, 300, 511):
            lam, alive = hc.sht_lambda(N, lmax, m, p)
            if alive < 0:
                seen_skipped = True
                assert np.max(np.abs(ref[:, p])) < 1e-30
                continue
            sel = np.arange(m, lmax + 1) >= alive
            seen_scaled |= alive > m
            if sel.any():
                assert np.max(np.abs(lam[sel] - ref[sel, p])) < 1e-10       # l^2 eps growth of the recurrence at the pole
            if (~sel).any():
                assert np.max(np.abs(ref[~sel, p])) < 1e-100
    assert seen_scaled and seen_skipped


@pytest.mark.parametrize("nside,lmax", [(1, 3), (2, 5), (4, 11), (8, 23), (16, 47), (16, 63)])
def test_host_transforms_vs_oracle(hc, nside, lmax):
    f = np.random.default_rng(nside).normal(size=so.npix(nside))          # not band-limited: aliasing paths included
    for it in (0, 3):
        assert rel(hc.sht_map2alm(f, nside, lmax, it), so.map2alm(f, nside, lmax, it)) < 1e-12
    a = so.map2alm(f, nside, lmax, 0)
    assert rel(hc.sht_alm2map(a, nside, lmax), so.alm2map(a, nside, lmax)) < 1e-12


def test_sht_stats_accounting():
    import ctypes as C
    import powerspectra_jl_b200 as ps
    out = (C.c_longlong * 6)()
    assert ps.lib().psb200_sht_stats(64, 191, out) == 0
    exec_, live, warps, R, chunks, steps = list(out)
    naive = sum((191 - m + 1) for m in range(192)) * 128
    assert R == 4 and steps == 16 and chunks == 1 and warps == 192 and live <= naive and exec_ >= live
    assert ps.lib().psb200_sht_stats(2048, 6143, out) == 0
    assert 0.55 < out[1] / (sum(6144 - m for m in range(6144)) * 4096) < 0.9       # rings skipped near the poles
    assert out[1] / out[0] > 0.7
    assert ps.lib().psb200_sht_stats(48, 10, out) == 1 and ps.lib().psb200_sht_stats(64, 256, out) == 1


# ---------------------------------------------------------------------------------------------------------------
# GPU: libpsb200 through the C ABI
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("nside,lmax", [(1, 3), (2, 7), (4, 11), (8, 23), (8, 31), (16, 47), (32, 95)])
def test_gpu_transforms_vs_oracle(ps, nside, lmax):
    f = np.random.default_rng(100 + nside).normal(size=so.npix(nside))
    for it in (0, 3):
        got = ps.map2alm(ps.HealpixMap(f), lmax=lmax, niter=it)
        assert got.lmax == lmax and got.mmax == lmax
        assert rel(got.alm, so.map2alm(f, nside, lmax, it)) < TOL
    a = so.map2alm(f, nside, lmax, 0)
    assert rel(ps.alm2map(ps.Alm(lmax, lmax, a), nside).pixels, so.alm2map(a, nside, lmax)) < TOL
    b = so.map2alm(f[::-1].copy(), nside, lmax, 0)
    assert np.allclose(ps.alm2cl_device(ps.Alm(lmax, lmax, a), ps.Alm(lmax, lmax, b)), so.alm2cl(a, b, lmax), rtol=1e-12, atol=1e-18)


@pytest.mark.gpu
@pytest.mark.parametrize("R", ["2", "8"])
def test_gpu_ring_pairs_per_lane_variants(ps, R, monkeypatch):
    nside, lmax = 32, 95
    f = np.random.default_rng(7).normal(size=so.npix(nside))
    base = ps.map2alm(ps.HealpixMap(f), lmax=lmax, niter=1).alm
    monkeypatch.setenv("PSB200_SHT_R", R)
    try:
        got = ps.map2alm(ps.HealpixMap(f), lmax=lmax, niter=1).alm
    finally:
        monkeypatch.delenv("PSB200_SHT_R")
    assert rel(got, base) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("nside,lmax", [(64, 191), (128, 383), (128, 511)])
def test_gpu_vs_host_build_medium(ps, hc, nside, lmax):
    """Every ring length 4..4 nside (all prime factors), scaled start values in use, several chunks per m."""
    f = np.random.default_rng(nside).normal(size=so.npix(nside))
    got = ps.map2alm(ps.HealpixMap(f), lmax=lmax, niter=1).alm
    ref = hc.sht_map2alm(f, nside, lmax, 1)
    assert rel(got, ref) < 1e-12
    assert rel(ps.alm2map(ps.Alm(lmax, lmax, ref), nside).pixels, hc.sht_alm2map(ref, nside, lmax)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("nside,lmax,niter", [(256, 767, 3), (512, 1535, 1), (512, 2047, 0)])
def test_gpu_vs_ring_based_cpu_restatement(ps, nside, lmax, niter):
    """Medium sizes against the independent CPU restatement (oracle/shtcpu.c: other FFT, other scaling scheme, no ring
    skipping): 1e-10 of the alm scale on a random, not band-limited map, lmax up to 4 nside - 1."""
    f = np.random.default_rng(nside + lmax).normal(size=so.npix(nside))
    ref = so.fast_map2alm(f, nside, lmax, niter)
    assert rel(ps.map2alm(ps.HealpixMap(f), lmax=lmax, niter=niter).alm, ref) < TOL
    assert rel(ps.alm2map(ps.Alm(lmax, lmax, ref), nside).pixels, so.fast_alm2map(ref, nside, lmax)) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("nside", [1024, 2048])
def test_gpu_full_size_properties(ps, nside):
    """BASELINE-size maps (nside 2048 <-> lmax 6143): a constant map is Y_00 alone; a band-limited field (lmax = 1.5 nside)
    synthesised on the device is recovered by map2alm as the Jacobi iterations proceed (the HEALPix quadrature is not
    exact: 1e-3 without iterations, ~1e-7 after the default 3, below 1e-9 after 8); Parseval ties alm2cl to the pixel
    variance; two runs are bit-identical (nothing is accumulated atomically)."""
    lmax = 3 * nside // 2
    one = ps.map2alm(ps.HealpixMap(np.ones(12 * nside * nside)), lmax=lmax, niter=0).alm
    assert abs(one[0] - np.sqrt(4 * np.pi)) < 1e-13
    assert np.max(np.abs(one[1:])) < 2e-3 and np.max(np.abs(one[lmax + 1:2 * lmax])) < 1e-13      # m = 1: zero by symmetry
    rng = np.random.default_rng(nside)
    n = so.alm_size(lmax)
    a = rng.normal(size=n) + 1j * rng.normal(size=n)
    a[:lmax + 1] = a[:lmax + 1].real
    a *= np.exp(-0.5 * (np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)]) / (0.4 * lmax)) ** 2)
    f = ps.alm2map(ps.Alm(lmax, lmax, a), nside)
    e0 = rel(ps.map2alm(f, lmax=lmax, niter=0).alm, a)
    back = ps.map2alm(f, lmax=lmax, niter=3)
    e3 = rel(back.alm, a)
    assert 1e-5 < e0 < 1e-2 and e3 < 1e-3 * e0 and e3 < 1e-6
    again = ps.map2alm(f, lmax=lmax, niter=3)
    assert np.array_equal(back.alm, again.alm)
    back = ps.map2alm(f, lmax=lmax, niter=8)
    assert rel(back.alm, a) < 1e-9
    cl = ps.alm2cl_device(back)
    parseval = float(np.sum((2 * np.arange(lmax + 1) + 1) * cl) / (4 * np.pi))
    assert abs(parseval - float(np.mean(f.pixels ** 2))) < 1e-6 * parseval


@pytest.mark.gpu
def test_gpu_workspace_window_spectra_vs_oracle(ps):
    """CovarianceWorkspace built from maps as in the reference (src/workspace.jl:110-135): effective_weight_alm! and
    window_function_W! (:141-213) against the oracle, then straight into coupledcov."""
    nside, lmax = 16, 40
    rng = np.random.default_rng(5)
    th, ph = so.pix2ang_ring(nside)
    th, ph = th.astype(float), ph.astype(float)

    def mask(k):
        return 0.5 * (1 + np.tanh(4 * (np.sin(th) * np.cos(ph - k) + 0.3 * np.cos(2 * th) + 0.2)))

    def var(k):
        return 1.0 + 0.5 * np.cos(th + k) ** 2 + 0.1 * rng.random(th.size)

    F = []
    for k, name in enumerate("AB"):
        F.append(ps.CovField(name, mask(k), mask(k + 0.5), ps.PolarizedHealpixMap(var(k), var(k + 1), var(k + 2))))
    ws = ps.CovarianceWorkspace.from_fields(F[0], F[1], F[0], F[1], lmax=lmax)
    N_ = ps.covariance.NULL
    # Omega_p = 4 pi / npix scales the noise-weighted products (:160-162)
    w_or = so.effective_weight_alm(nside, lmax, F[0].maskT.pixels, F[0].maskP.pixels, F[0].sigma2.q.pixels)
    w_gpu = ps.effective_weight_alm(ws, "QQ", "A", "A", "TP")
    assert rel(w_gpu.alm, w_or) < TOL
    assert not np.any(ps.effective_weight_alm(ws, "II", "A", "B", "TT").alm)               # i != j: zero alm (:170)
    # W^{0 PP, AB TT, AB PP}: mean over (QQ, UU) on the Y side; i != j kills the noise term -> zero
    W = ps.window_function_W(ws, N_, "PP", "A", "B", "TT", "A", "B", "PP")
    assert len(W) == lmax + 1 and not np.any(W.parent)
    # W^{TT PP, AA TT, BB PP} = 1/2 [cl(II_AA^TT, QQ_BB^PP) + cl(II_AA^TT, UU_BB^PP)]
    W = ps.window_function_W(ws, "TT", "PP", "A", "A", "TT", "B", "B", "PP")
    a = so.effective_weight_alm(nside, lmax, F[0].maskT.pixels, F[0].maskT.pixels, F[0].sigma2.i.pixels)
    bq = so.effective_weight_alm(nside, lmax, F[1].maskP.pixels, F[1].maskP.pixels, F[1].sigma2.q.pixels)
    bu = so.effective_weight_alm(nside, lmax, F[1].maskP.pixels, F[1].maskP.pixels, F[1].sigma2.u.pixels)
    ref = 0.5 * (so.alm2cl(a, bq, lmax) + so.alm2cl(a, bu, lmax))
    assert np.max(np.abs(W.parent - ref)) < TOL * np.max(np.abs(ref))
    # and the cache is the reference's: second lookup returns the stored vector
    assert ps.window_function_W(ws, "TT", "PP", "A", "A", "TT", "B", "B", "PP") is W
    # plain-mask spectrum feeding coupledcov TTTT end to end (values checked by the covariance parity tests)
    sp = {(s, x, y): ps.SpectralVector(np.ones(lmax + 1)) for s in ("TT",) for x in "AB" for y in "AB"}
    C_ = ps.coupledcov("TT", "TT", ws, sp, lmin=2)
    assert np.all(np.isfinite(C_.parent)) and np.allclose(C_.parent, C_.parent.T)


def _two_field_workspace(ps, nside, lmax, seed=5):
    rng = np.random.default_rng(seed)
    th, ph = so.pix2ang_ring(nside)
    th, ph = th.astype(float), ph.astype(float)
    mask = lambda k: 0.5 * (1 + np.tanh(4 * (np.sin(th) * np.cos(ph - k) + 0.3 * np.cos(2 * th) + 0.2)))
    var = lambda k: 1.0 + 0.5 * np.cos(th + k) ** 2 + 0.1 * rng.random(th.size)
    F = [ps.CovField(name, mask(k), mask(k + 0.5), ps.PolarizedHealpixMap(var(k), var(k + 1), var(k + 2)))
         for k, name in enumerate("AB")]
    return ps.CovarianceWorkspace.from_fields(F[0], F[1], F[0], F[1], lmax=lmax)


def _batched_weights_case(ps, ngpus):
    """psb200_map2alm_many (unique maps uploaded once, products dealt to `ngpus` devices) against the one-at-a-time path:
    every effective weight the TTTT and EEEE window spectra of a two-field workspace need, bit for bit."""
    nside, lmax = 32, 80
    N_ = ps.covariance.NULL
    W_keys = [(N_, N_, "A", "A", "TT", "B", "B", "TT"), (N_, "TT", "A", "B", "TT", "A", "B", "TT"),
              ("TT", "TT", "A", "A", "TT", "B", "B", "TT"), ("PP", "PP", "A", "A", "PP", "B", "B", "PP"),
              (N_, "PP", "B", "B", "PP", "A", "A", "PP"), ("TT", "PP", "A", "A", "TP", "B", "B", "PT")]
    keys = ps.weights_needed(W_keys)
    assert ("QQ", "A", "A", "PP") in keys and ("II", "A", "B", "TT") in keys
    one = _two_field_workspace(ps, nside, lmax)
    many = _two_field_workspace(ps, nside, lmax)
    n = ps.precompute_effective_weights(many, keys, ngpus=ngpus)
    assert n == len([k for k in keys if k[0] == N_ or k[1] == k[2]])            # i != j noise weights are zero: not computed
    assert ps.precompute_effective_weights(many, keys, ngpus=ngpus) == 0       # all cached now
    for k in keys:
        a = ps.effective_weight_alm(one, *k)
        b = ps.effective_weight_alm(many, *k)
        assert np.array_equal(a.alm, b.alm), k
    for wk in W_keys:
        assert np.array_equal(ps.window_function_W(one, *wk).parent, ps.window_function_W(many, *wk).parent)


@pytest.mark.gpu
def test_gpu_batched_effective_weights(ps):
    _batched_weights_case(ps, 1)


@pytest.mark.gpu
def test_gpu_batched_effective_weights_across_gpus(ps):
    if ps.lib().psb200_device_count() < 2:
        pytest.skip("needs two GPUs")
    _batched_weights_case(ps, 2)
    _batched_weights_case(ps, 0)


@pytest.mark.gpu
def test_gpu_sht_argument_errors(ps):
    with pytest.raises(ValueError):
        ps.HealpixMap(np.ones(12 * 9))                                   # nside 3
    f = ps.HealpixMap(np.ones(48))
    with pytest.raises(ValueError):
        ps.map2alm(f, lmax=8)                                            # lmax > 4 nside - 1
    with pytest.raises(ValueError):
        ps.map2alm(f, lmax=3, niter=-1)
    import ctypes as C
    DP = ps._lib.DP
    m = np.ones(48)
    out = np.zeros(10, dtype=np.complex128)
    mp, op = (DP * 1)(m.ctypes.data_as(DP)), (DP * 1)(out.ctypes.data_as(DP))
    sc = np.ones(1)
    call = lambda ix, ng=1: ps.lib().psb200_map2alm_many(2, 3, 0, 1, mp, 1, (C.c_int * 3)(*ix), sc.ctypes.data_as(DP), op, ng)
    assert call([0, -1, -1]) == 0 and abs(out[0] - np.sqrt(4 * np.pi)) < 1e-14
    assert call([-1, 0, -1]) == 1 and call([0, 1, -1]) == 1 and call([0, -1, 0]) == 1 and call([0, -1, -1], 99) == 1
    assert ps.map2alm(f, lmax=0, niter=0).alm.size == 1
    from powerspectra_jl_b200 import healpix
    healpix.release_transform_buffers()                                  # and the next call rebuilds its plan
    assert abs(ps.map2alm(f, lmax=3, niter=0).alm[0] - np.sqrt(4 * np.pi)) < 1e-14
