"""W-spectrum production, first slice: psb200_zonal_alm -- the m = 0 map2alm of zonal maps by
Gauss-Legendre quadrature -- against the host statement of the same sum (synthetic.ZonalSky.al0) and against closed forms.
Reference: effective_weight_alm! / window_function_W!, /root/reference/src/workspace.jl:141-213."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_zonal_alm_matches_host_quadrature(ps):
    from powerspectra_jl_b200 import synthetic as syn
    lmax = 767
    sky = syn.ZonalSky(lmax)
    prof = [syn.mask_profile(sky.theta, s) for s in (1001, 1002, 1003, 1004)]
    # masks, mask products and a noise-weighted product: what effective_weight_alm! transforms
    fields = prof + [prof[0] * prof[1], prof[2] * prof[3], prof[0] * prof[0] * (1.0 + 0.5 * sky.x ** 2)]
    host = sky.al0(fields)
    dev = sky.al0_device(fields)
    scale = np.max(np.abs(host), axis=1, keepdims=True)
    assert np.max(np.abs(dev - host) / scale) < 2e-14
    # the window spectrum built from them (window_function_W!: alm2cl of two effective-weight alms)
    Wh, Wd = sky.cross(host[4], host[5]), sky.cross(dev[4], dev[5])
    assert np.max(np.abs(Wd - Wh)) < 1e-13 * np.max(np.abs(Wh))


def test_zonal_alm_closed_forms(ps):
    """f = 1 -> a_00 = sqrt(4 pi), everything else 0; f = P_7 -> only a_70 = sqrt(4 pi / 15); more fields than one launch holds."""
    from powerspectra_jl_b200 import synthetic as syn
    lmax = 200
    sky = syn.ZonalSky(lmax)
    x = sky.x
    p = [np.ones_like(x), x.copy()]
    for l in range(2, 8):
        p.append(((2 * l - 1) * x * p[-1] - (l - 1) * p[-2]) / l)
    fields = [np.ones_like(x), p[7]] + [p[k % 8] * (1.0 + 0.01 * k) for k in range(40)]      # 42 fields: two launches
    a = sky.al0_device(fields)
    assert a.shape == (42, lmax + 1)
    assert abs(a[0, 0] - np.sqrt(4 * np.pi)) < 1e-13 and np.max(np.abs(a[0, 1:])) < 1e-13
    assert abs(a[1, 7] - np.sqrt(4 * np.pi / 15.0)) < 1e-13
    assert np.max(np.abs(np.delete(a[1], 7))) < 1e-13
    assert np.max(np.abs(a[2:] - sky.al0(fields[2:]))) < 1e-13


def test_zonal_alm_error_codes(ps):
    lib, DP = ps.lib(), ps._lib.DP
    x = np.linspace(-0.9, 0.9, 16); w = np.ones(16); f = np.ones((2, 16)); a = np.zeros((2, 9))
    dp = lambda v: v.ctypes.data_as(DP)
    assert lib.psb200_zonal_alm(2, 16, dp(x), dp(w), dp(f), 16, 8, dp(a), 9) == 0
    assert lib.psb200_zonal_alm(0, 16, dp(x), dp(w), dp(f), 16, 8, dp(a), 9) == 1
    assert lib.psb200_zonal_alm(2, 16, dp(x), dp(w), dp(f), 8, 8, dp(a), 9) == 1        # ldf < nnodes
    assert lib.psb200_zonal_alm(2, 16, dp(x), dp(w), dp(f), 16, 8, dp(a), 4) == 1       # lda < lmax+1
