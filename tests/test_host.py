"""CPU-side tests: the C-ABI library loads and exports every symbol include/psb200.h declares,
argument checking and the no-device error, host containers, band partitioning."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(ps):
    hdr = open(os.path.join(ROOT, "include", "psb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(psb200_\w+)\s*\(", hdr))
    assert declared == set(ps._lib.PROTOTYPES), declared ^ set(ps._lib.PROTOTYPES)
    L = ps.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.psb200_version().startswith(b"psb200")


def test_no_cpu_fallback(ps):
    """Without a device every compute call fails loudly (code 5); bad arguments give 1 first."""
    L = ps.lib()
    V = np.ones(8)
    M = np.zeros((8, 8), order="F")
    vp, mp = V.ctypes.data_as(ps._lib.DP), M.ctypes.data_as(ps._lib.DP)
    assert L.psb200_mcm(7, 0, 7, vp, 8, mp, 8, None, 1) == 1
    assert L.psb200_mcm(0, 3, 2, vp, 8, mp, 8, None, 1) == 1
    assert L.psb200_mcm(0, 0, 7, vp, 8, mp, 2, None, 1) == 1
    if L.psb200_device_count() == 0:
        assert L.psb200_mcm(0, 0, 7, vp, 8, mp, 8, None, 1) == 5
        assert b"no CPU fallback" in L.psb200_last_error()
        with pytest.raises(ps.PSB200Error):
            ps.mcm("TT", ps.SpectralVector(V))
        assert np.all(M == 0.0)
    with pytest.raises(ValueError):
        ps.mcm("TT", ps.SpectralVector(V), lmin=5, lmax=3)


def test_cov_arity_checked(ps):
    L = ps.lib()
    v = np.ones(8)
    p = (ps._lib.DP * 8)(*[v.ctypes.data_as(ps._lib.DP)] * 8)
    Cm = np.zeros((8, 8), order="F")
    cp = Cm.ctypes.data_as(ps._lib.DP)
    assert L.psb200_cov(0, 0, 7, p, 4, p, 4, p, 7, 8, cp, 8, 1) == 1     # TTTT needs 8 W
    assert L.psb200_cov(6, 0, 7, p, 4, p, 2, p, 2, 8, cp, 8, 1) == 1     # TTEE takes no ratios
    assert L.psb200_cov(7, 0, 7, p, 4, p, 4, p, 8, 8, cp, 8, 1) == 1     # unknown block


def test_band_edges_balance(ps):
    from powerspectra_jl_b200 import device as dev
    for lmin, lmax, nb in [(0, 6143, 8), (2, 767, 4), (0, 12287, 8), (0, 5, 8), (10, 10, 3)]:
        e = dev.band_edges(lmin, lmax, nb, lenW=0)          # balance the reference's full-family terms
        assert e[0] == lmin and e[-1] == lmax + 1 and len(e) == nb + 1
        assert all(b >= a for a, b in zip(e, e[1:]))
        cost = [dev.terms("M00", lmax, a, b) for a, b in zip(e, e[1:])]
        assert sum(cost) == dev.terms("M00", lmax, lmin, lmax + 1)
        if lmax - lmin > 100 * nb:
            assert max(cost) / (sum(cost) / nb) < 1.02           # row granularity only
    # default: rows costed as the kernel runs them (l3 truncated at lenW = lmax+1): upper rows are cheap
    et = dev.band_edges(0, 6143, 2)
    ef = dev.band_edges(0, 6143, 2, lenW=0)
    assert et[1] < ef[1]
    l = np.arange(0, 6144)

    def kcost(a, b):
        return sum(int(np.minimum(2 * l1 + 1, 6144 - np.arange(0, 6144 - l1)).sum()) for l1 in range(a, b))
    # the cost model also charges the warp start skew and a per-block overhead, so plain term counts
    # of the two halves agree only roughly
    assert abs(kcost(et[0], et[1]) / kcost(et[1], et[2]) - 1) < 0.12
    assert dev.terms("M00", 6143, 0, 6144) == 77328286720
    assert dev.terms("M00", 767, 0, 768) == 151289984


def test_spectral_array_semantics(ps):
    """subset of /root/reference/test/test_spectralarray.jl: offsets, slicing, inverse."""
    A = ps.SpectralArray(np.arange(16.0).reshape(4, 4))
    assert A[0, 0] == 0.0 and A[3, 3] == 15.0
    B = ps.spectralzeros(range(2, 6), range(2, 6))
    B[2, 3] = 7.0
    assert B.parent[0, 1] == 7.0 and B.firstindex(0) == 2 and B.lastindex(1) == 5
    v = ps.SpectralVector(np.arange(10.0))
    s = v[2:5]
    assert isinstance(s, ps.SpectralArray) and s.offsets == (2,) and s[2] == 2.0
    with pytest.raises(IndexError):
        v[10]
    M = ps.SpectralArray(np.array([[2.0, 1.0], [1.0, 3.0]]), (2, 2))
    x = M.solve(ps.SpectralVector(np.array([1.0, 2.0]), 2))
    assert np.allclose(M.parent @ x.parent, [1.0, 2.0]) and x.offsets == (2,)
    assert np.allclose(M.inv().parent @ M.parent, np.eye(2))
    z = ps.SpectralVector(np.arange(2.0, 8.0), 2).zero_based(6)
    assert np.array_equal(z, [0, 0, 2, 3, 4, 5, 6])


def test_block_matrix_and_decouple(ps):
    rng = np.random.default_rng(0)
    a = ps.SpectralArray(rng.normal(size=(3, 3)) + 4 * np.eye(3), (2, 2))
    b = ps.SpectralArray(0.1 * rng.normal(size=(3, 3)), (2, 2))
    blk = ps.BlockSpectralMatrix([[a, b], [b, a]])
    assert blk.parent.shape == (6, 6) and np.array_equal(blk.getblock(1, 0).parent, b.parent)
    e, bb = ps.SpectralVector(rng.normal(size=3), 2), ps.SpectralVector(rng.normal(size=3), 2)
    x, y = blk.solve([e, bb])
    assert np.allclose(blk.parent @ np.concatenate([x.parent, y.parent]), np.concatenate([e.parent, bb.parent]))
    # decouple_covmat == inv(B1) A inv(B2)'   (test/test_covmat.jl:11-18)
    A = np.array([[1.0, 0.2, 0.3], [0.4, 2.0, 0.15], [0.3, 0.1, 1.44]])
    B1 = np.array([[1.2, 0.6, 0.1], [0.3, 1.4, 0.5], [0.44, 0.2, 1.3]])
    B2 = np.array([[1.6, 0.4, 0.1], [0.3, 1.4, 0.9], [0.45, 0.8, 1.7]])
    Cd = ps.decouple_covmat(ps.SpectralArray(A), ps.SpectralArray(B1), ps.SpectralArray(B2))
    assert np.allclose(Cd.parent, np.linalg.inv(B1) @ A @ np.linalg.inv(B2).T)


def test_alm2cl_and_workspace(ps):
    a = ps.Alm.zonal([1.0, 2.0, 3.0])
    b = ps.Alm.zonal([2.0, 1.0, -1.0])
    assert np.allclose(ps.alm2cl(a, b), [2.0, 2.0 / 3, -3.0 / 5])
    rng = np.random.default_rng(1)
    full = ps.Alm(2, 2, rng.normal(size=6) + 1j * rng.normal(size=6))
    cl = ps.alm2cl(full, full)
    assert np.isclose(cl[2], (abs(full.alm[2]) ** 2 + 2 * abs(full.alm[4]) ** 2 + 2 * abs(full.alm[5]) ** 2) / 5)
    ws = ps.CovarianceWorkspace(("a", "b", "a", "b"), 4)
    with pytest.raises(KeyError):
        ps.window_function_W(ws, "TT", "TT", "a", "a", "TT", "b", "b", "TT")
    calls = []
    ws = ps.CovarianceWorkspace(("a", "b", "a", "b"), 4, provider=lambda *k: calls.append(k) or np.ones(5))
    w1 = ps.window_function_W(ws, "TT", "TT", "a", "a", "TT", "b", "b", "TT")
    w2 = ps.window_function_W(ws, "TT", "TT", "a", "a", "TT", "b", "b", "TT")
    assert w1 is w2 and len(calls) == 1                      # cached like workspace.W_spectra


def test_host_band_edges_follow_the_copies(ps):
    """TT at lmax 6143 is copy-bound on several GPUs (302 MB of result, 5 ms of kernels): the host-call bands must move
    towards equal bytes (first band much shorter than the kernel-balanced one); EEEE is kernel-bound and keeps the
    kernel-balanced edges to within a few per cent."""
    import ctypes as C
    from powerspectra_jl_b200 import device as dev
    kb = dev.band_edges(0, 6143, 8, lenW=6144)
    tt, ee = (C.c_int * 9)(), (C.c_int * 9)()
    assert ps.lib().psb200_host_band_edges(0, 0, 0, 6143, 6144, 8, tt) == 0
    assert ps.lib().psb200_host_band_edges(1, 1, 0, 6143, 6144, 8, ee) == 0
    assert tt[1] < 0.6 * kb[1]
    assert abs(ee[1] - kb[1]) < 0.1 * kb[1]
    N = 6144
    share = lambda e, g: sum(2 * (N - i) - 1 for i in range(e[g], e[g + 1])) / N ** 2
    assert max(share(tt, g) for g in range(8)) < 0.2 < max(share(kb, g) for g in range(8))


def test_folded_bands_partition(ps):
    """device.folded_bands: every row belongs to exactly one band of exactly one rank; each rank owns a low and a high piece."""
    from powerspectra_jl_b200 import device as dev
    for lmin, lmax, world in ((0, 6143, 8), (2, 767, 4), (0, 40, 2), (5, 9, 3), (0, 6143, 1)):
        owners = dev.folded_bands(lmin, lmax, world)
        assert len(owners) == world
        seen = np.zeros(lmax + 1 - lmin, dtype=int)
        for bands in owners:
            assert len(bands) == (1 if world == 1 else 2)
            for lo, hi in bands:
                assert lmin <= lo <= hi <= lmax + 1
                seen[lo - lmin:hi - lmin] += 1
        assert np.all(seen == 1)
        if world > 1 and lmax - lmin > 8 * world:
            assert all(b[0][1] <= b[1][0] for b in owners)        # low piece below the high piece


def test_band_edges_properties_random(ps):
    """Property test (hypothesis): for any shape the two partitioners return a monotone cover with the right ends,
    no band is empty while rows remain for the later ones only if the cost demands it, and with the reference
    term count (lenW = 0) every boundary sits exactly where the prefix sum crosses k/n of the total."""
    from hypothesis import given, settings, strategies as st
    from powerspectra_jl_b200 import device as dev

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 40), st.integers(0, 900), st.integers(1, 9), st.integers(0, 1200))
    def check(lmin, span, nb, lenW):
        lmax = lmin + span
        for lw in (0, lenW):
            e = dev.band_edges(lmin, lmax, nb, lenW=lw)
            assert len(e) == nb + 1 and e[0] == lmin and e[-1] == lmax + 1
            assert all(b >= a for a, b in zip(e, e[1:]))
        e = dev.band_edges(lmin, lmax, nb, lenW=0)
        l = np.arange(lmin, lmax + 1)
        cum = np.concatenate([[0], np.cumsum((2 * l + 1) * (lmax - l + 1))]).astype(float)
        for k in range(1, nb):
            i = e[k] - lmin                                  # rows [lmin, e[k]) belong to the first k bands
            assert cum[i] >= cum[-1] * k / nb * (1 - 1e-12)
            if i > 0 and e[k] > e[k - 1]:
                assert cum[i - 1] < cum[-1] * k / nb * (1 + 1e-12)
        # bands of the host-level multi-GPU calls (kernel seconds and copy seconds balanced jointly)
        import ctypes as C
        for api, code in ((0, 0), (0, 4), (1, 1), (2, 0)):
            h = (C.c_int * (nb + 1))()
            assert ps.lib().psb200_host_band_edges(api, code, lmin, lmax, max(lenW, 1), nb, h) == 0
            h = list(h)
            assert h[0] == lmin and h[-1] == lmax + 1 and all(b >= a for a, b in zip(h, h[1:]))
        # QuickPol column bands
        bl, bh = min(span, lenW % 50), min(span, lenW % 37)
        q = dev.quickpol_edges(lmax, bl, bh, nb)
        assert len(q) == nb + 1 and q[0] == 0 and q[-1] == lmax + 1
        assert all(b >= a for a, b in zip(q, q[1:]))
    check()


def test_host_result_buffers(ps):
    """psb200_host_alloc / free / placement (no device needed: registration is skipped without one)."""
    L = ps.lib()
    assert L.psb200_host_numa_nodes() >= 1
    assert L.psb200_host_alloc(0, 0) is None and b"host_alloc" in L.psb200_last_error()
    assert L.psb200_host_alloc(4096, 2) is None
    for interleave in (False, True):
        H = ps._lib.HostMatrix(700, interleave=interleave)        # 3.9 MB: two 2 MB pieces
        assert H._p % (2 << 20) == 0
        A = H.array
        assert A.shape == (700, 700) and A.flags.f_contiguous and not A.any()
        A[3, 5] = 2.5
        flat = np.ctypeslib.as_array(C.cast(H._p, ps._lib.DP), shape=(700 * 700,))
        assert flat[3 + 5 * 700] == 2.5                            # column-major, like parent(SpectralArray)
        pl = H.placement()
        assert pl is None or sum(pl) >= 1
        p = H._p
        H.free()
        assert L.psb200_host_free(p) == 1                          # already released
    assert L.psb200_host_free(None) == 0
    A = ps.pinned_spectralzeros(range(2, 302), interleave=True)       # the mirror of the shim's pinned_spectralzeros
    assert A.axes(0) == range(2, 302) and A.parent.flags.f_contiguous and not A.parent.any()
    A[5, 7] = 1.25
    assert A.parent[3, 5] == 1.25
    ps.free_pinned(A)
    with pytest.raises(ValueError):
        ps.free_pinned(ps.spectralzeros(4, 4))
    cnt = (C.c_int * 4)()
    assert L.psb200_host_placement(C.c_void_p(12345), cnt, 4) == -1


def _deliver(ps, lmin, lmax, a, b, nsub, nout, mode, scale=1, chunk_kb=64, nch=3, nthreads=2, pad=0):
    N = lmax - lmin + 1
    ld = N + pad
    outs = [np.asfortranarray(np.full((ld, N), np.nan)) for _ in range(nout)]   # column-major, leading dimension ld >= N
    DP = ps._lib.DP
    arr = (DP * nout)(*[o.ctypes.data_as(DP) for o in outs])
    rc = ps.lib().psb200_selftest_delivery(lmin, lmax, a, b, nsub, nout, mode, scale, chunk_kb, nch, nthreads, arr, ld)
    assert rc == 0, ps.lib().psb200_last_error()
    return outs


def _expected(lmin, lmax, a, b, o, scale):
    """What a band [a, b) owes the caller: column l1 below the diagonal holds fac(l1) x(l1, l2), row l1 right of it
    fac(l2) x(l1, l2), with the raw values x the test hook puts into its stand-in slabs."""
    N = lmax - lmin + 1
    E = np.full((N, N), np.nan)
    l2 = np.arange(lmin, lmax + 1, dtype=np.float64)
    for l1 in range(a, b):
        x = (o + 1.0) + 1e-3 * l1 + 1e-7 * l2[l1 - lmin:]
        E[l1 - lmin:, l1 - lmin] = (2 * l1 + 1) * x if scale else x
        E[l1 - lmin, l1 - lmin:] = (2 * l2[l1 - lmin:] + 1) * x if scale else x
    return E


@pytest.mark.parametrize("scale", [0, 1])
@pytest.mark.parametrize("lmin,lmax,a,b,nsub,nout", [(0, 255, 0, 256, 4, 1), (0, 255, 0, 256, 1, 2), (2, 300, 2, 301, 16, 5),
                                                     (0, 511, 100, 380, 8, 2), (0, 511, 380, 512, 2, 1), (5, 40, 7, 8, 1, 1)])
def test_staged_and_mirror_delivery_equal_direct(ps, lmin, lmax, a, b, nsub, nout, scale):
    """The direct delivery of a band writes its L-shaped region -- rows >= a of its columns, its rows of the columns to the
    right -- completely, only, and with the values of src/modecoupling.jl:90-91 / src/covariance.jl:119; the staged
    delivery (what pageable result arrays get) and the mirror delivery (block columns only over PCIe, the symmetric side
    written by the scatter workers) leave the same bytes for any chunk size / ring length / worker count."""
    N = lmax - lmin + 1
    direct = _deliver(ps, lmin, lmax, a, b, nsub, nout, 0, scale, pad=3)
    for o, D in enumerate(direct):
        assert np.isnan(D[N:, :]).all()           # the padding rows of the leading dimension stay untouched
        assert np.array_equal(D[:N, :], _expected(lmin, lmax, a, b, o, scale), equal_nan=True)
    for chunk_kb, nch, nthreads in [(4, 1, 1), (4, 2, 3), (16, 3, 2), (64, 12, 8), (1024, 2, 4)]:
        if chunk_kb * 1024 < N * 8:
            continue
        for mode in (1, 2) + ((3,) if scale == 0 else ()):
            got = _deliver(ps, lmin, lmax, a, b, nsub, nout, mode, scale, chunk_kb, nch, nthreads, pad=3)
            for S, D in zip(got, direct):
                assert np.array_equal(S, D, equal_nan=True), (mode, chunk_kb, nch, nthreads)


def test_staged_delivery_argument_errors(ps):
    L = ps.lib()
    out = np.zeros((8, 8), order="F")
    arr = (ps._lib.DP * 1)(out.ctypes.data_as(ps._lib.DP))
    assert L.psb200_selftest_delivery(0, 7, 0, 8, 1, 1, 1, 1, 64, 2, 2, arr, 4) == 1        # ld < N
    assert L.psb200_selftest_delivery(0, 7, 0, 9, 1, 1, 1, 1, 64, 2, 2, arr, 8) == 1        # band beyond the matrix
    assert L.psb200_selftest_delivery(0, 7, 0, 8, 1, 1, 1, 1, 64, 40, 2, arr, 8) == 1       # ring too long
    assert L.psb200_selftest_delivery(0, 7, 0, 8, 1, 1, 4, 1, 64, 2, 2, arr, 8) == 1        # unknown mode
    assert L.psb200_selftest_delivery(0, 7, 0, 8, 1, 1, 3, 1, 64, 2, 2, arr, 8) == 1        # direct mirror is for symmetric copies


def test_staged_delivery_random_shapes(ps):
    """Seeded random bands, sub-band counts, leading dimensions, chunk sizes, ring lengths and worker counts."""
    rng = np.random.default_rng(11)
    for _ in range(40):
        lmax = int(rng.integers(30, 400))
        lmin = int(rng.integers(0, 3))
        N = lmax - lmin + 1
        a = int(rng.integers(lmin, lmax))
        b = int(rng.integers(a + 1, lmax + 2))
        nsub, nout, pad = int(rng.integers(1, 17)), int(rng.integers(1, 6)), int(rng.integers(0, 5))
        scale = int(rng.integers(0, 2))
        chunk_kb = int(rng.integers((N * 8 + 1023) // 1024, 120))
        nch, nthreads = int(rng.integers(1, 33)), int(rng.integers(1, 12))
        direct = _deliver(ps, lmin, lmax, a, b, nsub, nout, 0, scale, pad=pad)
        for mode in (1, 2) + ((3,) if scale == 0 else ()):
            got = _deliver(ps, lmin, lmax, a, b, nsub, nout, mode, scale, chunk_kb, nch, nthreads, pad=pad)
            for S, D in zip(got, direct):
                assert np.array_equal(S, D, equal_nan=True), (mode, lmin, lmax, a, b, nsub, nout, pad, scale, chunk_kb, nch, nthreads)
