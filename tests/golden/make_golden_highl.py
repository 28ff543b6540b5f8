"""Multiprecision known answers for mode-coupling matrix ENTRIES at BASELINE's full sizes (lmax 6143 and 12287).

Independent of oracle/ and of the CUDA path: mpmath at 50 digits, plain forward three-term recurrence in l3 from
l3 = |l1 - l2| (stable here: for the two families of this path, (0,0,0) and (0,-2,2), all but a few values at either end
of a family lie in the classical region -- checked below by repeating a sample of the families at 400 digits), the
normalisation sum (2 l3 + 1) f^2 = 1, the sign rule sgn f(l1 + l2) = (-1)^(l1 - l2), and then the four sums of
/root/reference/src/modecoupling.jl:3-66 over the window spectrum V taken as exact binary doubles:

    Xi_TT = sum_l3            (2 l3 + 1) f00^2    V[l3] / 4 pi        -> M00 = (2 l2 + 1) Xi_TT      (:78-95)
    Xi_TE = sum_{l1+l2+l3 even} (2 l3 + 1) f00 f22 V[l3] / 4 pi        -> M02                         (:99-119)
    Xi_EE = sum_{even}          (2 l3 + 1) f22^2   V[l3] / 4 pi        -> M++                         (:123-139)
    Xi_EB = sum_{odd}           (2 l3 + 1) f22^2   V[l3] / 4 pi        -> M--                         (:143-159)

with l3 from |l1 - l2| to min(l1 + l2, len(V) - 1).  Stored per entry: the four Xi, and the sums of |terms| (the
condition of each sum, for the condition-aware bound of tests/conftest.py).  The window spectra are the synthetic
cross-mask spectra the GPU tests use (powerspectra_jl_b200/synthetic.py, seeds 1001 x 1002), stored verbatim so that
the fixture does not depend on how numpy rounds on another machine.

The same families give known answers for all seven covariance blocks at lmax 6143 -- TTTT, EEEE, TETE (the benchmark
step), TTTE, TEEE_planck, TEEE, TTEE; /root/reference/src/covariance.jl:92-122, :153-183, :261-302, :208-235, :376-402,
:337-372, :422-446 -- over inputs built from exactly rounded
arithmetic only (tests/highl_inputs.py; the fixture stores their digest, not the vectors): 32 entries per block, with
the condition sum  S_abs = sum_k |coefficient_k| sum_l3 |term|.

    python tests/golden/make_golden_highl.py        # ~2 minutes on 8 cores; writes mcm_entries_mp.npz, cov_entries_mp.npz
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def family(mp, l1, l2, m2, m3):
    """(l3 l1 l2; 0 m2 m3) with m2 + m3 = 0, l3 = |l1-l2| .. l1+l2, as a list of mpf."""
    d, s = abs(l2 - l1), l1 + l2 + 1
    a = lambda j: mp.sqrt((mp.mpf(j) ** 2 - d * d) * (mp.mpf(s) ** 2 - mp.mpf(j) ** 2))
    f, fm = [mp.mpf(1)], mp.mpf(0)
    for j in range(d, l1 + l2):
        fn = -((2 * j + 1) * (m3 - m2) * f[-1] + a(j) * fm) / a(j + 1)
        fm = f[-1]
        f.append(fn)
    norm = mp.sqrt(sum((2 * (d + t) + 1) * v * v for t, v in enumerate(f)))
    sgn = 1 if (f[-1] > 0) == ((l1 - l2) % 2 == 0) else -1
    return [sgn * v / norm for v in f]


def entry(args):
    l1, l2, V, dps = args
    import mpmath as mp
    mp.mp.dps = dps
    d = abs(l2 - l1)
    f00, f22 = family(mp, l1, l2, 0, 0), family(mp, l1, l2, -2, 2)
    last = min(l1 + l2, len(V) - 1)
    xi = [mp.mpf(0)] * 4
    sa = [mp.mpf(0)] * 4
    for l3 in range(d, last + 1):
        t = l3 - d
        w = (2 * l3 + 1) * mp.mpf(float(V[l3]))
        even = (l1 + l2 + l3) % 2 == 0
        terms = (w * f00[t] ** 2, w * f00[t] * f22[t] if even else 0, w * f22[t] ** 2 if even else 0,
                 0 if even else w * f22[t] ** 2)
        for k in range(4):
            xi[k] += terms[k]
            sa[k] += abs(terms[k])
    fourpi = 4 * mp.pi
    return [float(x / fourpi) for x in xi], [float(x / fourpi) for x in sa]


def cov_entry(args):
    """TTTT, EEEE, TETE at one (l1, l2): value and condition sum of each, formulas as in the reference."""
    l1, l2, inputs, dps = args
    import mpmath as mp
    mp.mp.dps = dps
    d = abs(l2 - l1)
    f00, f22 = family(mp, l1, l2, 0, 0), family(mp, l1, l2, -2, 2)
    fourpi = 4 * mp.pi
    F = lambda x, l: mp.mpf(float(x[l]))

    def xi(W, w, parity):
        """(sum, sum of |terms|) of (2 l3 + 1) w(l3) W[l3] / 4 pi; parity None = every l3, 0 = l1+l2+l3 even."""
        s = a = mp.mpf(0)
        for l3 in range(d, min(l1 + l2, len(W) - 1) + 1):
            if parity is not None and (l1 + l2 + l3) % 2 != parity:
                continue
            t = (2 * l3 + 1) * w[l3 - d] * mp.mpf(float(W[l3]))
            s += t
            a += abs(t)
        return s / fourpi, a / fourpi

    w00 = [v * v for v in f00]
    w22 = [v * v for v in f22]
    w02 = [u * v for u, v in zip(f00, f22)]
    out = {}
    for name, w, par in (("TTTT", w00, None), ("EEEE", w22, 0)):
        (sp, rt, W) = inputs[name]
        ip, jq, iq, jp = sp
        r_ip, r_jq, r_iq, r_jp = rt
        X = [xi(Wk, w, par) for Wk in W]
        coef = [mp.sqrt(F(ip, l1) * F(ip, l2) * F(jq, l1) * F(jq, l2)),
                mp.sqrt(F(iq, l1) * F(iq, l2) * F(jp, l1) * F(jp, l2)),
                mp.sqrt(F(ip, l1) * F(ip, l2)) * F(r_jq, l1) * F(r_jq, l2),
                mp.sqrt(F(jq, l1) * F(jq, l2)) * F(r_ip, l1) * F(r_ip, l2),
                mp.sqrt(F(iq, l1) * F(iq, l2)) * F(r_jp, l1) * F(r_jp, l2),
                mp.sqrt(F(jp, l1) * F(jp, l2)) * F(r_iq, l1) * F(r_iq, l2),
                F(r_ip, l1) * F(r_jq, l1) * F(r_ip, l2) * F(r_jq, l2),
                F(r_iq, l1) * F(r_jp, l1) * F(r_iq, l2) * F(r_jp, l2)]
        out[name] = (float(sum(c * x[0] for c, x in zip(coef, X))), float(sum(abs(c) * x[1] for c, x in zip(coef, X))))
    (sp, rt, W) = inputs["TETE"]
    TTip, EEjq, TEiq, TEjp = sp
    rT, rP = rt
    X = [xi(W[0], w02, 0), xi(W[1], w00, None), xi(W[2], w02, 0), xi(W[3], w02, 0), xi(W[4], w02, 0)]
    coef = [mp.sqrt(F(TTip, l1) * F(TTip, l2) * F(EEjq, l1) * F(EEjq, l2)),
            (F(TEiq, l1) * F(TEjp, l2) + F(TEjp, l1) * F(TEiq, l2)) / 2,
            mp.sqrt(F(TTip, l1) * F(TTip, l2)) * F(rP, l1) * F(rP, l2),
            mp.sqrt(F(EEjq, l1) * F(EEjq, l2)) * F(rT, l1) * F(rT, l2),
            F(rT, l1) * F(rT, l2) * F(rP, l1) * F(rP, l2)]
    out["TETE"] = (float(sum(c * x[0] for c, x in zip(coef, X))), float(sum(abs(c) * x[1] for c, x in zip(coef, X))))

    # TTTE (:208-235, f00^2, every l3), TEEE_planck (:376-402, f22^2, even), TEEE (:337-372, f00 f22, even): one shape
    for name, w, par in (("TTTE", w00, None), ("TEEE_planck", w22, 0), ("TEEE", w02, 0)):
        (sp, rt, W) = inputs[name]
        a, b, c1, c2 = sp                       # TTTE: TTip TTjp TEiq TEjq;  TEEE*: EEjq EEjp TEip TEiq
        ra, rb = rt
        X = [xi(Wk, w, par) for Wk in W]
        if name == "TTTE":
            s1, s2 = F(c2, l1) + F(c2, l2), F(c1, l1) + F(c1, l2)        # (TEjq1 + TEjq2), (TEiq1 + TEiq2)
        else:
            s1, s2 = F(c1, l1) + F(c1, l2), F(c2, l1) + F(c2, l2)        # (TEip1 + TEip2), (TEiq1 + TEiq2)
        coef = [mp.sqrt(F(a, l1) * F(a, l2)) * s1 / 2, mp.sqrt(F(b, l1) * F(b, l2)) * s2 / 2,
                s1 * F(ra, l1) * F(ra, l2) / 2, s2 * F(rb, l1) * F(rb, l2) / 2]
        out[name] = (float(sum(c * x[0] for c, x in zip(coef, X))), float(sum(abs(c) * x[1] for c, x in zip(coef, X))))
    # TTEE (:422-446, f00^2, every l3)
    (sp, rt, W) = inputs["TTEE"]
    TEip, TEiq, TEjq, TEjp = sp
    X = [xi(W[0], w00, None), xi(W[1], w00, None)]
    coef = [(F(TEip, l1) * F(TEjq, l2) + F(TEjq, l1) * F(TEip, l2)) / 2, (F(TEiq, l1) * F(TEjp, l2) + F(TEjp, l1) * F(TEiq, l2)) / 2]
    out["TTEE"] = (float(sum(c * x[0] for c, x in zip(coef, X))), float(sum(abs(c) * x[1] for c, x in zip(coef, X))))
    return out


def general_family(mp, j2, j3, m2, m3):
    """(j j2 j3; -(m2+m3) m2 m3), j = max(|j2-j3|, |m2+m3|) .. j2+j3: Schulten-Gordon three-term recurrence upwards from
    the first j, normalised and signed as WignerFamilies does.  Returns (first j, values)."""
    m1 = -(m2 + m3)
    if m1 == 0:
        return abs(j2 - j3), family(mp, j2, j3, m2, m3)
    jmin, jmax = max(abs(j2 - j3), abs(m1)), j2 + j3
    A = lambda j: mp.sqrt((mp.mpf(j) ** 2 - (j2 - j3) ** 2) * ((j2 + j3 + 1) ** 2 - mp.mpf(j) ** 2) * (mp.mpf(j) ** 2 - m1 * m1))
    B = lambda j: -(2 * j + 1) * (j2 * (j2 + 1) * m1 - j3 * (j3 + 1) * m1 - j * (j + 1) * (m3 - m2))
    f = [mp.mpf(1)]
    if jmax > jmin:
        f.append(-B(jmin) * f[0] / (jmin * A(jmin + 1)))            # A(jmin) = 0
    for j in range(jmin + 1, jmax):
        f.append(-(B(j) * f[-1] + (j + 1) * A(j) * f[-2]) / (j * A(j + 1)))
    norm = mp.sqrt(sum((2 * (jmin + t) + 1) * v * v for t, v in enumerate(f)))
    sgn = 1 if (f[-1] > 0) == ((j2 - j3 - m1) % 2 == 0) else -1
    return jmin, [sgn * v / norm for v in f]


def quickpol_entry(args):
    """Xi[l'', l] of quickpolXi! (/root/reference/src/beam.jl:72-101, Xisum :17-28) and the sum of |terms|."""
    lpp, l, nu1, nu2, s1, s2, W, dps = args
    import mpmath as mp
    mp.mp.dps = dps
    if abs(s1) > l or abs(s2) > l or abs(nu1) > lpp or abs(nu2) > lpp:
        return 0.0, 0.0                   # a projection larger than its angular momentum: the symbol is 0
    a1, f1 = general_family(mp, l, lpp, -s1, -nu1)
    a2, f2 = general_family(mp, l, lpp, -s2, -nu2)
    lo, hi = max(a1, a2), min(l + lpp, len(W) - 1)
    s = a = mp.mpf(0)
    for lp in range(lo, hi + 1):
        t = mp.mpf(float(W[lp])) * f1[lp - a1] * f2[lp - a2]
        s += t
        a += abs(t)
    sgn = -1 if (s1 + s2 + nu1 + nu2) % 2 else 1
    return float(sgn * s), float(a)


def check_general_family_against_sympy():
    import mpmath as mp
    from sympy import N
    from sympy.physics.wigner import wigner_3j
    mp.mp.dps = 50
    for (j2, j3, m2, m3) in [(7, 9, -2, -3), (12, 12, 2, 2), (30, 25, -4, 1), (9, 40, 2, -12), (5, 5, 0, 0), (6, 8, -2, 2)]:
        jmin, f = general_family(mp, j2, j3, m2, m3)
        for t, v in enumerate(f):
            exact = N(wigner_3j(jmin + t, j2, j3, -(m2 + m3), m2, m3), 40)
            assert abs(float(v) - float(exact)) < 1e-15, (j2, j3, m2, m3, jmin + t, float(v), float(exact))


def pairs_for(lmax, n, rng):
    """l1 <= l2 pairs: near and far from the diagonal, lowest spin-2 rows, last rows, edges of 8-way bands."""
    P = {(2, 2), (2, lmax), (3, lmax - 1), (lmax, lmax), (lmax - 1, lmax), (lmax // 2, lmax // 2), (lmax // 2, lmax),
         (2, lmax // 2 + 1), (17, 18), (lmax // 3, 2 * lmax // 3 + 1)}
    while len(P) < n:
        l1 = int(rng.integers(2, lmax + 1))
        kind = rng.integers(0, 3)
        if kind == 0:
            l2 = min(lmax, l1 + int(rng.integers(0, 40)))           # near the diagonal
        elif kind == 1:
            l2 = int(rng.integers(l1, lmax + 1))                    # anywhere right of it
        else:
            l1 = int(rng.integers(2, 200))                          # short families against a far column
            l2 = int(rng.integers(lmax // 2, lmax + 1))
        P.add((l1, l2))
    return sorted(P)


def main():
    from powerspectra_jl_b200 import synthetic as syn
    rng = np.random.default_rng(20261017)
    out = {}
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        for lmax, n in ((6143, 96), (12287, 48)):
            V = np.ascontiguousarray(syn.mask_spectra(lmax, seeds=(1001, 1002))[(0, 1)])
            P = pairs_for(lmax, n, rng)
            res = pool.map(entry, [(l1, l2, V, 50) for l1, l2 in P], chunksize=2)
            xi = np.array([r[0] for r in res])
            sabs = np.array([r[1] for r in res])
            # the recurrence is run again at 400 digits for every eighth entry: the 50-digit values must not move
            chk = pool.map(entry, [(l1, l2, V, 400) for l1, l2 in P[::8]], chunksize=1)
            dev = max(abs(a - b) / max(abs(b), 1e-300) for r, q in zip(res[::8], chk) for a, b in zip(r[0], q[0]))
            assert dev < 1e-14, dev
            out[f"V_{lmax}"] = V
            out[f"pairs_{lmax}"] = np.array(P, dtype=np.int32)
            out[f"xi_{lmax}"] = xi          # columns: Xi_TT, Xi_TE, Xi_EE, Xi_EB
            out[f"sabs_{lmax}"] = sabs
            print(lmax, len(P), "entries; 50 vs 400 digits:", dev, flush=True)
    np.savez_compressed(os.path.join(HERE, "mcm_entries_mp.npz"), **out)

    sys.path.insert(0, os.path.dirname(HERE))
    import highl_inputs
    lmax = 6143
    inputs = highl_inputs.cov_inputs(lmax)
    P = pairs_for(lmax, 32, rng)
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(cov_entry, [(l1, l2, inputs, 50) for l1, l2 in P], chunksize=1)
        chk = pool.map(cov_entry, [(l1, l2, inputs, 400) for l1, l2 in P[::8]], chunksize=1)
    dev = max(abs(r[b][0] - q[b][0]) / abs(q[b][0]) for r, q in zip(res[::8], chk) for b in r)
    assert dev < 1e-14, dev
    cov = {"pairs": np.array(P, dtype=np.int32), "lmax": np.int32(lmax), "inputs_sha256": np.array(highl_inputs.digest(inputs))}
    for b in ("TTTT", "EEEE", "TETE", "TTTE", "TEEE_planck", "TEEE", "TTEE"):
        cov[b] = np.array([r[b][0] for r in res])
        cov[b + "_sabs"] = np.array([r[b][1] for r in res])
    np.savez_compressed(os.path.join(HERE, "cov_entries_mp.npz"), **cov)
    print("cov", len(P), "entries per block; 50 vs 400 digits:", dev, flush=True)

    # QuickPol Xi: lmax 6143, band +-128, window of lmax + 1 entries, four spin cases
    check_general_family_against_sympy()
    band = 128
    W = highl_inputs.quickpol_window(lmax)
    cases = [(2, -2, 2, 2), (0, 0, 0, 0), (1, 3, -2, 2), (-12, 5, 4, -3)]          # (nu1, nu2, s1, s2)
    ent = {(2, 2), (2, 2 + band), (lmax, lmax), (lmax, lmax - band), (lmax - band, lmax), (13, 14), (lmax // 2, lmax // 2 + 7)}
    while len(ent) < 24:
        lpp = int(rng.integers(2, lmax + 1))
        l = int(rng.integers(max(2, lpp - band), min(lmax, lpp + band) + 1))
        ent.add((lpp, l))
    ent = sorted(ent)
    jobs = [(lpp, l, *c, W, 60) for c in cases for (lpp, l) in ent]
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(quickpol_entry, jobs, chunksize=1)
        chk = pool.map(quickpol_entry, [(*j[:7], 400) for j in jobs[::6]], chunksize=1)
    dev = max(abs(r[0] - q[0]) / max(abs(q[0]), 1e-300) for r, q in zip(res[::6], chk))
    assert dev < 1e-14, dev
    np.savez_compressed(os.path.join(HERE, "quickpol_entries_mp.npz"), lmax=np.int32(lmax), band=np.int32(band),
                        cases=np.array(cases, dtype=np.int32), entries=np.array(ent, dtype=np.int32),
                        xi=np.array([r[0] for r in res]).reshape(len(cases), len(ent)),
                        sabs=np.array([r[1] for r in res]).reshape(len(cases), len(ent)),
                        window_sha256=np.array(highl_inputs.digest({"W": ([W],)})))
    print("quickpol", len(cases), "x", len(ent), "entries; 60 vs 400 digits:", dev, flush=True)


if __name__ == "__main__":
    main()
