// psb200_quickpol.cuh -- QuickPol Xi matrix: general-spin Wigner-3j families on sm_100a.
//
// Replaces the pair loop of quickpolXi! (/root/reference/src/beam.jl:72-101) together with Xisum
// (:16-28) and the WignerF / wigner3j_f! calls it makes (:86-93):
//
//   Xi[l'', l] = (-1)^(s1+s2+nu1+nu2) sum_{l'} W[l'] (l' l l''; s1+nu1, -s1, -nu1) (l' l l''; s2+nu2, -s2, -nu2)
//
// for l'' = 2..lmax and l inside the stored band of row l'' (specrowrange, :59-63).  Unlike the (0,0,0) and
// (0,-2,2) families of the MCM / covariance kernels, these families have m1 = s+nu != 0 and real
// non-classical regions at both ends, so this kernel does what the north star describes: the three-term
// Schulten-Gordon recurrence swept FORWARD from j_min and BACKWARD from j_max (each in its stable
// direction), both branches matched inside the classical region, with rescaling against overflow.
//
// One thread owns one (l'', l) pair and runs BOTH families in lockstep over l' so the product
// W[l'] f1(l') f2(l') is accumulated on the fly; no 3j value is ever stored:
//   forward   j = min(nmin1,nmin2) .. c     L_k(j), sums  nL_k = sum (2j+1) L_k^2,  S_LL = sum W L_1 L_2
//   backward  j = nmax .. c+1               U_k(j), sums  nU_k = sum (2j+1) U_k^2,  S_UU = sum W U_1 U_2
//   match     lam_k = least-squares scale of L_k onto U_k over {c, c+1}
//   result    Xi = (lam_1 lam_2 S_LL + S_UU) / sqrt(N_1 N_2),   N_k = lam_k^2 nL_k + nU_k
// (the sign (-1)^(s1+s2+nu1+nu2) cancels against the two sign conventions f_k(nmax) ~ (-1)^(l-l''-m1_k),
// because U_k(nmax) = +1 for both).  c is the middle of the intersection of the two classical regions
//   j(j+1) in [ m1^2 + (lam2-lam3)^2 , m1^2 + (lam2+lam3)^2 ],  lam2^2 = l(l+1)-s^2, lam3^2 = l''(l''+1)-nu^2.
//
// Recurrence, divided by j(j+1) so that it is also valid at j = 0:
//   At(j+1) f(j+1) + Yt(j) f(j) + At(j) f(j-1) = 0
//   At(j)^2 = (j^2-d^2)(S^2-j^2)(j^2-m1^2)/j^2,   d = |l-l''|, S = l+l''+1
//   Yt(j)   = -(2j+1) [ m1 (l(l+1) - l''(l''+1)) / (j(j+1)) - (s - nu) ]
// One rsqrt per family and one reciprocal per step: At = P rsqrt(P), 1/At = rsqrt(P).
//
// The pair function is __host__ __device__ and free of CUDA-only constructs so tests/ can compile it
// with g++ (tests/hostcheck) and compare it with the oracle without a GPU.  That build is test
// infrastructure; libpsb200.so only ever runs the __global__ kernel below.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define PSB_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define PSB_HD static inline
#endif

namespace psb {

#if defined(__CUDA_ARCH__)
#define PSB_RSQRT(x) rsqrt(x)
#else
#define PSB_RSQRT(x) (1.0 / sqrt(x))
#endif

#ifndef PSB200_QP_UNROLL
#define PSB200_QP_UNROLL 4              // steps between two overflow tests (see qp_rescale for the bound)
#endif
#ifndef PSB200_QP_FLAT
#define PSB200_QP_FLAT 1                // 1: flattened (column, row) thread mapping (default: 21.8 vs 27.0 ms at lmax 6143, band +-128); 0: one block row per column
#endif
#ifndef PSB200_QP_MINBLOCKS
#define PSB200_QP_MINBLOCKS 1           // resident 64-thread blocks per SM the register allocation must allow
#endif
constexpr double QP_BIG = 1e100;        // rescale threshold
constexpr double QP_SMALL = 1e-100;

struct QpFam {
    double mm;      // m1^2
    double c0;      // m1 (l(l+1) - l''(l''+1))
    double dm;      // m3 - m2 = s - nu
    int nmin;
};

// Pair-independent factors of the coefficients, tabulated once per call over j = 0 .. 2 lmax+1
// (variant TAB): At(j) = a(j) B_k(j) with a(j)^2 = (j^2-d^2)(S^2-j^2) the only per-pair part, so ONE
// rsqrt per step serves both families and no reciprocal is left in the loop:
//   BBk[j] = (B_k, 1/B_k),  B_k(j) = sqrt(j^2 - m1_k^2) / j   ((0,0) for j <= |m1_k| and j = 0)
//   IJ2[j] = -(2j+1) / (j(j+1))   (0 at j = 0),   Yt_k(j) = c0_k IJ2[j] + (2j+1) dm_k
struct alignas(16) QpD2 { double x, y; };
struct QpTabs {
    const double* IJ2;
    const QpD2* BB0;
    const QpD2* BB1;
};

PSB_HD void qp_tab_entry(int j, int m1a, int m1b, double* ij2, QpD2* bb0, QpD2* bb1)
{
    const double dj = (double)j;
    *ij2 = j > 0 ? -(double)(2 * j + 1) / (dj * (double)(j + 1)) : 0.0;
    const int a0 = m1a < 0 ? -m1a : m1a, a1 = m1b < 0 ? -m1b : m1b;
    QpD2 z; z.x = 0.0; z.y = 0.0;
    *bb0 = z; *bb1 = z;
    if (j > a0 && j > 0) { const double b = sqrt((double)(j - a0) * (double)(j + a0)) / dj; bb0->x = b; bb0->y = 1.0 / b; }
    if (j > a1 && j > 0) { const double b = sqrt((double)(j - a1) * (double)(j + a1)) / dj; bb1->x = b; bb1->y = 1.0 / b; }
}

// classical region of family (s, nu) at (l, l''), in j
PSB_HD void qp_classical(int l, int lpp, int s, int nu, double* jlo, double* jhi)
{
    const double m1 = (double)(s + nu);
    const double a2 = (double)l * (double)(l + 1) - (double)s * (double)s;
    const double a3 = (double)lpp * (double)(lpp + 1) - (double)nu * (double)nu;
    const double l2 = sqrt(a2 > 0.0 ? a2 : 0.0), l3 = sqrt(a3 > 0.0 ? a3 : 0.0);
    const double xlo = m1 * m1 + (l2 - l3) * (l2 - l3), xhi = m1 * m1 + (l2 + l3) * (l2 + l3);
    *jlo = 0.5 * (sqrt(1.0 + 4.0 * xlo) - 1.0);
    *jhi = 0.5 * (sqrt(1.0 + 4.0 * xhi) - 1.0);
}

// State of one sweep (both families): v = L(j) or U(j), p = At * (previous value), vp = previous value,
// n = sum (2j+1) v^2, s = sum W v_0 v_1.  All in registers after inlining.
struct QpSweep {
    double v0, p0, vp0, n0;
    double v1, p1, vp1, n1;
    double s;
};

// One step of a sweep at index j.  The norm and W sums take the values AT j; then both families move to
// j+1 (forward: jx = j+1) or j-1 (backward: jx = j) with
//     q = Yt(j) v + p,   p' = At(jx) v,   v' = -q / At(jx).
// dj = (double)jx, w2 = 2j+1 are carried as doubles (no int->double conversions in the loop).
// The simple variant also carries rprev = 1/(the other index of the pair {j, j+1}) for 1/(j(j+1)).
// A family that is not live yet (forward sweep, jx <= |m1_k|) has At = 1/At = 0 and stays at exactly 0.
// pw = &W[j], pij = &IJ2[j], pb0/pb1 = &BB0[jx], &BB1[jx]: the sweeps carry pointers so that the unrolled steps
// load with immediate offsets (no per-load address arithmetic).
template <bool TAB, bool WSUM>
PSB_HD void qp_step(QpSweep& S, const double* __restrict__ pw, const double* __restrict__ pij,
                    const QpD2* __restrict__ pb0, const QpD2* __restrict__ pb1, double dj, double w2, double& rprev,
                    double d2, double ss, const QpFam& F0, const QpFam& F1)
{
    S.n0 = fma(w2 * S.v0, S.v0, S.n0);
    S.n1 = fma(w2 * S.v1, S.v1, S.n1);
    if constexpr (WSUM) S.s = fma(*pw * S.v0, S.v1, S.s);
    const double jj = dj * dj;
    const double t12 = (jj - d2) * (ss - jj);
    double an0, ian0, an1, ian1, y0, y1;
    if constexpr (TAB) {
        const double r = PSB_RSQRT(t12), a = t12 * r;
        const QpD2 b0 = *pb0, b1 = *pb1;
        const double t = *pij;
        an0 = a * b0.x; ian0 = r * b0.y;
        an1 = a * b1.x; ian1 = r * b1.y;
        y0 = fma(F0.c0, t, F0.dm * w2);
        y1 = fma(F1.c0, t, F1.dm * w2);
    } else {
        const double rj = 1.0 / dj, r2 = rj * rj, ij = rj * rprev;
        const double P0 = t12 * ((jj - F0.mm) * r2), P1 = t12 * ((jj - F1.mm) * r2);
        const double r0 = P0 > 0.0 ? PSB_RSQRT(P0) : 0.0, r1 = P1 > 0.0 ? PSB_RSQRT(P1) : 0.0;
        an0 = P0 * r0; ian0 = r0;
        an1 = P1 * r1; ian1 = r1;
        y0 = -w2 * fma(F0.c0, ij, -F0.dm);
        y1 = -w2 * fma(F1.c0, ij, -F1.dm);
        rprev = rj;
    }
    const double q0 = fma(y0, S.v0, S.p0), q1 = fma(y1, S.v1, S.p1);
    S.vp0 = S.v0; S.p0 = an0 * S.v0; S.v0 = -q0 * ian0;
    S.vp1 = S.v1; S.p1 = an1 * S.v1; S.v1 = -q1 * ian1;
}

// Rescaling against overflow.  One step multiplies a value by at most |Yt| / At < (2j+1)(S + |s-nu|) sqrt(j) ~ 1e11
// for lmax <= 12287 (a crude bound: real growth in the non-classical regions is a small factor per step), so with a
// test against 1e100 every FOURTH step values stay below 1e145 and the norm terms (2j+1) v^2 below 1e295 -- inside
// the double range.  Eight steps between tests would not be provably safe.  The branch is almost never taken.
PSB_HD void qp_rescale(QpSweep& S)
{
    const double a0 = fabs(S.v0), a1 = fabs(S.v1);
    if ((a0 > a1 ? a0 : a1) > QP_BIG) {
        const double k0 = a0 > QP_BIG ? QP_SMALL : 1.0, k1 = a1 > QP_BIG ? QP_SMALL : 1.0;
        S.v0 *= k0; S.p0 *= k0; S.vp0 *= k0; S.n0 *= k0 * k0;
        S.v1 *= k1; S.p1 *= k1; S.vp1 *= k1; S.n1 *= k1 * k1;
        S.s *= k0 * k1;
    }
}

// Steps j = ja, ja+DIR, ..., jb (inclusive; nothing if the range is empty).  DIR = +1 forward, -1 backward.
template <bool TAB, bool WSUM, int DIR>
PSB_HD void qp_run(QpSweep& S, int ja, int jb, double& rprev, double d2, double ss,
                   const QpFam& F0, const QpFam& F1, const QpTabs& T, const double* __restrict__ W)
{
    int n = DIR > 0 ? jb - ja + 1 : ja - jb + 1;
    if (n <= 0) return;
    constexpr int OFF = DIR > 0 ? 1 : 0;                  // jx = j + OFF
    double dj = (double)(ja + OFF), w2 = (double)(2 * ja + 1);
    if constexpr (!TAB) {
        // 1/(j(j+1)) needs the reciprocal of the index that is NOT jx: j (forward) or j+1 (backward)
        const int other = DIR > 0 ? ja : ja + 1;
        rprev = other > 0 ? 1.0 / (double)other : 0.0;
    }
    const double* pw = WSUM ? W + ja : W;
    const double* pij = TAB ? T.IJ2 + ja : nullptr;
    const QpD2* pb0 = TAB ? T.BB0 + (ja + OFF) : nullptr;
    const QpD2* pb1 = TAB ? T.BB1 + (ja + OFF) : nullptr;
    constexpr int UN = PSB200_QP_UNROLL;
    for (; n >= UN; n -= UN) {
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            qp_step<TAB, WSUM>(S, pw + u * DIR, pij + u * DIR, pb0 + u * DIR, pb1 + u * DIR, dj, w2, rprev, d2, ss, F0, F1);
            dj += (double)DIR; w2 += 2.0 * DIR;
        }
        if constexpr (WSUM) pw += UN * DIR;
        if constexpr (TAB) { pij += UN * DIR; pb0 += UN * DIR; pb1 += UN * DIR; }
        qp_rescale(S);
    }
    for (; n > 0; --n) {
        qp_step<TAB, WSUM>(S, pw, pij, pb0, pb1, dj, w2, rprev, d2, ss, F0, F1);
        dj += (double)DIR; w2 += 2.0 * DIR;
        if constexpr (WSUM) pw += DIR;
        if constexpr (TAB) { pij += DIR; pb0 += DIR; pb1 += DIR; }
    }
    qp_rescale(S);
}

// One entry of the Xi matrix.  W[0..lenW-1] is 0-based in l'; terms with l' > lenW-1 are dropped.
template <bool TAB>
PSB_HD double quickpol_pair_t(int l, int lpp, int nu1, int nu2, int s1, int s2, const double* __restrict__ W, int lenW,
                              const QpTabs& T)
{
    const int as1 = s1 < 0 ? -s1 : s1, as2 = s2 < 0 ? -s2 : s2, an1 = nu1 < 0 ? -nu1 : nu1, an2 = nu2 < 0 ? -nu2 : nu2;
    if (as1 > l || as2 > l || an1 > lpp || an2 > lpp) return 0.0;      // projection above the angular momentum
    const int d = l > lpp ? l - lpp : lpp - l, nmax = l + lpp;
    const int m1a = s1 + nu1, m1b = s2 + nu2;
    QpFam F0, F1;
    F0.nmin = (m1a < 0 ? -m1a : m1a) > d ? (m1a < 0 ? -m1a : m1a) : d;
    F1.nmin = (m1b < 0 ? -m1b : m1b) > d ? (m1b < 0 ? -m1b : m1b) : d;
    const int nlo = F0.nmin > F1.nmin ? F0.nmin : F1.nmin;
    const int jA = F0.nmin < F1.nmin ? F0.nmin : F1.nmin;
    const int jW = nmax < lenW - 1 ? nmax : lenW - 1;
    if (nlo > nmax || jW < nlo) return 0.0;                            // no common term inside the window
    const double dl = (double)l * (double)(l + 1) - (double)lpp * (double)(lpp + 1);
    F0.mm = (double)m1a * (double)m1a; F0.c0 = (double)m1a * dl; F0.dm = (double)(s1 - nu1);
    F1.mm = (double)m1b * (double)m1b; F1.c0 = (double)m1b * dl; F1.dm = (double)(s2 - nu2);
    const double d2 = (double)d * (double)d, ss = (double)(nmax + 1) * (double)(nmax + 1);
    double rprev = 0.0;

    if (nlo == nmax) {
        // The families share only j = nmax, where U_k = 1: Xi = W[nmax] / sqrt(N_1 N_2).  The longer family
        // is swept backward over its whole (short: <= |m1| terms) range for its norm; the other one has the
        // single term j = nmax, its coefficients below that are 0 and it adds nothing more to its norm.
        QpSweep B = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0};
        qp_run<TAB, false, -1>(B, nmax, jA + 1, rprev, d2, ss, F0, F1, T, W);
        B.n0 = fma((double)(2 * jA + 1) * B.v0, B.v0, B.n0);          // the last value, at j = jA
        B.n1 = fma((double)(2 * jA + 1) * B.v1, B.v1, B.n1);
        return (W[nmax] / sqrt(B.n0)) / sqrt(B.n1);      // N_k can reach 1e200 each: never multiply them
    }

    // matching point: middle of the intersection of the classical regions, nlo <= c <= nmax-1
    int c;
    {
        double lo0, hi0, lo1, hi1;
        qp_classical(l, lpp, s1, nu1, &lo0, &hi0);
        qp_classical(l, lpp, s2, nu2, &lo1, &hi1);
        const double mid = 0.5 * ((lo0 > lo1 ? lo0 : lo1) + (hi0 < hi1 ? hi0 : hi1));
        c = (int)mid;
        if (c < nlo) c = nlo;
        if (c > nmax - 1) c = nmax - 1;
    }
    const int cw = c < jW ? c : jW;                      // last forward step that still sees the window

    // ---------------- forward sweep: L_k(j), j = jA .. c  (values end at c+1) ----------------
    QpSweep L = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (F0.nmin == jA) L.v0 = 1.0;
    if (F1.nmin == jA) L.v1 = 1.0;
    qp_run<TAB, false, +1>(L, jA, nlo - 1, rprev, d2, ss, F0, F1, T, W);    // only the earlier family is live
    if (F0.nmin > jA) L.v0 = 1.0;                        // the later one starts at j = nlo
    if (F1.nmin > jA) L.v1 = 1.0;
    qp_run<TAB, true, +1>(L, nlo, cw, rprev, d2, ss, F0, F1, T, W);
    qp_run<TAB, false, +1>(L, cw + 1, c, rprev, d2, ss, F0, F1, T, W);
    // now L.v_k = L_k(c+1), L.vp_k = L_k(c)

    // ---------------- backward sweep: U_k(j), j = nmax .. c+1  (values end at c) ----------------
    QpSweep U = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0};
    const int uw = jW > c ? jW : c;                      // first backward step that sees the window
    qp_run<TAB, false, -1>(U, nmax, uw + 1, rprev, d2, ss, F0, F1, T, W);
    qp_run<TAB, true, -1>(U, uw, c + 1, rprev, d2, ss, F0, F1, T, W);
    // now U.v_k = U_k(c), U.vp_k = U_k(c+1)

    const double lam0 = (U.v0 * L.vp0 + U.vp0 * L.v0) / (L.vp0 * L.vp0 + L.v0 * L.v0);
    const double lam1 = (U.v1 * L.vp1 + U.vp1 * L.v1) / (L.vp1 * L.vp1 + L.v1 * L.v1);
    const double N0 = fma(lam0 * lam0, L.n0, U.n0), N1 = fma(lam1 * lam1, L.n1, U.n1);
    return (fma(lam0 * lam1, L.s, U.s) / sqrt(N0)) / sqrt(N1);
}

#ifdef __CUDACC__
// Band storage of BandedMatrices (parent(Xi)): Xb[(band_hi + l'' - l) + l*ldb], one column per l.
// grid.x = columns of this launch, heaviest (largest l) first; threads run over the band rows of a
// column, so a warp holds 32 consecutive l'' of one l (family lengths within a warp differ by <= 64)
// and its stores are contiguous.
struct QpArgs {
    int nu1, nu2, s1, s2;
    int lmax, lenW;
    int band_lo, band_hi;
    int col_lo, col_hi;        // columns l in [col_lo, col_hi)
    long ldb;
    const double* W;
    double* Xb;                // points at column 0 of the band storage
    QpTabs T;                  // table variant only (else null)
};

constexpr int QP_THREADS = 64;

// tables of the TAB variant for j = 0 .. n-1
__global__ void quickpol_tables_kernel(double* __restrict__ ij2, QpD2* __restrict__ bb0, QpD2* __restrict__ bb1, int n,
                                       int m1a, int m1b)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) qp_tab_entry(j, m1a, m1b, ij2 + j, bb0 + j, bb1 + j);
}

template <bool TAB>
__global__ void __launch_bounds__(QP_THREADS, PSB200_QP_MINBLOCKS) quickpol_kernel(const QpArgs A)
{
    const int nb = A.band_lo + A.band_hi + 1;
#if PSB200_QP_FLAT
    // (column, band row) flattened into one index so that no warp is mostly empty -- with 64-thread blocks laid over the
    // 2*128+1 band rows of one column 11 % of the lanes idle and the last block of every column is one warp short
    const long idx = (long)blockIdx.x * QP_THREADS + (long)threadIdx.x;
    const long cidx = idx / nb;
    if (cidx >= (long)(A.col_hi - A.col_lo)) return;
    const int l = A.col_hi - 1 - (int)cidx;
    const int r = (int)(idx - cidx * nb);
    if (l < 2) return;
#else
    const int l = A.col_hi - 1 - (int)blockIdx.x;
    const int r = (int)blockIdx.y * QP_THREADS + (int)threadIdx.x;     // band row: l'' = l + r - band_hi
    if (l < A.col_lo || l < 2 || r >= nb) return;
#endif
    const int lpp = l + r - A.band_hi;
    if (lpp < 2 || lpp > A.lmax) return;
    A.Xb[(long)r + (long)l * A.ldb] = quickpol_pair_t<TAB>(l, lpp, A.nu1, A.nu2, A.s1, A.s2, A.W, A.lenW, A.T);
}
#endif

}  // namespace psb
