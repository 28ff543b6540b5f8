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

// Coefficients of both families at index jx >= 1:  an_k = At_k(jx), ian_k = 1/At_k(jx).
// rjx = 1/jx is returned for the simple variant's 1/(j(j+1)); the table variant does not need it.
template <bool TAB>
PSB_HD void qp_coefs(int jx, double d2, double ss, const QpFam& F0, const QpFam& F1, const QpTabs& T,
                     double* an0, double* ian0, double* an1, double* ian1, double* rjx)
{
    const double dj = (double)jx, jj = dj * dj;
    const double t12 = (jj - d2) * (ss - jj);
    if constexpr (TAB) {
        const double r = PSB_RSQRT(t12), a = t12 * r;
        const QpD2 b0 = T.BB0[jx], b1 = T.BB1[jx];
        *an0 = a * b0.x; *ian0 = r * b0.y;
        *an1 = a * b1.x; *ian1 = r * b1.y;
        *rjx = 0.0;
    } else {
        const double rj = 1.0 / dj, r2 = rj * rj;
        const double P0 = t12 * ((jj - F0.mm) * r2), P1 = t12 * ((jj - F1.mm) * r2);
        // a family that is not live yet (jx <= |m1_k|) has P <= 0: keep its state at exactly 0
        const double r0 = P0 > 0.0 ? PSB_RSQRT(P0) : 0.0, r1 = P1 > 0.0 ? PSB_RSQRT(P1) : 0.0;
        *an0 = P0 * r0; *ian0 = r0;
        *an1 = P1 * r1; *ian1 = r1;
        *rjx = rj;
    }
}

// Yt_k(j) for both families; ij = 1/(j(j+1)) (simple variant, 0 at j = 0)
template <bool TAB>
PSB_HD void qp_y(int j, double w2, double ij, const QpFam& F0, const QpFam& F1, const QpTabs& T, double* y0, double* y1)
{
    if constexpr (TAB) {
        const double t = T.IJ2[j];
        *y0 = fma(F0.c0, t, F0.dm * w2);
        *y1 = fma(F1.c0, t, F1.dm * w2);
    } else {
        *y0 = -w2 * fma(F0.c0, ij, -F0.dm);
        *y1 = -w2 * fma(F1.c0, ij, -F1.dm);
    }
}

// sum_j (2j+1) U(j)^2 of both families swept backward over their whole ranges (U(nmax) = 1).  Only
// used for the degenerate pairs whose two families overlap in the single term j = nmax.
template <bool TAB>
PSB_HD void qp_norm_backward(const QpFam& F0, const QpFam& F1, const QpTabs& T, double d2, double ss, int nmax,
                             double* n0, double* n1)
{
    double g0 = 1.0, b0 = 0.0, g1 = 1.0, b1 = 0.0;
    *n0 = 0.0; *n1 = 0.0;
    double rj1 = 1.0 / (double)(nmax + 1);
    const int jend = F0.nmin < F1.nmin ? F0.nmin : F1.nmin;
    for (int j = nmax; j >= jend; --j) {
        const double w2 = (double)(2 * j + 1);
        if (j >= F0.nmin) *n0 = fma(w2 * g0, g0, *n0);
        if (j >= F1.nmin) *n1 = fma(w2 * g1, g1, *n1);
        if (j == jend) break;
        double an0, ian0, an1, ian1, rj, y0, y1;
        qp_coefs<TAB>(j, d2, ss, F0, F1, T, &an0, &ian0, &an1, &ian1, &rj);
        qp_y<TAB>(j, w2, rj * rj1, F0, F1, T, &y0, &y1);
        const double q0 = fma(y0, g0, b0), q1 = fma(y1, g1, b1);
        b0 = an0 * g0; g0 = -q0 * ian0;          // below its own nmin a family gets ian = 0 and is not summed
        b1 = an1 * g1; g1 = -q1 * ian1;
        rj1 = rj;
        if (fabs(g0) > QP_BIG) { g0 *= QP_SMALL; b0 *= QP_SMALL; *n0 *= QP_SMALL * QP_SMALL; }
        if (fabs(g1) > QP_BIG) { g1 *= QP_SMALL; b1 *= QP_SMALL; *n1 *= QP_SMALL * QP_SMALL; }
    }
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
    const int jW = nmax < lenW - 1 ? nmax : lenW - 1;
    if (nlo > nmax || jW < nlo) return 0.0;                            // no common term inside the window
    const double dl = (double)l * (double)(l + 1) - (double)lpp * (double)(lpp + 1);
    F0.mm = (double)m1a * (double)m1a; F0.c0 = (double)m1a * dl; F0.dm = (double)(s1 - nu1);
    F1.mm = (double)m1b * (double)m1b; F1.c0 = (double)m1b * dl; F1.dm = (double)(s2 - nu2);
    const double d2 = (double)d * (double)d, ss = (double)(nmax + 1) * (double)(nmax + 1);

    if (nlo == nmax) {
        // the families share only j = nmax, where U_k = 1: Xi = W[nmax] / sqrt(N_1 N_2)
        double n0, n1;
        qp_norm_backward<TAB>(F0, F1, T, d2, ss, nmax, &n0, &n1);
        return (W[nmax] / sqrt(n0)) / sqrt(n1);          // N_k can reach 1e200 each: never multiply them
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

    // ---------------- forward sweep: j = min(nmin) .. c ----------------
    // f = L(j), p = At(j) L(j-1), fp = L(j-1).  A family whose nmin = |m1_k| lies above the other's start
    // stays at exactly 0 until j reaches it (its coefficients are 0 there), then is set to 1.
    double f0 = 0.0, p0 = 0.0, fp0 = 0.0, nL0 = 0.0;
    double f1 = 0.0, p1 = 0.0, fp1 = 0.0, nL1 = 0.0;
    double sLL = 0.0;
    {
        int j = F0.nmin < F1.nmin ? F0.nmin : F1.nmin;
        double rj = (!TAB && j > 0) ? 1.0 / (double)j : 0.0;
        for (; j <= c; ++j) {
            if (j == F0.nmin) f0 = 1.0;
            if (j == F1.nmin) f1 = 1.0;
            const double w2 = (double)(2 * j + 1);
            nL0 = fma(w2 * f0, f0, nL0);
            nL1 = fma(w2 * f1, f1, nL1);
            if (j >= nlo && j <= jW) sLL = fma(W[j] * f0, f1, sLL);
            double an0, ian0, an1, ian1, rjp, y0, y1;
            qp_coefs<TAB>(j + 1, d2, ss, F0, F1, T, &an0, &ian0, &an1, &ian1, &rjp);
            qp_y<TAB>(j, w2, rj * rjp, F0, F1, T, &y0, &y1);
            const double q0 = fma(y0, f0, p0), q1 = fma(y1, f1, p1);
            fp0 = f0; p0 = an0 * f0; f0 = -q0 * ian0;
            fp1 = f1; p1 = an1 * f1; f1 = -q1 * ian1;
            rj = rjp;
            if (fabs(f0) > QP_BIG) { f0 *= QP_SMALL; p0 *= QP_SMALL; fp0 *= QP_SMALL; nL0 *= QP_SMALL * QP_SMALL; sLL *= QP_SMALL; }
            if (fabs(f1) > QP_BIG) { f1 *= QP_SMALL; p1 *= QP_SMALL; fp1 *= QP_SMALL; nL1 *= QP_SMALL * QP_SMALL; sLL *= QP_SMALL; }
        }
    }
    // now f_k = L_k(c+1), fp_k = L_k(c)

    // ---------------- backward sweep: j = nmax .. c+1 ----------------
    double g0 = 1.0, b0 = 0.0, gp0 = 0.0, nU0 = 0.0;      // g = U(j), b = At(j+1) U(j+1), gp = U(j+1)
    double g1 = 1.0, b1 = 0.0, gp1 = 0.0, nU1 = 0.0;
    double sUU = 0.0;
    {
        double rj1 = TAB ? 0.0 : 1.0 / (double)(nmax + 1);
        for (int j = nmax; j > c; --j) {
            const double w2 = (double)(2 * j + 1);
            nU0 = fma(w2 * g0, g0, nU0);
            nU1 = fma(w2 * g1, g1, nU1);
            if (j <= jW) sUU = fma(W[j] * g0, g1, sUU);
            double an0, ian0, an1, ian1, rj, y0, y1;
            qp_coefs<TAB>(j, d2, ss, F0, F1, T, &an0, &ian0, &an1, &ian1, &rj);
            qp_y<TAB>(j, w2, rj * rj1, F0, F1, T, &y0, &y1);
            const double q0 = fma(y0, g0, b0), q1 = fma(y1, g1, b1);
            gp0 = g0; b0 = an0 * g0; g0 = -q0 * ian0;
            gp1 = g1; b1 = an1 * g1; g1 = -q1 * ian1;
            rj1 = rj;
            if (fabs(g0) > QP_BIG) { g0 *= QP_SMALL; b0 *= QP_SMALL; gp0 *= QP_SMALL; nU0 *= QP_SMALL * QP_SMALL; sUU *= QP_SMALL; }
            if (fabs(g1) > QP_BIG) { g1 *= QP_SMALL; b1 *= QP_SMALL; gp1 *= QP_SMALL; nU1 *= QP_SMALL * QP_SMALL; sUU *= QP_SMALL; }
        }
    }
    // now g_k = U_k(c), gp_k = U_k(c+1)

    const double lam0 = (g0 * fp0 + gp0 * f0) / (fp0 * fp0 + f0 * f0);
    const double lam1 = (g1 * fp1 + gp1 * f1) / (fp1 * fp1 + f1 * f1);
    const double N0 = fma(lam0 * lam0, nL0, nU0), N1 = fma(lam1 * lam1, nL1, nU1);
    return (fma(lam0 * lam1, sLL, sUU) / sqrt(N0)) / sqrt(N1);
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
__global__ void __launch_bounds__(QP_THREADS) quickpol_kernel(const QpArgs A)
{
    const int l = A.col_hi - 1 - (int)blockIdx.x;
    const int nb = A.band_lo + A.band_hi + 1;
    const int r = (int)blockIdx.y * QP_THREADS + (int)threadIdx.x;     // band row: l'' = l + r - band_hi
    if (l < A.col_lo || l < 2 || r >= nb) return;
    const int lpp = l + r - A.band_hi;
    if (lpp < 2 || lpp > A.lmax) return;
    A.Xb[(long)r + (long)l * A.ldb] = quickpol_pair_t<TAB>(l, lpp, A.nu1, A.nu2, A.s1, A.s2, A.W, A.lenW, A.T);
}
#endif

}  // namespace psb
