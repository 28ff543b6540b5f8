// psb200_lowrows.cuh -- rows l1 < 2 of the spin-2 jobs (the reference's "don't-care" region).
//
// For l1 < 2 the symbol (l3 l1 l2; 0 -2 2) is identically zero (|m2| > l1), and no reference test pins
// those entries.  But `master(...; lmin = 0)` is the reference default
// (/root/reference/src/modecoupling.jl:319,341) and hands rows/columns 0 and 1 to the decoupling solve, so
// they must hold what the reference's family routine produces when it is asked anyway, not zeros: the
// WignerFamilies.jl evaluation runs its recurrences on the (at most three) values j = |l1-l2| .. l1+l2
// without checking |m| <= l and normalises them to sum (2j+1) f^2 = 1.  The tuned kernel's closed-form start
// value and its even-parity identity are both 0/0 there, so these <= 2 rows (<= 2 (lmax+1) pairs, three
// terms each) are evaluated by this separate kernel, which follows the published two-sided scheme step by
// step (Luscombe & Luban 1998: ratio recursions inward from both ends, three-term recurrence across the
// middle, least-squares match, sum normalisation, sgn f(jmax) = (-1)^(l1-l2)); SURVEY.md 8(c) "don't-care
// region".  One thread per pair; the cost is nothing.
#pragma once
#include "psb200_common.cuh"

namespace psb {

__host__ __device__ constexpr bool job_has_spin2(int job) { return job_family(job) != FAM_00; }

// Reduced coefficients for m1 = 0:  X(j) f(j+1) + Y(j) f(j) + Z(j) f(j-1) = 0,
//   X(j) = a(j+1), Z(j) = a(j), a(j)^2 = (j^2-d^2)(s^2-j^2), Y(j) = (2j+1)(m3-m2).
struct LowFam {
    double d2, s2, y;           // (l1-l2)^2, (l1+l2+1)^2, m3-m2
    __device__ double X(int j) const { const double p = (double)(j + 1); return sqrt((p * p - d2) * (s2 - p * p)); }
    __device__ double Z(int j) const { const double p = (double)j; const double a2 = (p * p - d2) * (s2 - p * p); return a2 > 0.0 ? sqrt(a2) : 0.0; }
    __device__ double Y(int j) const { return (double)(2 * j + 1) * y; }
};

// Family of length n <= 3 on j = nmin .. nmin+n-1 into psi[].  spin2 = false: (0,0,0); true: (0,-2,2).
__device__ inline void low_family(int l1, int l2, bool spin2, double* psi, int nmin, int n)
{
    LowFam w;
    w.d2 = (double)(l1 - l2) * (double)(l1 - l2);
    w.s2 = (double)(l1 + l2 + 1) * (double)(l1 + l2 + 1);
    w.y = spin2 ? 4.0 : 0.0;
    const int nmax = nmin + n - 1;
#define PSI(j) psi[(j) - nmin]
    if (n == 1) {
        PSI(nmin) = 1.0;
    } else if (!spin2) {
        PSI(nmin) = 1.0;
        PSI(nmin + 1) = 0.0;
        for (int j = nmin + 1; j < nmax; ++j) PSI(j + 1) = -(w.Z(j) / w.X(j)) * PSI(j - 1);
    } else {
        // ratios r(j) = f(j)/f(j-1) downward from nmax while |r| < 1
        int nplus = nmax;
        {
            int j = nmax;
            double r = -w.Z(j) / w.Y(j);
            PSI(j) = r;
            while (fabs(r) < 1.0 && j - 1 > nmin) {
                --j;
                r = -w.Z(j) / (w.Y(j) + w.X(j) * r);
                PSI(j) = r;
            }
            nplus = j;
        }
        // ratios s(j) = f(j)/f(j+1) upward from nmin while |s| < 1
        int nminus = nmin;
        {
            int j = nmin;
            double s = -w.X(j) / w.Y(j);
            PSI(j) = s;
            while (fabs(s) < 1.0 && j + 1 < nplus) {
                ++j;
                s = -w.X(j) / (w.Y(j) + w.Z(j) * s);
                PSI(j) = s;
            }
            nminus = j;
        }
        const int nc = (nminus + nplus) / 2;
        const double s_at = PSI(nminus), r_at = PSI(nplus);
        PSI(nminus) = 1.0;
        for (int j = nminus - 1; j >= nmin; --j) PSI(j) = PSI(j) * PSI(j + 1);
        double Lc, Lc1;
        {
            double fm = 1.0, f0 = 1.0 / s_at;
            int j = nminus + 1;
            while (j <= nc) {
                PSI(j) = f0;
                const double fp = -(w.Y(j) * f0 + w.Z(j) * fm) / w.X(j);
                fm = f0; f0 = fp; ++j;
            }
            Lc = fm; Lc1 = f0;
        }
        PSI(nplus) = 1.0;
        for (int j = nplus + 1; j <= nmax; ++j) PSI(j) = PSI(j) * PSI(j - 1);
        double Uc, Uc1;
        {
            double gp = 1.0, g0 = 1.0 / r_at;
            int j = nplus - 1;
            while (j >= nc + 1) {
                PSI(j) = g0;
                const double gm = -(w.X(j) * gp + w.Y(j) * g0) / w.Z(j);
                gp = g0; g0 = gm; --j;
            }
            Uc = g0; Uc1 = gp;
        }
        const double lam = (Uc * Lc + Uc1 * Lc1) / (Lc * Lc + Lc1 * Lc1);
        for (int j = nmin; j <= nc; ++j) PSI(j) *= lam;
    }
    double norm = 0.0;
    for (int j = nmin; j <= nmax; ++j) norm += (double)(2 * j + 1) * PSI(j) * PSI(j);
    double sc = 1.0 / sqrt(norm);
    const bool neg = ((l1 - l2) & 1) != 0;
    if ((PSI(nmax) < 0.0) != neg) sc = -sc;
    for (int j = nmin; j <= nmax; ++j) PSI(j) *= sc;
#undef PSI
}

// One thread per pair (l1, l2), l1 in [row_lo, min(row_hi, 2)), l2 = l1 .. lmax.
template <int JOB>
__global__ void __launch_bounds__(128) low_rows_kernel(const PairArgs A)
{
    constexpr int NACC = job_nacc(JOB);
    const int l1 = A.row_lo + (int)blockIdx.y;
    const int l2 = l1 + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (l1 >= A.row_hi || l1 >= 2 || l2 > A.lmax) return;
    const int nmin = l2 - l1, n = 2 * l1 + 1;
    double f0[3], f2[3];
    low_family(l1, l2, false, f0, nmin, n);
    low_family(l1, l2, true, f2, nmin, n);
    double x[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) x[q] = 0.0;
    for (int t = 0; t < n; ++t) {
        const int j = nmin + t;
        if (j >= A.lenW) break;
        const bool even = (t & 1) == 0;                    // l1 + l2 + j = 2 l1 + 2 d + t
        const double k = (double)(2 * j + 1);
        const double g00 = k * f0[t] * f0[t], g22 = k * f2[t] * f2[t], g02 = k * f0[t] * f2[t];
        if constexpr (JOB == JOB_M02 || JOB == JOB_TEEE) {
            if (even) {
#pragma unroll
                for (int q = 0; q < NACC; ++q) x[q] += g02 * __ldg(A.W[q] + j);
            }
        } else if constexpr (JOB == JOB_MPP || JOB == JOB_EEEE || JOB == JOB_TEEEP) {
            if (even) {
#pragma unroll
                for (int q = 0; q < NACC; ++q) x[q] += g22 * __ldg(A.W[q] + j);
            }
        } else if constexpr (JOB == JOB_MMM) {
            if (!even) x[0] += g22 * __ldg(A.W[0] + j);
        } else if constexpr (JOB == JOB_MPPMMM) {
            if (even) x[0] += g22 * __ldg(A.W[0] + j); else x[1] += g22 * __ldg(A.W[0] + j);
        } else if constexpr (JOB == JOB_TETE) {
            x[1] += g00 * __ldg(A.W[1] + j);
            if (even) {
                x[0] += g02 * __ldg(A.W[0] + j);
                x[2] += g02 * __ldg(A.W[2] + j);
                x[3] += g02 * __ldg(A.W[3] + j);
                x[4] += g02 * __ldg(A.W[4] + j);
            }
        } else if constexpr (JOB == JOB_MASTER) {
            x[0] += g00 * __ldg(A.W[0] + j);
            if (even) {
                x[1] += g02 * __ldg(A.W[1] + j);
                x[2] += g02 * __ldg(A.W[2] + j);
                x[3] += g22 * __ldg(A.W[3] + j);
            } else {
                x[4] += g22 * __ldg(A.W[3] + j);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NACC; ++q) x[q] *= INV_4PI;
    epilogue<JOB>(A, l1, l2, x);
}

}  // namespace psb
