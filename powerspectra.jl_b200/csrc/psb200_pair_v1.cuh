// psb200_pair_v1.cuh -- straightforward pair kernel: one thread owns one (l1,l2) pair and runs
// the forward Schulten-Gordon recurrence over the full family with sqrt / divide evaluated
// inline, normalising with sum (2 l3+1) f^2 = 1 exactly as the reference's dependency does.
// It is the simple on-device cross-check of the tuned kernel (psb200_pair_v2.cuh) and is
// selected with PSB200_KERNEL=v1; it is NOT a CPU fallback -- it needs the same GPU.
//
// Reduced recurrence for m1 = 0 (both families of this path):
//     a(j+1) f(j+1) + (2j+1)(m3-m2) f(j) + a(j) f(j-1) = 0,   a(j)^2 = (j^2-d^2)(s^2-j^2),
// d = l2-l1, s = l1+l2+1.  (0,0,0): the middle term vanishes, odd-parity terms are 0.
#pragma once
#include "psb200_common.cuh"

namespace psb {

constexpr int V1_THREADS = 128;

template <int JOB>
__global__ void __launch_bounds__(V1_THREADS) pair_kernel_v1(const PairArgs A)
{
    constexpr int FAM = job_family(JOB);
    constexpr int NW = job_nw(JOB);
    constexpr int NACC = job_nacc(JOB);

    const int l1 = A.row_hi - 1 - (int)blockIdx.y;            // heavy rows (large l1) first
    const int l2 = l1 + (int)blockIdx.x * V1_THREADS + (int)threadIdx.x;
    if (l2 > A.lmax) return;

    const int d = l2 - l1;
    const double dd = (double)d * (double)d;
    const double ss = (double)(l1 + l2 + 1) * (double)(l1 + l2 + 1);
    const int nt = 2 * l1 + 1;                                // family length, t = j - d

    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    double n00 = 0.0, n22 = 0.0;

    // state at step t (j = d + t): f = f22(j), fm = f22(j-1), h = f00(j) (0 on odd t), a = a(j)
    double f = 1.0, fm = 0.0, h = 1.0, hm = 0.0, a = 0.0;
    double flast = 1.0, hlast = 1.0;

    for (int t = 0; t < nt; ++t) {
        const int j = d + t;
        const double k = (double)(2 * j + 1);
        const bool even = (t & 1) == 0;
        const bool inW = j < A.lenW;

        // ---- accumulate this term ----
        if constexpr (FAM == FAM_00) {
            if (even) {
                const double g = h * h;
                n00 = fma(k, g, n00);
                if (inW) {
                    const double kg = k * g;
#pragma unroll
                    for (int q = 0; q < NW; ++q) acc[q] = fma(kg, __ldg(A.W[q] + j), acc[q]);
                }
            }
        } else if constexpr (FAM == FAM_22) {
            const double g = f * f;
            n22 = fma(k, g, n22);
            if (inW) {
                const double kg = k * g;
                if constexpr (JOB == JOB_MPPMMM) {
                    const double w = __ldg(A.W[0] + j);
                    if (even) acc[0] = fma(kg, w, acc[0]); else acc[1] = fma(kg, w, acc[1]);
                } else if constexpr (JOB == JOB_MMM) {
                    if (!even) acc[0] = fma(kg, __ldg(A.W[0] + j), acc[0]);
                } else {
                    if (even) {
#pragma unroll
                        for (int q = 0; q < NW; ++q) acc[q] = fma(kg, __ldg(A.W[q] + j), acc[q]);
                    }
                }
            }
        } else {  // FAM_02
            n22 = fma(k, f * f, n22);
            if constexpr (JOB == JOB_MASTER) {
                if (inW) {
                    const double kg = k * (f * f);
                    if (even) acc[3] = fma(kg, __ldg(A.W[3] + j), acc[3]);
                    else acc[4] = fma(kg, __ldg(A.W[3] + j), acc[4]);
                }
            }
            if (even) {
                n00 = fma(k, h * h, n00);
                if (inW) {
                    const double kp = k * (h * f);
                    if constexpr (JOB == JOB_MASTER) {
                        acc[0] = fma(k * (h * h), __ldg(A.W[0] + j), acc[0]);
                        acc[1] = fma(kp, __ldg(A.W[1] + j), acc[1]);
                        acc[2] = fma(kp, __ldg(A.W[2] + j), acc[2]);
                    } else if constexpr (JOB == JOB_TETE) {
                        acc[0] = fma(kp, __ldg(A.W[0] + j), acc[0]);
                        acc[1] = fma(k * (h * h), __ldg(A.W[1] + j), acc[1]);
                        acc[2] = fma(kp, __ldg(A.W[2] + j), acc[2]);
                        acc[3] = fma(kp, __ldg(A.W[3] + j), acc[3]);
                        acc[4] = fma(kp, __ldg(A.W[4] + j), acc[4]);
                    } else {
#pragma unroll
                        for (int q = 0; q < NW; ++q) acc[q] = fma(kp, __ldg(A.W[q] + j), acc[q]);
                    }
                }
            }
        }
        flast = f;
        if (even) hlast = h;

        // ---- advance j -> j+1 ----
        if (t + 1 < nt) {
            const double jp = (double)(j + 1);
            const double an = sqrt((jp * jp - dd) * (ss - jp * jp));   // a(j+1) > 0 for j+1 <= jmax
            if constexpr (FAM != FAM_00) {
                const double fn = -(4.0 * k * f + a * fm) / an;
                fm = f; f = fn;
            }
            if constexpr (FAM != FAM_22) {
                const double hn = -(a * hm) / an;                       // 0 on odd t+1
                hm = h; h = hn;
            }
            a = an;
        }
    }

    // ---- normalise, sign, scale by 1/4pi ----
    double x[NACC];
    if constexpr (FAM == FAM_00) {
        const double sc = INV_4PI / n00;
#pragma unroll
        for (int q = 0; q < NACC; ++q) x[q] = acc[q] * sc;
    } else if constexpr (FAM == FAM_22) {
        const double sc = INV_4PI / n22;
#pragma unroll
        for (int q = 0; q < NACC; ++q) x[q] = acc[q] * sc;
    } else {
        // sgn f00(jmax) = sgn f22(jmax) = (-1)^d after normalisation => the product's sign
        // correction is sgn(h_last) * sgn(f_last).
        double sc = INV_4PI / sqrt(n00 * n22);
        if ((hlast < 0.0) != (flast < 0.0)) sc = -sc;
#pragma unroll
        for (int q = 0; q < NACC; ++q) x[q] = acc[q] * sc;
        if constexpr (JOB == JOB_TETE) x[1] = acc[1] * (INV_4PI / n00);
        if constexpr (JOB == JOB_MASTER) {
            x[0] = acc[0] * (INV_4PI / n00);
            x[3] = acc[3] * (INV_4PI / n22);
            x[4] = acc[4] * (INV_4PI / n22);
        }
    }
    epilogue<JOB>(A, l1, l2, x);
}

}  // namespace psb
