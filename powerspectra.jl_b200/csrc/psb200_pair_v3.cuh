// psb200_pair_v3.cuh -- closed-form pair kernel for sm_100a (FP64-pipe bound by design); the default kernel.
//
// No recurrence is run at all.  With d = l2-l1, L = 2 l1+1, t = j-d, m = j+d and g(n) = binom(2n,n)/4^n
// (table `gam`), the squared (0,0,0) symbol of an even-parity term (l1+l2+j even <=> t even) factorises into a
// part that depends on (row, t) and a part that depends on (row, m):
//     f00(j)^2 = g(J-j) g(J-l1) g(J-l2) / (g(J) (2J+1)),  2J = j+l1+l2
//              = [ g(t/2) g(l1 - t/2) ] * [ g(m/2) / (g(m/2 + l1) (m + L)) ]  =  PT[t] * PV[m].
// Spin 2, l1+l2+j even (two steps of the m-ladder, as in psb200_pair_v2.cuh), x = j(j+1), a = l1(l1+1), b = l2(l2+1):
//     f22(j) = f00(j) N(x)/D,   2 N(x) = u^2 - (2ab+1),  u = x - (a+b-1),
//     D^2 = (l1-1) l1 (l1+1)(l1+2) (l2-1) l2 (l2+1)(l2+2).
// Spin 2, l1+l2+j odd: f00 vanishes; the m-ladder at m = 1 gives f22(j) = (x-a-b+2) f11(j)/D1 and the j-recurrence of
// the (0,-1,1) family between two even-parity neighbours collapses to f11(j) = a(j) f00(j-1) / (2 sqrt(ab)),
// a(j)^2 = (j^2-d^2)(s^2-j^2) = [t (L-t)] [m (m+L)].  Hence
//     4 D^2 f22(j)^2 = (x-a-b+2)^2 * QT[t] * QV[m],   QT[t] = t (L-t) PT[t-1],  QV[m] = m (m+L) PV[m-1]
// -- again one falling-index and one rising-index per-row table.  (All three identities are checked against exact
// sympy values and the long-double oracle: tests/test_oracle.py::test_closed_form_products.)
//
// So one pair-step = ONE multiply PT*PV (+ the polynomial in x for spin 2) + the accumulator FMAs; there is no
// per-pair state, no start value, no error growth along the family: every term carries ~4 roundings.
//
// Work decomposition: exactly the lockstep scheme of psb200_pair_v2.cuh.  block = ONE WARP = NR consecutive l1 rows x
// (32/NR) R pairs of ONE parity of d each (d = d_lo + 2 o); all pairs advance j = d_lo + 2 tau in lockstep, so the
// window values W'_q[j] are warp-uniform shared-memory broadcasts.  A step handles the even-parity term j and, for
// the jobs that need odd parity (M--, fused M++/M--, MASTER), the odd-parity term j+1.  The table indices of a
// thread's pairs fall (t/2 = tau - o) and rise (m/2 - d_lo = tau + o) by one per pair and per step: each thread loads
// ONE new entry of each table per step (conflict-free layout de-interleaved modulo R), the other R-1 values rotate
// through registers.  Tables are staged per chunk of TC steps from the global g / 1/g / 1/n tables.
#pragma once
#include "psb200_common.cuh"
#include "psb200_pair_v2.cuh"      // cp_async16, cp_async_wait_all, block_sync

namespace psb {

#ifndef PSB200_V3_R_LIGHT
#define PSB200_V3_R_LIGHT 12       // <= 2 accumulators per pair (the shared-memory path is their co-limiter: fewer bytes per pair-step)
#endif
#ifndef PSB200_V3_R_MID
#define PSB200_V3_R_MID 8          // 4-5 accumulators
#endif
#ifndef PSB200_V3_R_HEAVY
#define PSB200_V3_R_HEAVY 8        // 8 accumulators
#endif
#ifndef PSB200_V3_R_BOTH
#define PSB200_V3_R_BOTH 6         // even + odd parity (fused M++/M--, MASTER)
#endif
#ifndef PSB200_V3_NR
#define PSB200_V3_NR 2
#endif
__host__ __device__ constexpr int v3_r(int job)
{
    // the two-parity jobs carry four rotating windows: fewer pairs per thread keep them in registers
    if (job == JOB_MPPMMM || job == JOB_MASTER) return PSB200_V3_R_BOTH;
    return job_nacc(job) <= 2 ? PSB200_V3_R_LIGHT : (job_nacc(job) <= 5 ? PSB200_V3_R_MID : PSB200_V3_R_HEAVY);
}
// rows per warp, by job class like R (defaults: PSB200_V3_NR everywhere)
// MEASURED with the ring-staged kernel (B200, lmax 6143, ms per bench step; profiles/r02_v4_variants_probe_*.jsonl):
// NR 1 / 2 / 4 / 8 everywhere 99.4 / 93.7 / 91.6 / 104.2; NR 4 with R 8 for the covariance jobs 90.3; the light jobs
// (TT: 5.0 vs 5.3 ms) prefer NR 2 at R 8, and NR 4 at R 12 (4.7 ms).
#ifndef PSB200_V3_NR_LIGHT
#define PSB200_V3_NR_LIGHT (2 * PSB200_V3_NR)
#endif
#ifndef PSB200_V3_NR_MID
#define PSB200_V3_NR_MID (2 * PSB200_V3_NR)
#endif
#ifndef PSB200_V3_NR_HEAVY
#define PSB200_V3_NR_HEAVY (2 * PSB200_V3_NR)
#endif
#ifndef PSB200_V3_NR_BOTH
#define PSB200_V3_NR_BOTH (2 * PSB200_V3_NR)
#endif
__host__ __device__ constexpr int v3_nr(int job)
{
    if (job == JOB_MPPMMM || job == JOB_MASTER) return PSB200_V3_NR_BOTH;
    return job_nacc(job) <= 2 ? PSB200_V3_NR_LIGHT : (job_nacc(job) <= 5 ? PSB200_V3_NR_MID : PSB200_V3_NR_HEAVY);
}
__host__ __device__ constexpr bool v3_nr_ok(int n) { return n == 1 || n == 2 || n == 4 || n == 8; }
static_assert(v3_nr_ok(PSB200_V3_NR_LIGHT) && v3_nr_ok(PSB200_V3_NR_MID) && v3_nr_ok(PSB200_V3_NR_HEAVY) && v3_nr_ok(PSB200_V3_NR_BOTH),
              "rows per warp: 1, 2, 4 or 8");

__host__ __device__ constexpr bool v3_has_even(int job) { return job != JOB_MMM; }
__host__ __device__ constexpr bool v3_has_odd(int job) { return job == JOB_MMM || job == JOB_MPPMMM || job == JOB_MASTER; }
__host__ __device__ constexpr bool v3_spin2(int job) { return job_family(job) != FAM_00; }
__host__ __device__ constexpr int v3_ntab(int job) { return (v3_has_even(job) ? 1 : 0) + (v3_has_odd(job) ? 1 : 0); }
__host__ __device__ constexpr int v3_span(int job) { return (32 / v3_nr(job)) * v3_r(job); }   // pairs per row per warp
// x column(s): the spin-2 jobs read x = j(j+1) of the step (and x+1 for the odd-parity term) from the staged W' row --
// a broadcast shared-memory load -- instead of advancing it with two or three warp-uniform FP64 adds per step
// (the FP64 pipe is the bound; the shared-memory path has headroom).
#ifndef PSB200_V3_XCOL
#define PSB200_V3_XCOL 1
#endif
__host__ __device__ constexpr bool v3_xcol(int job) { return PSB200_V3_XCOL && v3_spin2(job); }
__host__ __device__ constexpr int v3_xc_even(int job) { return job_nw(job); }                                   // column of x
__host__ __device__ constexpr int v3_xc_odd(int job) { return job_nw(job) + (v3_has_even(job) ? 1 : 0); }       // column of x+1
__host__ __device__ constexpr int v3_nqp(int job)                                                                // W' columns (even)
{
    const int nx = v3_xcol(job) ? (v3_has_even(job) ? 1 : 0) + (v3_has_odd(job) ? 1 : 0) : 0;
    return (job_nw(job) + nx + 1) & ~1;
}
__host__ __device__ constexpr int v3_tc(int job)                                                 // steps per staged chunk
{
#ifdef PSB200_V3_TC
    return PSB200_V3_TC / v3_r(job) * v3_r(job);
#else
    return (v3_nqp(job) * v3_ntab(job) <= 2 ? 256 : 128) / v3_r(job) * v3_r(job);
#endif
}
constexpr int V3_TC_MAX = 256;
constexpr int V3_PB_MAX = 32 * 8;
__host__ __device__ constexpr int v3_szt(int job) { return v3_tc(job) + v3_span(job) + v3_r(job); }   // table entries per row per chunk
__host__ __device__ constexpr int v3_sub(int job) { return v3_szt(job) / v3_r(job) + 1; }             // de-interleaved sub-table stride
// doubles between the tables of consecutive rows of a warp (see v2_tstride: bank-conflict-free row groups)
__host__ __device__ constexpr int v3_tstride(int job)
{
    const int n = v3_ntab(job) * v3_r(job) * v3_sub(job);
    if (v3_ntab(job) == 1 && v3_nr(job) == 4) return n + ((8 - n % 16) + 16) % 16;
    return n + (n & 1);
}
__host__ __device__ constexpr int v3_smem_doubles(int job)
{
    return 2 * v3_nr(job) * v3_tstride(job) + v3_tc(job) * v3_ntab(job) * v3_nqp(job) + 2;
}

struct V3Tables {
    const double* gam;    // g(n) = binom(2n,n)/4^n
    const double* igam;   // 1/g(n)
    const double* INV;    // 1/n, INV[0] = 0
    int nS;               // length of all three
    const int4* blocks;   // (first l1, d_lo, end of the row band the tile belongs to, -), heaviest first
    const double* Wp;     // [row j][v3_nqp columns] = (2j+1) W_q[j] / 4pi, zero rows past lenW
};

// W'[j][q] for one call: (2j+1) W_q[j] / 4pi in columns q < NW (zero past lenW), then x_j = j(j+1) and x_j + 1 where
// the job reads them (v3_xcol), zero padding to the even column count.
template <int JOB>
__global__ void v3_prep_w(double* __restrict__ Wp, int rows, int lenW,
                          const double* w0, const double* w1, const double* w2, const double* w3,
                          const double* w4, const double* w5, const double* w6, const double* w7)
{
    constexpr int NW = job_nw(JOB), NQP = v3_nqp(JOB);
    const double* W[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    const double k = (double)(2 * j + 1) * INV_4PI;
    const double x = (double)j * (double)(j + 1);
#pragma unroll
    for (int q = 0; q < NQP; ++q) {
        double v = 0.0;
        if (q < NW) v = (j < lenW) ? k * W[q][j] : 0.0;
        else if (v3_xcol(JOB) && v3_has_even(JOB) && q == v3_xc_even(JOB)) v = x;
        else if (v3_xcol(JOB) && v3_has_odd(JOB) && q == v3_xc_odd(JOB)) v = x + 1.0;
        Wp[(size_t)j * NQP + q] = v;
    }
}

// Resident warps per SM the register allocation must allow (one warp per block), from an estimate of the live
// doubles per thread: accumulators, rotating windows, spin-2 constants, the step's window values.
__host__ __device__ constexpr int v3_reg_estimate(int job)
{
    const int r = v3_r(job);
    return 2 * (r * job_nacc(job) + v3_ntab(job) * 2 * (2 * r - 1) + (v3_spin2(job) ? 2 * r : 0) + v3_nqp(job) * v3_ntab(job)) + 36;
}
__host__ __device__ constexpr int v3_min_blocks(int job)
{
#ifdef PSB200_V3_MINB
    return PSB200_V3_MINB;
#else
    const int e = v3_reg_estimate(job);
    return e <= 128 ? 16 : (e <= 168 ? 12 : (e <= 200 ? 10 : 8));
#endif
}

template <int JOB>
__global__ void __launch_bounds__(32, v3_min_blocks(JOB)) pair_kernel_v3(const PairArgs A, const V3Tables T)
{
    constexpr bool EV = v3_has_even(JOB), OD = v3_has_odd(JOB), S2 = v3_spin2(JOB);
    constexpr int NWQ = job_nw(JOB), NACC = job_nacc(JOB), NQP = v3_nqp(JOB);
    constexpr int R = v3_r(JOB), NR = v3_nr(JOB), LPR = 32 / NR, SPAN = v3_span(JOB);
    constexpr int NTAB = v3_ntab(JOB);                 // 1: one parity; 2: (even, odd) entries side by side
    constexpr int RPS = NTAB;                          // W' rows per step
    constexpr int TC = v3_tc(JOB), SZT = v3_szt(JOB), SUB = v3_sub(JOB), TSTR = v3_tstride(JOB);
    constexpr bool XCOL = v3_xcol(JOB);
    constexpr int XE = v3_xc_even(JOB), XO = v3_xc_odd(JOB);

    extern __shared__ __align__(16) double smem[];
    double* shU = smem;                                // falling index: PT | QT | (PT, QT), one table per row
    double* shV = shU + NR * TSTR;                     // rising index:  PV | QV | (PV, QV)
    double* shW = shV + NR * TSTR;                     // [TC][RPS][NQP]   (TSTR is even: 16-byte aligned)

    const int4 blk = T.blocks[blockIdx.x];
    const int l1_first = blk.x, d_lo = blk.y, band_hi = blk.z;
    const int tid = threadIdx.x;
    const int rg = tid / LPR;                          // my row group
    const int l1 = l1_first + rg;                      // my row
    const int e = (tid % LPR) * R;                     // offset of my first pair inside the window
    const int dmax = (l1 < band_hi) ? A.lmax - l1 : -1;   // last valid d of my row (rows past the band: none)
    const int l1_last = min(l1_first + NR, band_hi) - 1;  // longest family of the warp
    // last step: the last pair (offset SPAN-1) of the last row finishes its family, or the window spectrum ends
    const int tau_end = (A.lenW - 1 - d_lo < 0) ? -1 : min(SPAN - 1 + l1_last, (A.lenW - 1 - d_lo) / 2);
    if (d_lo > A.lmax - l1_first || tau_end < 0) {     // nothing to sum: the stored values are exact zeros
        if (d_lo <= A.lmax - l1_first) {
            const double z[NACC > 0 ? NACC : 1] = {};
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int d = d_lo + 2 * (e + r);
                if (d <= dmax) epilogue<JOB>(A, l1, l1 + d, z);
            }
        }
        return;
    }
    const double* myU = shU + rg * TSTR;
    const double* myV = shV + rg * TSTR;

    double acc[R][NACC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[r][q] = 0.0;
    // spin 2: per-pair constants of  u = x - (a+b-1),  2N = u^2 - (2ab+1)
    double se[S2 ? R : 1], cc[S2 ? R : 1];
    if constexpr (S2) {
        const double a = (double)l1 * (double)(l1 + 1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int l2 = l1 + d_lo + 2 * (e + r);
            const double b = (double)l2 * (double)(l2 + 1);
            se[r] = a + b - 1.0;
            cc[r] = fma(2.0 * a, b, 1.0);
        }
    }
    // warp-uniform: x = j(j+1) of the even-parity term (advanced by 4j+6 per step) and 2j+3 (odd term: u_odd = u + 2j+3)
    double xj = (double)d_lo * (double)(d_lo + 1);
    double xinc = 4.0 * (double)d_lo + 6.0;
    double k2 = 2.0 * (double)d_lo + 3.0;
    (void)xj; (void)xinc; (void)k2;
    // rotating windows (2R-1 live entries each); the second pair holds the odd-parity tables when both are needed
    double wU0[2 * R - 1], wV0[2 * R - 1], wU1[NTAB > 1 ? 2 * R - 1 : 1], wV1[NTAB > 1 ? 2 * R - 1 : 1];
#pragma unroll
    for (int k = 0; k < 2 * R - 1; ++k) {
        wU0[k] = 0.0; wV0[k] = 0.0;
        if constexpr (NTAB > 1) { wU1[k] = 0.0; wV1[k] = 0.0; }
    }

    for (int tau0 = 0; tau0 <= tau_end; tau0 += TC) {
        block_sync();
        // ================= stage this chunk's tables =================
        // falling index: entry idx <-> nu = t/2 = tau0 - SPAN + idx - 1 of the even term (odd term: t = 2 nu + 1)
        for (int i = tid; i < NR * SZT; i += 32) {
            const int g = i / SZT, idx = i - g * SZT;
            const int lg = l1_first + g;
            const int pos = (idx % R) * SUB + idx / R;
            const int nu = tau0 - SPAN + idx - 1;
            double pt = 0.0, qt = 0.0;
            if (nu >= 0 && nu <= lg) {
                pt = __ldg(T.gam + nu) * __ldg(T.gam + (lg - nu));
                if constexpr (OD) qt = ((double)(2 * nu + 1) * (double)(2 * (lg - nu))) * pt;
            }
            if constexpr (NTAB > 1) reinterpret_cast<double2*>(shU + g * TSTR)[pos] = make_double2(pt, qt);
            else shU[g * TSTR + pos] = EV ? pt : qt;
        }
        // rising index: entry idx <-> sg = m/2 = tau0 + idx + d_lo - 1 of the even term (odd term: m = 2 sg + 1)
        for (int i = tid; i < NR * SZT; i += 32) {
            const int g = i / SZT, idx = i - g * SZT;
            const int lg = l1_first + g, Lg = 2 * lg + 1;
            const int pos = (idx % R) * SUB + idx / R;
            const int sg = tau0 + idx + d_lo - 1;
            double pv = 0.0, qv = 0.0;
            if (sg >= 0 && 2 * sg + Lg + 1 < T.nS) {
                pv = __ldg(T.gam + sg) * __ldg(T.igam + (sg + lg)) * __ldg(T.INV + (2 * sg + Lg));
                if constexpr (OD) qv = ((double)(2 * sg + 1) * (double)(2 * sg + 1 + Lg)) * pv;
            }
            if constexpr (NTAB > 1) reinterpret_cast<double2*>(shV + g * TSTR)[pos] = make_double2(pv, qv);
            else shV[g * TSTR + pos] = EV ? pv : qv;
        }
        // W' rows j = d_lo + 2 (tau0 + step) (+1 for the odd term), step < TC  (16-byte cp.async; rows past lenW are zero)
        {
            constexpr int CPR = NQP / 2;                       // 16-byte pieces per row
            if constexpr (RPS == 2) {
                // rows j and j+1 of consecutive steps are consecutive rows: one contiguous run
                const double* src = T.Wp + (size_t)(d_lo + 2 * tau0) * NQP;
                constexpr int NCH = TC * 2 * CPR;
                for (int c = tid; c < NCH; c += 32) cp_async16(shW + 2 * c, src + 2 * c);
            } else {
                const double* src = T.Wp + (size_t)(d_lo + 2 * tau0 + (EV ? 0 : 1)) * NQP;
                constexpr int NCH = TC * CPR;
                for (int c = tid; c < NCH; c += 32) {
                    const int row = c / CPR, piece = c % CPR;
                    cp_async16(shW + 2 * c, src + (size_t)row * 2 * NQP + 2 * piece);
                }
            }
            cp_async_wait_all();
        }
        block_sync();

        if (tau0 == 0) {
            // prime the carried part of the rising windows (k = 0..R-2 <-> idx = 1 + e + k)
#pragma unroll
            for (int k = 0; k < R - 1; ++k) {
                const int idx = 1 + e + k;
                const int pos = (idx % R) * SUB + idx / R;
                if constexpr (NTAB > 1) {
                    const double2 v = reinterpret_cast<const double2*>(myV)[pos];
                    wV0[k] = v.x; wV1[k] = v.y;
                } else {
                    wV0[k] = myV[pos];
                }
            }
        }

        const int tg_end = min(TC, tau_end - tau0 + 1);
        for (int tg = 0; tg < tg_end; tg += R) {
            // ---- the R new entries of every window ----
            {
                const int bu = (tg - e + SPAN) / R;        // idx = tg + 1 - e + u + SPAN
                const int bv = (tg + e) / R + 1;           // idx = tg + e + R + u
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    const int pu = ((1 + u) % R) * SUB + bu + (1 + u) / R;
                    const int pv = u * SUB + bv;
                    if constexpr (NTAB > 1) {
                        const double2 a = reinterpret_cast<const double2*>(myU)[pu];
                        const double2 c = reinterpret_cast<const double2*>(myV)[pv];
                        wU0[R - 1 + u] = a.x; wU1[R - 1 + u] = a.y;
                        wV0[R - 1 + u] = c.x; wV1[R - 1 + u] = c.y;
                    } else {
                        wU0[R - 1 + u] = myU[pu];
                        wV0[R - 1 + u] = myV[pv];
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < R; ++s) {
                // window spectra (and x columns) of this step: warp-uniform broadcast reads
                double w[NQP], wo[NQP];
                {
                    const double* wr = shW + (size_t)(tg + s) * (RPS * NQP);
                    if constexpr (NQP == 2 && NWQ == 1 && !XCOL && RPS == 1) {
                        w[0] = wr[0];
                    } else {
#pragma unroll
                        for (int q = 0; q < NQP; q += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(wr + q);
                            w[q] = v.x; w[q + 1] = v.y;
                        }
                    }
                    if constexpr (RPS == 2) {
                        // odd-parity row: its window value (the last W) and x+1
                        constexpr int QW = NWQ - 1;
#pragma unroll
                        for (int q = (QW & ~1); q < NQP; q += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(wr + NQP + q);
                            wo[q] = v.x; wo[q + 1] = v.y;
                        }
                    }
                }
                if constexpr (XCOL && EV) xj = w[XE];
                double xo1 = 0.0;                                   // x_{j+1} + 1 of the odd-parity term
                if constexpr (XCOL && OD) xo1 = (RPS == 2) ? wo[XO] : w[XO];
                (void)xo1;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int kU = s - r + R - 1, kV = r + s;
                    double u = 0.0;
                    if constexpr (S2 && (EV || !XCOL)) u = xj - se[r];
                    if constexpr (EV) {
                        const double g = wU0[kU] * wV0[kV];                 // f00(j)^2
                        if constexpr (!S2) {
#pragma unroll
                            for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(g, w[q], acc[r][q]);
                        } else {
                            const double nn = fma(u, u, -cc[r]);            // 2 N(x)
                            const double gn = g * nn;                       // 2 D f00 f22
                            if constexpr (JOB == JOB_M02 || JOB == JOB_TEEE) {
#pragma unroll
                                for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(gn, w[q], acc[r][q]);
                            } else if constexpr (JOB == JOB_TETE) {
                                acc[r][0] = fma(gn, w[0], acc[r][0]);
                                acc[r][1] = fma(g, w[1], acc[r][1]);
                                acc[r][2] = fma(gn, w[2], acc[r][2]);
                                acc[r][3] = fma(gn, w[3], acc[r][3]);
                                acc[r][4] = fma(gn, w[4], acc[r][4]);
                            } else if constexpr (JOB == JOB_MASTER) {
                                const double gnn = gn * nn;
                                acc[r][0] = fma(g, w[0], acc[r][0]);
                                acc[r][1] = fma(gn, w[1], acc[r][1]);
                                acc[r][2] = fma(gn, w[2], acc[r][2]);
                                acc[r][3] = fma(gnn, w[3], acc[r][3]);
                            } else if constexpr (JOB == JOB_MPPMMM) {
                                acc[r][0] = fma(gn * nn, w[0], acc[r][0]);
                            } else {                                        // MPP, EEEE, TEEEP: 4 D^2 f22^2
                                const double gnn = gn * nn;
#pragma unroll
                                for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(gnn, w[q], acc[r][q]);
                            }
                        }
                    }
                    if constexpr (OD) {
                        // odd-parity term j+1: 4 D^2 f22^2 = (x' - a - b + 2)^2 QT QV,  x' - a - b + 2 = u + 2j + 3
                        const double uo = XCOL ? xo1 - se[r] : u + k2;       // (x' + 1) - (a + b - 1)
                        const double h = (NTAB > 1 ? wU1[kU] * wV1[kV] : wU0[kU] * wV0[kV]) * uo;
                        if constexpr (JOB == JOB_MMM) acc[r][0] = fma(h * uo, w[0], acc[r][0]);
                        else acc[r][NACC - 1] = fma(h * uo, wo[NWQ - 1], acc[r][NACC - 1]);   // MPPMMM (W0), MASTER (W3)
                    }
                }
                if constexpr (S2 && !XCOL) { xj += xinc; xinc += 8.0; }
                if constexpr (OD && !XCOL) k2 += 4.0;
            }
            // ---- rotate: next group's carried entries ----
#pragma unroll
            for (int k = 0; k < R - 1; ++k) {
                wU0[k] = wU0[k + R]; wV0[k] = wV0[k + R];
                if constexpr (NTAB > 1) { wU1[k] = wU1[k + R]; wV1[k] = wV1[k + R]; }
            }
        }
    }

    // ---- epilogue: one stored value per output and valid pair ----
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int d = d_lo + 2 * (e + r);
        if (d <= dmax) {
            if constexpr (S2) {
                // 1/(4 D^2) and 1/(2 D); rows l1 < 2 never come here (psb200_lowrows.cuh)
                const int l2 = l1 + d;
                const double id2 = 0.25 / (((double)(l1 - 1) * (double)l1 * ((double)(l1 + 1) * (double)(l1 + 2)))
                                           * ((double)(l2 - 1) * (double)l2 * ((double)(l2 + 1) * (double)(l2 + 2))));
                const double id1 = sqrt(id2);
                if constexpr (JOB == JOB_M02 || JOB == JOB_TEEE) {
#pragma unroll
                    for (int q = 0; q < NACC; ++q) acc[r][q] *= id1;
                } else if constexpr (JOB == JOB_TETE) {
                    acc[r][0] *= id1; acc[r][2] *= id1; acc[r][3] *= id1; acc[r][4] *= id1;
                } else if constexpr (JOB == JOB_MASTER) {
                    acc[r][1] *= id1; acc[r][2] *= id1; acc[r][3] *= id2; acc[r][4] *= id2;
                } else {
#pragma unroll
                    for (int q = 0; q < NACC; ++q) acc[r][q] *= id2;
                }
            }
            epilogue<JOB>(A, l1, l1 + d, acc[r]);
        }
    }
}

// host side of one launch; returns a cudaError_t value (0 = ok)
template <int JOB>
int launch_pair_v3(const PairArgs& A, const V3Tables& T, int nblocks, cudaStream_t st)
{
    constexpr int smem = v3_smem_doubles(JOB) * (int)sizeof(double);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pair_kernel_v3<JOB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_done[dev] = true;
    }
    if (nblocks > 0) pair_kernel_v3<JOB><<<nblocks, 32, smem, st>>>(A, T);
    return (int)cudaGetLastError();
}

}  // namespace psb
