// psb200_pair_v4.cuh -- closed-form pair kernel with fully asynchronous table staging (sm_100a).
//
// Arithmetic and work decomposition are those of psb200_pair_v3.cuh (closed-form products PT[t] PV[m], lockstep
// warps, register-rotated windows).  What changes is where the per-row tables come from.  ncu on v3 (profiles/
// r02_ncu_v3_summary.txt): 16-19 % of the warp samples of the covariance jobs (47 % for the TT matrix) sit in the
// per-chunk staging loops -- dependent global loads, integer division, FP64 products -- during which the other one or
// two warps of the scheduler cannot keep the FP64 pipe full (72-78 % pipe utilisation).  Here the warp never leaves its
// main loop for longer than a few dozen instructions:
//   * every per-row table factorises into two ROW-INDEPENDENT one-dimensional sequences read at a row shift,
//         PT_l[nu]    = G0[nu]    * G0[l - nu]          QT_l[nu]    = G1[nu]    * G2[l - nu]
//         PV_l[sigma] = G0[sigma] * H0[sigma + l]       QV_l[sigma] = G1[sigma] * H1[sigma + l]
//     G0[n] = g(n) = binom(2n,n)/4^n, G1 = (2n+1) g, G2 = 2n g, H0 = 1/((2n+1) g), H1 = (2n+2) H0 (global, ~0.3 MB each,
//     zero for n < 0 so that the family ends need no test);
//   * work advances in sub-chunks of P = 30 or 32 steps.  While sub-chunk c is computed, the raw factors and the W' rows
//     of sub-chunk c+1 are in flight as cp.async copies (one 8-byte element per lane and sequence, no registers held);
//     at the boundary each lane multiplies ITS element pair and stores the product into a ring (de-interleaved modulo R
//     like v3's tables), then issues the copies for sub-chunk c+2.  No global-load latency is ever waited for after the
//     prologue, and no table entry is computed twice (v3 re-stages SPAN+R entries of every chunk).
// Shared memory per warp drops from 16 KB to ~10 KB (rings of SPAN+P entries instead of chunk tables).
#pragma once
#include "psb200_pair_v3.cuh"

namespace psb {

constexpr int V4_PAD = 1024;          // zero entries in front of every global sequence (indices down to -V4_PAD are valid)

// PSB200_V4_EMBED = 1: the products of pass c+1 are computed INSIDE the groups of sub-chunk c (one row of the warp per
// group), from raw factors issued a whole sub-chunk earlier, so that the boundary shrinks to a wait that never waits, a
// warp sync and the next copies.  MEASURED (B200, lmax 6143): no gain for the covariance jobs (TTTT 18.02 vs 18.12 ms,
// EEEE 26.07 vs 26.10), a loss for the light and two-parity jobs (TT 5.29 vs 4.92, fused M++/M-- 24.5 vs 21.2: the
// extra sub-chunk of ring lookahead costs shared memory) -- the boundary is not what keeps the FP64 pipe at 82-85 %.
// Default 0: products at the boundary.
#ifndef PSB200_V4_EMBED
#define PSB200_V4_EMBED 0
#endif
constexpr bool V4_EMBED = PSB200_V4_EMBED != 0;
__host__ __device__ constexpr int v4_p(int job) { return (32 / v3_r(job)) * v3_r(job); }          // steps per sub-chunk
__host__ __device__ constexpr int v4_g(int job) { return 32 / v3_r(job); }                         // groups per sub-chunk
// Ring length / R: the live range is LPR + G sub-indices; rounded up to a power of two so that the wrap is a mask and
// -- the point -- the LPR consecutive sub-indices one row group reads in one instruction fall into distinct banks
// whether or not they wrap (with a length of 21 every wrapped window put two lanes 16 doubles apart on one bank:
// 17 % of the shared-memory wavefronts of the first build were conflict replays).
__host__ __device__ constexpr int v4_sublen(int job)
{
    // live sub-indices: LPR + G with products at the boundary; one more sub-chunk of lookahead when they are embedded
    const int need = V4_EMBED ? 32 / v3_nr(job) - 1 + 2 * v4_g(job) : 32 / v3_nr(job) + v4_g(job);
    return need <= 16 ? 16 : (need <= 32 ? 32 : 64);
}
// Distance of the R residue sub-tables of a ring: odd, so that the staging lanes of one sub-index (consecutive residues)
// store to distinct banks (with the power-of-two distance the stores of a pass were R-way conflicts).
__host__ __device__ constexpr int v4_subp(int job) { return v4_sublen(job) + 1; }
// Doubles between the rings of consecutive rows of a warp.  With 64-bit loads a half-warp is one wavefront; for NR = 4
// it holds two row groups of 8 lanes reading the same sub-indices of their own rings: conflict-free iff the stride is
// 8 modulo 16 doubles (NR = 8: four groups of 4 lanes, stride 4 modulo 16).
__host__ __device__ constexpr int v4_tstride(int job)
{
    const int n = v3_ntab(job) * v3_r(job) * v4_subp(job), nr = v3_nr(job);
    int want = -1;                                             // required stride modulo 16 doubles
    if (v3_ntab(job) == 1 && nr >= 4) want = 32 / nr;          // 64-bit loads: 2 (NR 4) or 4 (NR 8) row groups per half-warp
    if (v3_ntab(job) == 2 && nr == 8) want = 8;                // 128-bit loads: a quarter-warp holds 2 groups only for NR 8
    if (want < 0) return n + (n & 1);
    return n + ((want - n % 16) + 16) % 16;
}
__host__ __device__ constexpr int v4_wrows(int job) { return 2 * v4_p(job); }
__host__ __device__ constexpr int v4_raw_slot(int job) { return v3_ntab(job) * (2 + 2 * v3_nr(job)) * 32; }   // one pass
__host__ __device__ constexpr int v4_raw_doubles(int job) { return (V4_EMBED ? 2 : 1) * v4_raw_slot(job); }
__host__ __device__ constexpr int v4_smem_doubles(int job)
{
    return 2 * v3_nr(job) * (v4_tstride(job) + (v4_tstride(job) & 1)) + v4_wrows(job) * v3_ntab(job) * v3_nqp(job)
         + v4_raw_doubles(job) + 2;
}

struct V4Tables {
    const double *G0, *G1, *G2, *H0, *H1;   // element 0 of each sequence; [-V4_PAD, 0) are zeros
    int nS;
    const int4* blocks;                      // (first l1, d_lo, end of the tile's row band, -), heaviest first
    const double* Wp;                        // [row j][v3_nqp columns], see v3_prep_w
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

template <int JOB>
__global__ void __launch_bounds__(32, v3_min_blocks(JOB)) pair_kernel_v4(const PairArgs A, const V4Tables T)
{
    constexpr bool EV = v3_has_even(JOB), OD = v3_has_odd(JOB), S2 = v3_spin2(JOB);
    constexpr int NWQ = job_nw(JOB), NACC = job_nacc(JOB), NQP = v3_nqp(JOB);
    constexpr int R = v3_r(JOB), NR = v3_nr(JOB), LPR = 32 / NR, SPAN = v3_span(JOB);
    constexpr int NTAB = v3_ntab(JOB), RPS = NTAB;
    constexpr bool XCOL = v3_xcol(JOB);
    constexpr int XE = v3_xc_even(JOB), XO = v3_xc_odd(JOB);
    constexpr int P = v4_p(JOB), G = v4_g(JOB), SUBLEN = v4_sublen(JOB), SUBP = v4_subp(JOB);
    constexpr int RS = 32, RSLOT = v4_raw_slot(JOB);    // stride of the raw arrays, doubles per raw slot
    constexpr int RPG = (NR + G - 1) / G;               // rows whose products one group computes (embedded mode)
    constexpr int TSTR = v4_tstride(JOB) + (v4_tstride(JOB) & 1);
    constexpr int WR = v4_wrows(JOB), ROWD = RPS * NQP;       // W' ring rows, doubles per ring row
    constexpr int CMIN = -((LPR + G - 1) / G);                // first prologue pass: covers the priming entries

    extern __shared__ __align__(16) double smem[];
    double* shU = smem;                                // product rings, falling index: PT | QT | (PT, QT) per row
    double* shV = shU + NR * TSTR;                     // rising index:  PV | QV | (PV, QV)
    double* shW = shV + NR * TSTR;                     // [WR][RPS][NQP]
    double* raw = shW + WR * ROWD;                     // raw factors: [slot][NTAB][2 + 2 NR][32]

    const int4 blk = T.blocks[blockIdx.x];
    const int l1_first = blk.x, d_lo = blk.y, band_hi = blk.z;
    const int tid = threadIdx.x;
    const int rg = tid / LPR, eR = tid % LPR;          // my row group; my pair offset / R
    const int l1 = l1_first + rg;
    const int e = eR * R;
    const int dmax = (l1 < band_hi) ? A.lmax - l1 : -1;
    const int l1_last = min(l1_first + NR, band_hi) - 1;
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

    // ---- staging: pass c provides U entries nu in [cP, (c+1)P), V entries mu' = mu + 1 in [cP + SPAN, (c+1)P + SPAN)
    //      and the W' rows of steps [cP, (c+1)P); lane k < P owns element k of the pass ----
    const int kq = tid / R, kr = tid % R;              // ring coordinates of my element: sub-index offset, residue
    const double* seqA0 = EV ? T.G0 : T.G1;            // falling: [nu]
    const double* seqB0 = EV ? T.G0 : T.G2;            //          [l - nu]
    const double* seqD0 = EV ? T.H0 : T.H1;            // rising:  [sigma + l]   (C = A sequences at [sigma])
    auto issue_raw = [&](int c) {                      // raw factors of pass c into slot c & 1 (embedded) / slot 0
        if (tid < P) {
            const int nu = c * P + tid;
            const int sg = d_lo + c * P + SPAN + tid - 1;              // sigma = d_lo + mu' - 1
#pragma unroll
            for (int h = 0; h < NTAB; ++h) {
                const double* sA = h == 0 ? seqA0 : T.G1;
                const double* sB = h == 0 ? seqB0 : T.G2;
                const double* sD = h == 0 ? seqD0 : T.H1;
                double* rw = raw + (V4_EMBED ? (c & 1) * RSLOT : 0) + h * (2 + 2 * NR) * RS;
                cp_async8(rw + tid, sA + nu);
                cp_async8(rw + (1 + NR) * RS + tid, sA + sg);
#pragma unroll
                for (int g = 0; g < NR; ++g) {
                    cp_async8(rw + (1 + g) * RS + tid, sB + (l1_first + g - nu));
                    cp_async8(rw + (2 + NR + g) * RS + tid, sD + (sg + l1_first + g));
                }
            }
        }
    };
    auto issue_w = [&](int c) {
        // W' ring row of step cP + k: row j = d_lo + 2 tau (+1: odd-only jobs), RPS consecutive rows per step; lane k
        // copies ring row k piece by piece (immediate offsets, no index arithmetic)
        if (tid < P) {
            constexpr int CPR = ROWD / 2;                              // 16-byte pieces per ring row
            double* dst = shW + (size_t)((c & 1) * P + tid) * ROWD;
            const double* src = T.Wp + (size_t)(d_lo + 2 * (c * P + tid) + ((EV || RPS == 2) ? 0 : 1)) * NQP;
#pragma unroll
            for (int piece = 0; piece < CPR; ++piece) cp_async16(dst + 2 * piece, src + 2 * piece);
        }
    };
    auto commit = [&]() { asm volatile("cp.async.commit_group;\n" ::: "memory"); };
    auto wait_all = [&]() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); };
    // products of rows [g0, g1) of pass c: lane k owns element k; ring sub-index U (cG + k/R) mod SUBLEN,
    // V (cG + LPR + k/R) mod SUBLEN with cgm = (cG) mod SUBLEN
    auto product_rows = [&](int c, int cgm, int g0, int g1) {
        if (tid < P) {
            const int su = (cgm + kq) & (SUBLEN - 1);
            const int sv = (cgm + LPR + kq) & (SUBLEN - 1);
            const double* slot = raw + (V4_EMBED ? (c & 1) * RSLOT : 0);
            for (int g = g0; g < g1; ++g) {
                double pu[NTAB], pv[NTAB];
#pragma unroll
                for (int h = 0; h < NTAB; ++h) {
                    const double* rw = slot + h * (2 + 2 * NR) * RS;
                    pu[h] = rw[tid] * rw[(1 + g) * RS + tid];
                    pv[h] = rw[(1 + NR) * RS + tid] * rw[(2 + NR + g) * RS + tid];
                }
                if constexpr (NTAB > 1) {
                    reinterpret_cast<double2*>(shU + g * TSTR)[kr * SUBP + su] = make_double2(pu[0], pu[1]);
                    reinterpret_cast<double2*>(shV + g * TSTR)[kr * SUBP + sv] = make_double2(pv[0], pv[1]);
                } else {
                    shU[g * TSTR + kr * SUBP + su] = pu[0];
                    shV[g * TSTR + kr * SUBP + sv] = pv[0];
                }
            }
        }
    };

    // ---- per-pair state ----
    double acc[R][NACC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[r][q] = 0.0;
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
    double xj = (double)d_lo * (double)(d_lo + 1);
    double xinc = 4.0 * (double)d_lo + 6.0;
    double k2 = 2.0 * (double)d_lo + 3.0;
    (void)xj; (void)xinc; (void)k2;
    double wU0[2 * R - 1], wV0[2 * R - 1], wU1[NTAB > 1 ? 2 * R - 1 : 1], wV1[NTAB > 1 ? 2 * R - 1 : 1];
#pragma unroll
    for (int k = 0; k < 2 * R - 1; ++k) {
        wU0[k] = 0.0; wV0[k] = 0.0;
        if constexpr (NTAB > 1) { wU1[k] = 0.0; wV1[k] = 0.0; }
    }

    // ---- prologue: passes CMIN .. 0 synchronously (the V ring needs mu' from 1, the U ring zeros for nu < 0) ----
    {
        int cgm = ((CMIN * G) % SUBLEN + SUBLEN) % SUBLEN;
        for (int c = CMIN; c <= 0; ++c) {
            issue_raw(c);
            if (c == 0) issue_w(0);
            commit();
            wait_all();
            __syncwarp();
            product_rows(c, cgm, 0, NR);
            __syncwarp();
            cgm = (cgm + G) & (SUBLEN - 1);
        }
        // embedded mode: the raw factors of pass 1 must have landed when sub-chunk 0 starts (its groups multiply them);
        // boundary mode: pass 1 (raw + W') just takes off
        issue_raw(1);
        if constexpr (!V4_EMBED) issue_w(1);
        commit();
    }
    // prime the carried part of the rising windows: k = 0..R-2 <-> mu' = e + k + 1: residue k+1, sub-index eR
#pragma unroll
    for (int k = 0; k < R - 1; ++k) {
        const int pos = (k + 1) * SUBP + eR;
        if constexpr (NTAB > 1) {
            const double2 v = reinterpret_cast<const double2*>(myV)[pos];
            wV0[k] = v.x; wV1[k] = v.y;
        } else {
            wV0[k] = myV[pos];
        }
    }

    // reader ring coordinates at group t = tau / R: U sub-index (t - eR) mod SUBLEN, V (t + eR + 1) mod SUBLEN, W' row tau mod WR
    int qU = (SUBLEN - eR) % SUBLEN, qV = (eR + 1) % SUBLEN, wrow = 0;
    int cgm1 = G % SUBLEN;                                   // (c G) mod SUBLEN of the pass whose products come next (c + 1)
    for (int c = 0; c * P <= tau_end; ++c) {
        if constexpr (V4_EMBED) {
            // boundary: raw(c+1) -- issued a sub-chunk ago -- and W'(c) are in; the products of pass c that the groups
            // of the previous sub-chunk stored become visible; raw(c+2) and W'(c+1) take off
            wait_all();
            __syncwarp();
            if ((c + 1) * P <= tau_end) {
                issue_raw(c + 2);
                issue_w(c + 1);
                commit();
            }
        } else if (c > 0) {
            // boundary: the copies of pass c have landed -> products into the rings -> copies of pass c+1 take off
            wait_all();
            __syncwarp();
            product_rows(c, cgm1, 0, NR);
            __syncwarp();
            cgm1 = (cgm1 + G) & (SUBLEN - 1);
            if ((c + 1) * P <= tau_end) { issue_raw(c + 1); issue_w(c + 1); commit(); }
        }
        const int g_end = min(G, (tau_end - c * P) / R + 1);
        for (int gi = 0; gi < g_end; ++gi) {
            if constexpr (V4_EMBED) {
                // my share of the products of pass c+1 (rows gi RPG ..): independent of everything below, so its
                // LDS -> DMUL -> STS chains run under the group's FP64 stream
                if (gi * RPG < NR) product_rows(c + 1, cgm1, gi * RPG, min(NR, (gi + 1) * RPG));
            }
            // ---- the R new entries of every window: residue u, one sub-index for all u ----
#pragma unroll
            for (int u = 0; u < R; ++u) {
                if constexpr (NTAB > 1) {
                    const double2 a = reinterpret_cast<const double2*>(myU)[u * SUBP + qU];
                    const double2 b = reinterpret_cast<const double2*>(myV)[u * SUBP + qV];
                    wU0[R - 1 + u] = a.x; wU1[R - 1 + u] = a.y;
                    wV0[R - 1 + u] = b.x; wV1[R - 1 + u] = b.y;
                } else {
                    wU0[R - 1 + u] = myU[u * SUBP + qU];
                    wV0[R - 1 + u] = myV[u * SUBP + qV];
                }
            }
            const double* wgrp = shW + (size_t)wrow * ROWD;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                // window spectra (and x columns) of this step: warp-uniform broadcast reads
                double w[NQP], wo[NQP];
                {
                    const double* wr = wgrp + s * ROWD;
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
                        constexpr int QW = NWQ - 1;
#pragma unroll
                        for (int q = (QW & ~1); q < NQP; q += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(wr + NQP + q);
                            wo[q] = v.x; wo[q + 1] = v.y;
                        }
                    }
                }
                if constexpr (XCOL && EV) xj = w[XE];
                double xo1 = 0.0;
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
                        const double uo = XCOL ? xo1 - se[r] : u + k2;       // (x' + 1) - (a + b - 1)
                        const double h = (NTAB > 1 ? wU1[kU] * wV1[kV] : wU0[kU] * wV0[kV]) * uo;
                        if constexpr (JOB == JOB_MMM) acc[r][0] = fma(h * uo, w[0], acc[r][0]);
                        else acc[r][NACC - 1] = fma(h * uo, wo[NWQ - 1], acc[r][NACC - 1]);   // MPPMMM (W0), MASTER (W3)
                    }
                }
                if constexpr (S2 && !XCOL) { xj += xinc; xinc += 8.0; }
                if constexpr (OD && !XCOL) k2 += 4.0;
            }
            // ---- rotate windows, advance the ring coordinates ----
#pragma unroll
            for (int k = 0; k < R - 1; ++k) {
                wU0[k] = wU0[k + R]; wV0[k] = wV0[k + R];
                if constexpr (NTAB > 1) { wU1[k] = wU1[k + R]; wV1[k] = wV1[k + R]; }
            }
            qU = (qU + 1) & (SUBLEN - 1);
            qV = (qV + 1) & (SUBLEN - 1);
            wrow += R; if (wrow == WR) wrow = 0;
        }
        if constexpr (V4_EMBED) cgm1 = (cgm1 + G) & (SUBLEN - 1);
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");      // nothing may be in flight when the block retires

    // ---- epilogue: one stored value per output and valid pair ----
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int d = d_lo + 2 * (e + r);
        if (d <= dmax) {
            if constexpr (S2) {
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

template <int JOB>
int launch_pair_v4(const PairArgs& A, const V4Tables& T, int nblocks, cudaStream_t st)
{
    constexpr int smem = v4_smem_doubles(JOB) * (int)sizeof(double);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pair_kernel_v4<JOB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_done[dev] = true;
    }
    if (nblocks > 0) pair_kernel_v4<JOB><<<nblocks, 32, smem, st>>>(A, T);
    return (int)cudaGetLastError();
}

}  // namespace psb
