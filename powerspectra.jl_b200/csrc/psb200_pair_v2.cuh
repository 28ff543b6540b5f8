// psb200_pair_v2.cuh -- tuned pair kernel for sm_100a (FP64-pipe bound by design).
//
// Work decomposition
//   block  = ONE WARP = NR consecutive l1 rows x (32/NR) R consecutive d = l2-l1 of each row;
//   thread = R consecutive d of ONE row (R = 8 for the one- and two-accumulator jobs, 6 for the covariance
//   jobs); lanes [g LPR, (g+1) LPR) of the warp, LPR = 32/NR, hold row l1 + g over the SAME d window.
//   (Warp-sized blocks: table staging needs only __syncwarp, so the ~8-12 resident warps of an SM
//   drift apart and hide each other's staging latency; ncu showed 9-12% barrier stalls with
//   4-warp blocks.)
//   All pairs of a warp advance l3 = j in LOCKSTEP (j = tau + d_w, d_w = first d of the warp), so
//   every window spectrum read W'_q[j] is a warp-uniform shared-memory broadcast and feeds all
//   pairs with no per-pair loads.  A pair becomes live when j reaches its own jmin = d
//   (injection of the closed-form start value), and dies by itself at jmax (tables are 0 there).
//   Why several rows per warp: a lockstep warp spends SPAN-1 + 2 l1/DS + 1 steps on families that are
//   only 2 l1/DS + 1 steps long (SPAN = pairs per row per warp: the last pair starts SPAN-1 steps after the
//   first), and the last tile of every row is partly empty.  With one row per warp (SPAN = 32 R) 13.6% of
//   the executed pair-steps of a covariance job at lmax 6143 were dead; rows l1 and l1+1 share the d window,
//   the parity class and (to two steps) the family length, so putting NR of them side by side divides SPAN by NR
//   at the same register rotation depth R: dead pair-steps 13.6% -> 7.2% (NR = 2) -> 3.8% (NR = 4); the price is
//   one set of per-row tables per row (staging work x (NR TC + 32 R)/(TC + 32 R)).
//
// Recurrence (reduced Schulten-Gordon, m1 = 0), t = j-d, m = j+d, L = 2 l1+1:
//   a(j)^2 = (j^2-d^2)(s^2-j^2) = [t (L-t)] * [m (m+L)]   =>   a(j) = U[t] * V[m]
//   with PER-ROW one-dimensional tables U[t] = sqrt(t(L-t)) (index falls along a thread's pairs)
//   and V[m] = sqrt(m(m+L)) (index rises).  f22: f(j+1) = -(4(2j+1) f(j) + a(j) f(j-1)) / a(j+1)
//   costs 5 FP64 ops per term: no sqrt, no divide.  f00^2 obeys the two-step rational relation
//   g(j+2) = g(j) a(j+1)^2 / a(j+2)^2 = g(j) RT[t+1] RV[m+1]  (2 ops per two terms).
//   For the (0,0,0)-only jobs a warp therefore takes pairs of ONE parity of d (d = d_lo + 2o) and
//   steps l3 by 2: every pair is live on every step and odd-parity terms are never visited.
//   Tables are staged per chunk of V2_TC steps into shared memory (products of the global
//   sqrt(n), 1/sqrt(n), 1/n tables) in a layout de-interleaved modulo R so that the one new
//   value each thread needs per step is a conflict-free 64-bit load; the other R-1 values a
//   thread's pairs need are the ones its neighbours used one step earlier (register rotation).
//   W' = (2j+1) W / 4pi is pre-multiplied once per call and staged with cp.async.
//
// Normalisation: instead of running the family to jmax and dividing by sum (2j+1) f^2 (what the
// reference's dependency does), the start value is the closed-form "stretched" symbol
//   f00(d)^2 = g(d) g(l1) / (g(l2) (2 l2+1)),  g(n) = binom(2n,n) / 4^n,
//   f22(d)^2 = f00(d)^2 * (l2+1)(l2+2)(l1-1) l1 / ((l2-1) l2 (l1+1)(l1+2)),   same sign,
// so terms with l3 > lenW-1 (25% of every family at nV = lmax+1) are never evaluated.
// Rows l1 < 2 of the spin-2 jobs (true symbol 0, closed forms 0/0) are not this kernel's: psb200_lowrows.cuh.
#pragma once
#include "psb200_common.cuh"

namespace psb {

// Pairs per thread, by job weight.  More pairs per thread = fewer shared-memory bytes per pair-step (each
// new table entry serves R pairs) but more registers and a longer start skew.
// MEASURED (B200, lmax 6143, ms per bench step, NR = 1): R = 4/4/4 154.4, 6/6/4 151.2, 6/6/6 149.5, 8/6/4 150.8,
// 8/6/6 148.5 -- the shared-memory return path, not occupancy, is the co-limiter next to the FP64 pipe.
#ifndef PSB200_R_LIGHT
#define PSB200_R_LIGHT 8        // <= 2 accumulators per pair
#endif
#ifndef PSB200_R_MID
#define PSB200_R_MID 6          // 4-5 accumulators
#endif
#ifndef PSB200_R_HEAVY
#define PSB200_R_HEAVY 6        // 8 accumulators
#endif
__host__ __device__ constexpr int v2_r(int job)
{
    return job_nacc(job) <= 2 ? PSB200_R_LIGHT : (job_nacc(job) <= 5 ? PSB200_R_MID : PSB200_R_HEAVY);
}
constexpr int V2_R_MAX = 8;
// Rows per warp (1, 2 or 4), by job weight like R.
#ifndef PSB200_NR
#define PSB200_NR 2
#endif
#ifndef PSB200_NR_LIGHT
#define PSB200_NR_LIGHT PSB200_NR
#endif
#ifndef PSB200_NR_MID
#define PSB200_NR_MID PSB200_NR
#endif
#ifndef PSB200_NR_HEAVY
#define PSB200_NR_HEAVY PSB200_NR
#endif
__host__ __device__ constexpr int v2_nr(int job)
{
    return job_nacc(job) <= 2 ? PSB200_NR_LIGHT : (job_nacc(job) <= 5 ? PSB200_NR_MID : PSB200_NR_HEAVY);
}
static_assert(PSB200_NR_LIGHT == 1 || PSB200_NR_LIGHT == 2 || PSB200_NR_LIGHT == 4, "rows per warp: 1, 2 or 4");
static_assert(PSB200_NR_MID == 1 || PSB200_NR_MID == 2 || PSB200_NR_MID == 4, "rows per warp: 1, 2 or 4");
static_assert(PSB200_NR_HEAVY == 1 || PSB200_NR_HEAVY == 2 || PSB200_NR_HEAVY == 4, "rows per warp: 1, 2 or 4");
// (Measured and dropped in round 1: splitting a warp into lockstep groups that run at DIFFERENT l3 -- it shortens
// the skew too, but the W' row stops being one broadcast per step: 153.1 -> 156.9 / 163.1 ms per step.)
__host__ __device__ constexpr int v2_span(int job) { return (32 / v2_nr(job)) * v2_r(job); }   // pairs per row per warp
constexpr int V2_TC_MAX = 256;                 // upper bound of v2_tc()
// steps per staged chunk: longer chunks where the W' tile is small (fewer staging events)
__host__ __device__ constexpr int v2_tc(int job);
constexpr int V2_THREADS = 32;
__host__ __device__ constexpr int v2_pb(int r) { return V2_THREADS * r; }            // pairs per block (= per warp)
constexpr int V2_PB_MAX = V2_THREADS * V2_R_MAX;
__host__ __device__ constexpr int v2_szt(int job) { return v2_tc(job) + v2_span(job) + v2_r(job); }   // table entries per row per chunk
__host__ __device__ constexpr int v2_sub(int job) { return v2_szt(job) / v2_r(job) + 1; }             // de-interleaved sub-table stride

__host__ __device__ constexpr int v2_nqp(int job) { return (job_nw(job) + 1) & ~1; }   // W' columns (even)
// Family as the TUNED kernel evaluates it.  Jobs that only ever use EVEN-parity terms (M02, M++, EEEE,
// TETE, TEEE, TEEE_planck) never run the three-term spin-2 recurrence: for l1+l2+j even
//     (j l1 l2; 0 -2 2) = (j l1 l2; 0 0 0) * N(x) / D,      x = j(j+1), a = l1(l1+1), b = l2(l2+1),
//     N(x) = (x-a-b)(x-a-b+2)/2 - a b,   D = sqrt((l1-1) l1 (l1+1)(l1+2) (l2-1) l2 (l2+1)(l2+2))
// (two steps of the m-ladder; checked exactly against sympy, tests/test_oracle.py), so they ride the
// cheap two-step f00^2 recurrence with l3 step 2, exactly like the (0,0,0) jobs.  Only jobs that need
// odd parity (M--, fused M++/M--, MASTER) keep the spin-2 recurrence.
__host__ __device__ constexpr bool v2_even_only_spin2(int job)
{
    return job == JOB_M02 || job == JOB_MPP || job == JOB_EEEE || job == JOB_TETE || job == JOB_TEEEP || job == JOB_TEEE;
}
__host__ __device__ constexpr int v2_family(int job)
{
    return v2_even_only_spin2(job) ? FAM_00 : job_family(job);
}
__host__ __device__ constexpr int v2_ntab(int job) { return v2_family(job) == FAM_00 ? 1 : 2; }
__host__ __device__ constexpr int v2_tc(int job)
{
#ifdef PSB200_TC_ALL
    return PSB200_TC_ALL;
#else
    return (v2_nqp(job) <= 2 ? 256 : 128) / v2_r(job) * v2_r(job);
#endif
}
// Doubles between the tables of consecutive rows of a warp.  With 64-bit table loads (NTAB = 1) one shared-memory
// wavefront serves a half-warp; for NR = 4 that is two row groups of 8 lanes, each reading 8 consecutive doubles at
// the same relative position of its own table: conflict-free iff the stride is 8 modulo 16 doubles.  (128-bit
// loads are served per quarter-warp = at most one row group; NR = 2 puts one row group in each half-warp.)
__host__ __device__ constexpr int v2_tstride(int job)
{
    const int n = v2_ntab(job) * v2_r(job) * v2_sub(job);
    if (v2_ntab(job) == 1 && v2_nr(job) == 4) return n + ((8 - n % 16) + 16) % 16;
    return n + (n & 1);
}
__host__ __device__ constexpr int v2_smem_doubles(int job)
{
    return 2 * v2_nr(job) * v2_tstride(job) + v2_tc(job) * v2_nqp(job)
         + v2_pb(v2_r(job)) * (v2_family(job) == FAM_02 ? 2 : 1) + 2;
}

struct V2Tables {
    const double* S;      // sqrt(n)
    const double* IS;     // 1/sqrt(n), IS[0] = 0
    const double* INV;    // 1/n, INV[0] = 0
    const double* gam;    // binom(2n,n)/4^n
    int nS;
    const int4* blocks;   // (first l1, d_lo, end of the row band the tile belongs to, -), heaviest first
    const double* Wp;     // [row j][v2_nqp columns] = (2j+1) W_q[j] / 4pi, zero rows past lenW
};

// W'[j][q] for one call (interleaved so a step's window values are one or a few 128-bit loads)
__global__ void v2_prep_w(double* __restrict__ Wp, int rows, int nqp, int nw, int lenW,
                          const double* w0, const double* w1, const double* w2, const double* w3,
                          const double* w4, const double* w5, const double* w6, const double* w7)
{
    const double* W[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    const double k = (double)(2 * j + 1) * INV_4PI;
    for (int q = 0; q < nqp; ++q)
        Wp[(size_t)j * nqp + q] = (q < nw && j < lenW) ? k * W[q][j] : 0.0;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void block_sync()
{
    __syncwarp();            // one warp per block: staging is warp-private
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Resident warps per SM the register allocation must allow (one warp per block): 16 = 4 per
// SMSP for the light jobs, 12 = 3 per SMSP for the 8-accumulator covariance jobs.
__host__ __device__ constexpr int v2_min_blocks(int job)
{
    if (v2_r(job) > 4) {
        if (v2_family(job) == FAM_00 && job_nacc(job) <= 2) return 16;
        return (v2_r(job) >= 8 || v2_r(job) * job_nacc(job) >= 24) ? 8 : 12;
    }
    return (job == JOB_EEEE || job == JOB_TETE || job == JOB_MASTER) ? 12 : (job == JOB_M00 ? 32 : 16);
}

template <int JOB>
__global__ void __launch_bounds__(V2_THREADS, v2_min_blocks(JOB)) pair_kernel_v2(const PairArgs A, const V2Tables T)
{
    constexpr int FAM = v2_family(JOB);
    constexpr bool E2 = v2_even_only_spin2(JOB);      // even-parity spin-2 through f00^2 * (N/D)^k
    constexpr int NWQ = job_nw(JOB);
    constexpr int NACC = job_nacc(JOB);
    constexpr int NQP = v2_nqp(JOB);
    constexpr int R = v2_r(JOB);
    constexpr int NR = v2_nr(JOB), LPR = 32 / NR;     // rows per warp, lanes per row
    constexpr int SPAN = v2_span(JOB);                // pairs per row per warp
    constexpr int NTAB = v2_ntab(JOB);       // F00: ratio tables only; F22/F02: value + (negated) inverse
    constexpr int DS = (FAM == FAM_00) ? 2 : 1;   // stride of d inside a warp == step of l3
    constexpr int V2_TC = v2_tc(JOB);
    constexpr int SZT = v2_szt(JOB), SUB = v2_sub(JOB), TSTR = v2_tstride(JOB);

    // F22/F02 tables hold (value, inverse) pairs so one 128-bit load fetches both; F00 holds ratios.
    extern __shared__ __align__(16) double smem[];
    double* shU = smem;                                   // falling index: (U, -1/U) | RT, one table per row
    double* shV = shU + NR * TSTR;                        // rising index:  (V, 1/V)  | RV
    double* shW = shV + NR * TSTR;                        // [V2_TC][NQP]   (TSTR is even: 16-byte aligned)
    double* shF = shW + V2_TC * NQP;                      // start values f22(d) | g(d)
    double* shH = shF + 32 * R;                           // start values f00(d)     (F02 only)

    const int4 blk = T.blocks[blockIdx.x];
    const int l1_first = blk.x, d_lo = blk.y, band_hi = blk.z;
    const int tid = threadIdx.x, lane = tid & 31;
    const int rg = lane / LPR;                            // my row group
    const int l1 = l1_first + rg;                         // my row
    const int L = 2 * l1 + 1;
    const int e = (lane % LPR) * R;                       // pair offset of my first pair inside the window
    const int dmax = (l1 < band_hi) ? A.lmax - l1 : -1;  // last valid d of my row (rows past the band: none)
    const int l1_last = min(l1_first + NR, band_hi) - 1; // longest family of the warp
    // last step: the last pair (offset SPAN-1) of the last row finishes its family, or the window spectrum ends
    const int tau_end = (A.lenW - 1 - d_lo < 0) ? -1
                      : min(SPAN - 1 + (2 * l1_last) / DS, (A.lenW - 1 - d_lo) / DS);   // block-uniform
    const double* myU = shU + rg * TSTR;
    const double* myV = shV + rg * TSTR;

    // ---- start values (closed form), one per pair, parked in shared memory ----
    {
        const double gl1 = T.gam[min(l1, A.lmax)];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int d = d_lo + DS * (e + r);
            double g00 = 0.0, g22 = 0.0;
            if (d <= dmax) {
                const int l2 = l1 + d;
                g00 = T.gam[d] * gl1 / (T.gam[l2] * (double)(2 * l2 + 1));
                if (l1 >= 2) {
                    const double num = (double)(l2 + 1) * (double)(l2 + 2) * ((double)(l1 - 1) * (double)l1);
                    const double den = (double)(l2 - 1) * (double)l2 * ((double)(l1 + 1) * (double)(l1 + 2));
                    g22 = g00 * (num / den);
                }
            }
            if constexpr (FAM == FAM_00) shF[tid * R + r] = g00;
            if constexpr (FAM == FAM_22) shF[tid * R + r] = sqrt(g22);
            if constexpr (FAM == FAM_02) { shF[tid * R + r] = sqrt(g22); shH[tid * R + r] = sqrt(g00); }
        }
    }

    // ---- per-pair state ----
    // f = f22(j) (or g = f00(j)^2 for FAM_00); p = a(j) f22(j-1); hv = f00 chain (FAM_02 only):
    // f00(j) on even-parity steps, a(j) f00(j-1) on odd ones.
    double f[R], p[R], hv[R];
    double acc[R][NACC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        f[r] = 0.0; p[r] = 0.0; hv[r] = 0.0;
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[r][q] = 0.0;
    }
    // even-parity spin-2 through f00^2: per-pair constants of N(x) = t (t/2 + 1) - ab, t = x - (a+b)
    double e2_s[R], e2_ab[R];
    if constexpr (E2) {
        const double a = (double)l1 * (double)(l1 + 1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int l2 = l1 + d_lo + DS * (e + r);
            const double b = (double)l2 * (double)(l2 + 1);
            e2_s[r] = a + b;
            e2_ab[r] = a * b;
        }
    }
    // x = j (j+1) of the warp's current l3 (uniform), advanced by 4j+6 per step of 2
    double xj = (double)d_lo * (double)(d_lo + 1);
    double xinc = 4.0 * (double)d_lo + 6.0;
    // rotating windows (2R-1 live entries each)
    double wU0[2 * R - 1], wU1[2 * R - 1], wV0[2 * R - 1], wV1[2 * R - 1];
#pragma unroll
    for (int k = 0; k < 2 * R - 1; ++k) { wU0[k] = 0.0; wU1[k] = 0.0; wV0[k] = 0.0; wV1[k] = 0.0; }

    double k4 = 4.0 * (double)(2 * d_lo + 1);             // 4 (2j+1) at tau = 0   (DS == 1 jobs only)
    (void)k4;
    const bool warp_live = d_lo <= A.lmax - l1_first;     // block-uniform (the first row is the longest in d)

    for (int tau0 = 0; tau0 <= tau_end; tau0 += V2_TC) {
        block_sync();
        // ================= stage this chunk's tables =================
        // falling-index tables, entry idx <-> n = tau0 - SPAN + idx  (n = t+1 of the step that uses it)
        // (one table per row of the warp: g = row group, Lg = 2 (l1_first + g) + 1)
        for (int i = tid; i < NR * SZT; i += V2_THREADS) {
            const int g = i / SZT, idx = i - g * SZT;
            const int Lg = 2 * (l1_first + g) + 1;
            const int pos = (idx % R) * SUB + idx / R;
            if constexpr (FAM == FAM_00) {
                // ratio a(j+1)^2/a(j+2)^2, falling part at n = t+1.  The windows hand a pair the entry
                // nu = t/2 + 1 (the "next step" slot), hence n = 2 nu - 1.
                const int n = 2 * (tau0 - SPAN + idx) - 1;
                double v = 0.0;
                if (n >= 1 && n <= Lg - 2)
                    v = ((double)n * (double)(Lg - n)) * (__ldg(T.INV + n + 1) * __ldg(T.INV + (Lg - n - 1)));
                shU[g * TSTR + pos] = v;
            } else {
                const int n = tau0 - SPAN + idx;
                double u = 0.0, iu = 0.0;
                if (n >= 1 && n <= Lg - 1) {
                    u = __ldg(T.S + n) * __ldg(T.S + (Lg - n));
                    iu = -(__ldg(T.IS + n) * __ldg(T.IS + (Lg - n)));
                }
                reinterpret_cast<double2*>(shU + g * TSTR)[pos] = make_double2(u, iu);
            }
        }
        // rising-index tables, entry idx <-> m' = tau0 + idx + 2 d_lo  (m' = m+1 of the step that uses it)
        for (int i = tid; i < NR * SZT; i += V2_THREADS) {
            const int g = i / SZT, idx = i - g * SZT;
            const int Lg = 2 * (l1_first + g) + 1;
            const int pos = (idx % R) * SUB + idx / R;
            if constexpr (FAM == FAM_00) {
                // rising part at mp = m+1 (m = j + d even); slot mu = m/2 + 1, hence mp = 2 mu - 1
                const int mp = 2 * (tau0 + idx + d_lo) - 1;
                double v = 0.0;
                if (mp >= 1)
                    v = ((double)mp * (double)(mp + Lg)) * (__ldg(T.INV + mp + 1) * __ldg(T.INV + (mp + Lg + 1)));
                shV[g * TSTR + pos] = v;
            } else {
                const int mp = tau0 + idx + 2 * d_lo;
                double v = 0.0, iv = 0.0;
                if (mp >= 1) {
                    v = __ldg(T.S + mp) * __ldg(T.S + (mp + Lg));
                    iv = __ldg(T.IS + mp) * __ldg(T.IS + (mp + Lg));
                }
                reinterpret_cast<double2*>(shV + g * TSTR)[pos] = make_double2(v, iv);
            }
        }
        // W' rows j = d_lo + DS (tau0 + row), row < V2_TC   (16-byte cp.async; rows past lenW are zero)
        {
            const double* src = T.Wp + (size_t)(d_lo + DS * tau0) * NQP;
            constexpr int CPR = NQP / 2;                    // 16-byte pieces per row
            constexpr int NCH = V2_TC * CPR;
            for (int c = tid; c < NCH; c += V2_THREADS) {
                const int row = c / CPR, piece = c % CPR;
                cp_async16(shW + 2 * c, src + (size_t)row * DS * NQP + 2 * piece);
            }
            cp_async_wait_all();
        }
        block_sync();

        if (!warp_live) continue;            // dead warps only help staging

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

        const int tg_end = min(V2_TC, tau_end - tau0 + 1);
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
            const bool inject = ((tau0 + tg) == e);        // this group holds t = 0 of my pairs
#pragma unroll
            for (int s = 0; s < R; ++s) {
                // window spectra of this step: warp-uniform broadcast reads
                double w[NQP];
                {
                    const double* wr = shW + (size_t)(tg + s) * NQP;
                    if constexpr (NWQ == 1) {
                        w[0] = wr[0];
                    } else {
#pragma unroll
                        for (int q = 0; q < NQP; q += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(wr + q);
                            w[q] = v.x; w[q + 1] = v.y;
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const bool even = ((r + s) & 1) == 0;          // parity of l1+l2+l3, static
                    const int kU = s - r + R - 1, kV = r + s;
                    if (r == s) {                                   // t == 0 happens at sub-step s == r
                        if (inject) {
                            if constexpr (FAM == FAM_00) f[r] = shF[tid * R + r];
                            else f[r] = shF[tid * R + r];
                            if constexpr (FAM == FAM_02) hv[r] = shH[tid * R + r];
                        }
                    }
                    if constexpr (FAM == FAM_00) {
                        // f[r] holds g = f00(j)^2; this warp only visits even-parity j
                        if constexpr (!E2) {
#pragma unroll
                            for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(f[r], w[q], acc[r][q]);
                        } else {
                            // f22 = f00 N / D on even parity; the 1/D powers are applied in the epilogue
                            const double t = xj - e2_s[r];
                            const double nn = fma(t, fma(0.5, t, 1.0), -e2_ab[r]);
                            const double gn = f[r] * nn;                    // f00 f22 D
                            if constexpr (JOB == JOB_M02 || JOB == JOB_TEEE) {
#pragma unroll
                                for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(gn, w[q], acc[r][q]);
                            } else if constexpr (JOB == JOB_TETE) {
                                acc[r][0] = fma(gn, w[0], acc[r][0]);
                                acc[r][1] = fma(f[r], w[1], acc[r][1]);
                                acc[r][2] = fma(gn, w[2], acc[r][2]);
                                acc[r][3] = fma(gn, w[3], acc[r][3]);
                                acc[r][4] = fma(gn, w[4], acc[r][4]);
                            } else {                                        // MPP, EEEE, TEEEP: f22^2 D^2
                                const double gnn = gn * nn;
#pragma unroll
                                for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(gnn, w[q], acc[r][q]);
                            }
                        }
                        f[r] *= wU0[kU] * wV0[kV];
                    } else {
                        const double ee = f[r] * f[r];
                        if constexpr (JOB == JOB_MPP) { if (even) acc[r][0] = fma(ee, w[0], acc[r][0]); }
                        else if constexpr (JOB == JOB_MMM) { if (!even) acc[r][0] = fma(ee, w[0], acc[r][0]); }
                        else if constexpr (JOB == JOB_MPPMMM) {
                            if (even) acc[r][0] = fma(ee, w[0], acc[r][0]);
                            else acc[r][1] = fma(ee, w[0], acc[r][1]);
                        } else if constexpr (JOB == JOB_EEEE || JOB == JOB_TEEEP) {
                            if (even) {
#pragma unroll
                                for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(ee, w[q], acc[r][q]);
                            }
                        } else if constexpr (JOB == JOB_M02 || JOB == JOB_TEEE) {
                            if (even) {
                                const double pr = hv[r] * f[r];
#pragma unroll
                                for (int q = 0; q < NWQ; ++q) acc[r][q] = fma(pr, w[q], acc[r][q]);
                            }
                        } else if constexpr (JOB == JOB_MASTER) {
                            if (even) {
                                const double pr = hv[r] * f[r];
                                acc[r][0] = fma(hv[r] * hv[r], w[0], acc[r][0]);
                                acc[r][1] = fma(pr, w[1], acc[r][1]);
                                acc[r][2] = fma(pr, w[2], acc[r][2]);
                                acc[r][3] = fma(ee, w[3], acc[r][3]);
                            } else {
                                acc[r][4] = fma(ee, w[3], acc[r][4]);
                            }
                        } else if constexpr (JOB == JOB_TETE) {
                            if (even) {
                                const double pr = hv[r] * f[r];
                                acc[r][0] = fma(pr, w[0], acc[r][0]);
                                acc[r][1] = fma(hv[r] * hv[r], w[1], acc[r][1]);
                                acc[r][2] = fma(pr, w[2], acc[r][2]);
                                acc[r][3] = fma(pr, w[3], acc[r][3]);
                                acc[r][4] = fma(pr, w[4], acc[r][4]);
                            }
                        }
                        const double an = wU0[kU] * wV0[kV];        // a(j+1)
                        const double ian = wU1[kU] * wV1[kV];       // -1/a(j+1)
                        if constexpr (FAM == FAM_02) {
                            // f00(j+2) = -a(j+1) f00(j) / a(j+2): one factor per step
                            hv[r] *= even ? an : ian;
                        }
                        const double q2 = fma(k4, f[r], p[r]);     // 4(2j+1) f(j) + a(j) f(j-1)
                        p[r] = an * f[r];
                        f[r] = q2 * ian;
                    }
                }
                k4 += 8.0;
                if constexpr (E2) { xj += xinc; xinc += 8.0; }
            }
            // ---- rotate: next group's carried entries ----
#pragma unroll
            for (int k = 0; k < R - 1; ++k) {
                wU0[k] = wU0[k + R]; wV0[k] = wV0[k + R];
                if constexpr (NTAB > 1) { wU1[k] = wU1[k + R]; wV1[k] = wV1[k + R]; }
            }
        }
    }

    // ---- epilogue: one stored value (two for MPPMMM) per valid pair ----
    if (!warp_live) return;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int d = d_lo + DS * (e + r);
        if (d <= dmax) {
            if constexpr (E2) {
                // powers of 1/D; l1 < 2 (|m| > l, true symbol 0): exact zeros, as the recurrence path gives
                const int l2 = l1 + d;
                double id2 = 0.0;
                if (l1 >= 2)
                    id2 = 1.0 / (((double)(l1 - 1) * (double)l1 * ((double)(l1 + 1) * (double)(l1 + 2)))
                                 * ((double)(l2 - 1) * (double)l2 * ((double)(l2 + 1) * (double)(l2 + 2))));
                const double id1 = sqrt(id2);
                if constexpr (JOB == JOB_M02 || JOB == JOB_TEEE) {
#pragma unroll
                    for (int q = 0; q < NACC; ++q) acc[r][q] *= id1;
                } else if constexpr (JOB == JOB_TETE) {
                    acc[r][0] *= id1; acc[r][2] *= id1; acc[r][3] *= id1; acc[r][4] *= id1;
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
int launch_pair_v2(const PairArgs& A, const V2Tables& T, int nblocks, cudaStream_t st)
{
    constexpr int smem = v2_smem_doubles(JOB) * (int)sizeof(double);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pair_kernel_v2<JOB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_done[dev] = true;
    }
    if (nblocks > 0) pair_kernel_v2<JOB><<<nblocks, V2_THREADS, smem, st>>>(A, T);
    return (int)cudaGetLastError();
}

}  // namespace psb
