// psb200_common.cuh -- shared definitions of the pair kernels (sm_100a only).
//
// Vocabulary: a *pair* is (l1, l2) with l1 <= l2 of the upper triangle; a
// *family* is f(l3) = (l3 l1 l2; 0 m2 m3) for l3 = |l1-l2| .. l1+l2; a *term* is one family
// value; Xi is the l3 reduction of a squared / multiplied family against a window spectrum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace psb {

constexpr double INV_4PI = 0.079577471545947667884441881686257181;

// What one launch computes.  MCM jobs replace inner_mcm*! (/root/reference/src/modecoupling.jl:78-159),
// covariance jobs replace loop_cov*! (src/covariance.jl:92-446).
enum Job : int {
    JOB_M00 = 0, JOB_M02, JOB_MPP, JOB_MMM, JOB_MPPMMM,
    JOB_TTTT, JOB_EEEE, JOB_TTTE, JOB_TETE, JOB_TEEEP, JOB_TEEE, JOB_TTEE,
    JOB_MASTER,      // M00(W0), M02(W1), M02(W2), M++(W3), M--(W3) from ONE evaluation of f00 and f22
    JOB_COUNT
};

// Which families a job needs.
enum Fam : int { FAM_00 = 0, FAM_22 = 1, FAM_02 = 2 };

__host__ __device__ constexpr int job_family(int job)
{
    return (job == JOB_M00 || job == JOB_TTTT || job == JOB_TTTE || job == JOB_TTEE) ? FAM_00
         : (job == JOB_M02 || job == JOB_TETE || job == JOB_TEEE || job == JOB_MASTER) ? FAM_02
                                                                                     : FAM_22;
}
// Number of window spectra the job reads.
__host__ __device__ constexpr int job_nw(int job)
{
    return (job == JOB_TTTT || job == JOB_EEEE) ? 8
         : (job == JOB_TTTE || job == JOB_TEEEP || job == JOB_TEEE || job == JOB_MASTER) ? 4
         : (job == JOB_TETE) ? 5
         : (job == JOB_TTEE) ? 2 : 1;
}
// Number of Xi accumulators the job carries (MPPMMM: one W, two parities).
__host__ __device__ constexpr int job_nacc(int job)
{
    return job == JOB_MPPMMM ? 2 : (job == JOB_MASTER ? 5 : job_nw(job));
}
// Reference-counted families per pair (EE_BB counts the (0,-2,2) family twice).
__host__ __device__ constexpr int job_ref_families(int job)
{
    // MASTER: what `master` needs as distinct reference calls: TT 1 + TE 2 + ET 2 + EE_BB 2
    return job == JOB_MASTER ? 7 : (job == JOB_M02 || job == JOB_TETE || job == JOB_TEEE || job == JOB_MPPMMM) ? 2 : 1;
}

struct PairArgs {
    int lmin, lmax;        // rows / columns kept
    int lenW;              // valid length of every window vector (l3 sum stops at lenW-1)
    int row_lo, row_hi;    // l1 band [row_lo, row_hi)
    int nxb;               // further bands of the same launch (host side only: the tile list carries each tile's band end)
    int xb[6];             // [lo, hi) of up to three further bands
    long ld;               // leading dimension of the outputs
    const double* W[8];    // window spectra, 0-based in l3
    const double* sp[4];   // signal spectra, 0-based in l
    const double* rt[4];   // noise ratios
    double* out0;          // X[(l1-lmin)*ld + (l2-lmin)]
    double* out1;          // second output (MPPMMM, MASTER)
    double* out2;          // MASTER only
    double* out3;
    double* out4;
};

// Epilogue: combine the Xi of one pair into the stored value.  x[] already holds
// Xi_k = (1/4pi) sum (2 l3+1) w(l3) W_k(l3).  MCM jobs store raw Xi (the (2l+1) factors are
// applied by finish_kernel); covariance jobs store C[l1,l2] following, term by term,
// src/covariance.jl:110-118 (TTTT), :171-179 (EEEE), :226-231 (TTTE), :292-297 (TETE),
// :361-366 / :392-397 (TEEE), :438-441 (TTEE).
template <int JOB>
__device__ __forceinline__ void epilogue(const PairArgs& A, int l1, int l2, const double* x)
{
    const long o = (long)(l1 - A.lmin) * A.ld + (l2 - A.lmin);
#define SP(k, l) __ldg(A.sp[k] + (l))
#define RT(k, l) __ldg(A.rt[k] + (l))
    if constexpr (JOB == JOB_M00 || JOB == JOB_M02 || JOB == JOB_MPP || JOB == JOB_MMM) {
        A.out0[o] = x[0];
    } else if constexpr (JOB == JOB_MPPMMM) {
        A.out0[o] = x[0];
        A.out1[o] = x[1];
    } else if constexpr (JOB == JOB_MASTER) {
        A.out0[o] = x[0];
        A.out1[o] = x[1];
        A.out2[o] = x[2];
        A.out3[o] = x[3];
        A.out4[o] = x[4];
    } else if constexpr (JOB == JOB_TTTT || JOB == JOB_EEEE) {
        const double s0a = SP(0, l1), s0b = SP(0, l2), s1a = SP(1, l1), s1b = SP(1, l2);
        const double s2a = SP(2, l1), s2b = SP(2, l2), s3a = SP(3, l1), s3b = SP(3, l2);
        const double r0a = RT(0, l1), r0b = RT(0, l2), r1a = RT(1, l1), r1b = RT(1, l2);
        const double r2a = RT(2, l1), r2b = RT(2, l2), r3a = RT(3, l1), r3b = RT(3, l2);
        double c = sqrt(s0a * s0b * s1a * s1b) * x[0];
        c += sqrt(s2a * s2b * s3a * s3b) * x[1];
        c += sqrt(s0a * s0b) * x[2] * r1a * r1b;
        c += sqrt(s1a * s1b) * x[3] * r0a * r0b;
        c += sqrt(s2a * s2b) * x[4] * r3a * r3b;
        c += sqrt(s3a * s3b) * x[5] * r2a * r2b;
        c += x[6] * r0a * r1a * r0b * r1b;
        c += x[7] * r2a * r3a * r2b * r3b;
        A.out0[o] = c;
    } else if constexpr (JOB == JOB_TTTE) {
        const double te_jq = SP(3, l1) + SP(3, l2), te_iq = SP(2, l1) + SP(2, l2);
        double c = sqrt(SP(0, l1) * SP(0, l2)) * te_jq * x[0];
        c += sqrt(SP(1, l1) * SP(1, l2)) * te_iq * x[1];
        c += te_jq * x[2] * RT(0, l1) * RT(0, l2);
        c += te_iq * x[3] * RT(1, l1) * RT(1, l2);
        A.out0[o] = c / 2;
    } else if constexpr (JOB == JOB_TETE) {
        const double tt = SP(0, l1) * SP(0, l2), ee = SP(1, l1) * SP(1, l2);
        const double rt = RT(0, l1) * RT(0, l2);
        double c = sqrt(tt * ee) * x[0];
        c += 0.5 * (SP(2, l1) * SP(3, l2) + SP(3, l1) * SP(2, l2)) * x[1];
        c += sqrt(tt) * x[2] * RT(1, l1) * RT(1, l2);
        c += sqrt(ee) * x[3] * rt;
        c += x[4] * RT(0, l1) * RT(0, l2) * RT(1, l1) * RT(1, l2);
        A.out0[o] = c;
    } else if constexpr (JOB == JOB_TEEEP || JOB == JOB_TEEE) {
        const double te_ip = SP(2, l1) + SP(2, l2), te_iq = SP(3, l1) + SP(3, l2);
        double c = sqrt(SP(0, l1) * SP(0, l2)) * te_ip * x[0];
        c += sqrt(SP(1, l1) * SP(1, l2)) * te_iq * x[1];
        c += te_ip * x[2] * RT(0, l1) * RT(0, l2);
        c += te_iq * x[3] * RT(1, l1) * RT(1, l2);
        A.out0[o] = c / 2;
    } else if constexpr (JOB == JOB_TTEE) {
        double c = (SP(0, l1) * SP(2, l2) + SP(2, l1) * SP(0, l2)) * x[0];
        c += (SP(1, l1) * SP(3, l2) + SP(3, l1) * SP(1, l2)) * x[1];
        A.out0[o] = c / 2;
    }
#undef SP
#undef RT
}

}  // namespace psb
