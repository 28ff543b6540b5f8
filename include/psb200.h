/*
 * psb200.h -- C ABI of libpsb200.so, the B200 (sm_100a) replacement for the
 * Wigner-3j hot path of PowerSpectra.jl.
 *
 * The reference has no FFI seam of its own (pure Julia).  The seam this library
 * fills is the set of Julia inner-loop methods that `mcm` and `coupledcov` call;
 * a Julia shim overrides exactly those with `ccall`s:
 *
 *   psb200_mcm  replaces  inner_mcm00!   /root/reference/src/modecoupling.jl:78-95
 *                         inner_mcm02!   src/modecoupling.jl:99-119
 *                         inner_mcm++!   src/modecoupling.jl:123-139
 *                         inner_mcm--!   src/modecoupling.jl:143-159
 *                         (+ fill_3j! :69-75, Xi_TT/EE/EB/TE :3-66, and the
 *                          WignerFamilies.jl calls WignerF / wigner3j_f! they make)
 *   psb200_cov  replaces  loop_covTTTT!        src/covariance.jl:92-122
 *                         loop_covEEEE!        src/covariance.jl:153-183
 *                         loop_covTTTE!        src/covariance.jl:208-235
 *                         loop_covTETE!        src/covariance.jl:261-302
 *                         loop_covTEEE_planck! src/covariance.jl:376-402
 *                         loop_covTEEE!        src/covariance.jl:337-372
 *                         loop_covTTEE!        src/covariance.jl:422-446
 *
 * Conventions
 *   - Every `double*` of the host-level calls is caller-owned HOST memory, alive for
 *     the duration of the (blocking) call; nothing is retained after return.
 *   - Vectors are 0-based in l: x[l], l = 0..len-1 (a Julia SpectralVector over 0:len-1).
 *   - Matrices are column-major with leading dimension ld >= N, N = lmax-lmin+1:
 *     element (l1,l2) lives at A[(l1-lmin) + (l2-lmin)*ld]   (= parent(SpectralArray)).
 *     Both triangles are written:  M[l1,l2] = (2 l2+1) Xi, M[l2,l1] = (2 l1+1) Xi
 *     (src/modecoupling.jl:90-91);  C[l2,l1] = C[l1,l2]  (src/covariance.jl:119).
 *   - The l3 sum runs over [|l1-l2|, min(l1+l2, len-1)] of the window vector
 *     (src/modecoupling.jl:6-8); lmin only crops rows/columns.
 *   - Return value: 0 ok; 1 bad argument (maps to ArgumentError / the @assert at
 *     src/modecoupling.jl:80,225); 2 CUDA error; 3 collective error; 4 out of memory;
 *     5 no usable CUDA device; 6 singular matrix in a device-side solve (Julia: SingularException).  psb200_last_error() gives the message.  There is
 *     NO CPU fallback: without a B200-class device every compute call fails with 5.
 *   - Thread safety: calls may come from any OS thread; they are serialised inside.
 */
#ifndef PSB200_H
#define PSB200_H
#include <stddef.h>   /* size_t */

#ifdef __cplusplus
extern "C" {
#endif

/* psb200_mcm kinds */
enum {
    PSB200_M00 = 0,      /* :TT, :M00                         inner_mcm00! */
    PSB200_M02 = 1,      /* :TE :ET :TB :BT :M02 :M20         inner_mcm02! */
    PSB200_MPP = 2,      /* :M++                              inner_mcm++! */
    PSB200_MMM = 3,      /* :M--                              inner_mcm--! */
    PSB200_MPP_MMM = 4   /* both from ONE evaluation of the (0,-2,2) family: M -> M++, M2 -> M-- */
};

/* psb200_cov blocks (positional order of spectra / ratios / W = the reference signatures) */
enum {
    PSB200_TTTT = 0,        /* sp TTip TTjq TTiq TTjp | r ip jq iq jp | W1..W8 */
    PSB200_EEEE = 1,        /* sp EEip EEjq EEiq EEjp | r ip jq iq jp | W1..W8 */
    PSB200_TTTE = 2,        /* sp TTip TTjp TEiq TEjq | r ip jp       | W1..W4 */
    PSB200_TETE = 3,        /* sp TTip EEjq TEiq TEjp | r TTip PPjq   | W1..W5 */
    PSB200_TEEE_PLANCK = 4, /* sp EEjq EEjp TEip TEiq | r EEjq EEjp   | W1..W4 */
    PSB200_TEEE = 5,        /* same arguments, f00*f22 instead of f22^2 */
    PSB200_TTEE = 6         /* sp TEip TEiq TEjq TEjp | (no ratios)   | W1 W2  */
};

/* ---- host-level entry points (what the Julia shim ccalls) ------------------------- */

/* Mode-coupling matrix.  V[0..nV-1] = mask cross-spectrum (reference passes nV = lmax+1,
 * src/modecoupling.jl:197).  M is fully overwritten; M2 only for kind 4 (else NULL).
 * ngpus: 1, 2, 4, 8 ... row bands across that many devices of this box; 0 = all visible. */
int psb200_mcm(int kind, int lmin, int lmax, const double* V, int nV,
               double* M, long ldM, double* M2, int ngpus);

/* Optional fused call (not a reference method): everything `master` /
 * `maskedalm2spectra` asks of `mcm` (src/modecoupling.jl:348-362 -- TT, TE(=TB), ET(=BT),
 * (EE_BB, EB_BE)) from ONE evaluation of the two families per pair:
 *   M00    = inner_mcm00!(V_TT)      M02_TP = inner_mcm02!(V_TP)     M02_PT = inner_mcm02!(V_PT)
 *   Mpp    = inner_mcm++!(V_PP)      Mmm    = inner_mcm--!(V_PP)
 * V_XY = alm2cl(mask X of map 1, mask Y of map 2)[0..nV-1].  All five outputs N x N, leading dim ldM. */
int psb200_mcm_master(int lmin, int lmax, const double* V_TT, const double* V_TP, const double* V_PT,
                      const double* V_PP, int nV, double* M00, double* M02_TP, double* M02_PT,
                      double* Mpp, double* Mmm, long ldM, int ngpus);

/* Coupled covariance block.  spectra[k], ratios[k]: length >= lmax+1; W[k]: length lenW
 * (reference: workspace.lmax+1, src/workspace.jl:197).  nspec/nratio/nW must equal the
 * block's arity: TTTT/EEEE 4/4/8, TTTE 4/2/4, TETE 4/2/5, TEEE* 4/2/4, TTEE 4/0/2. */
int psb200_cov(int block, int lmin, int lmax,
               const double* const* spectra, int nspec,
               const double* const* ratios, int nratio,
               const double* const* W, int nW, int lenW,
               double* C, long ldC, int ngpus);

const char* psb200_last_error(void);   /* message of the last non-zero return on this thread */
int psb200_device_count(void);         /* CUDA devices visible to the library (0 if none) */
const char* psb200_version(void);

/* ---- device-level entry points ------------------------------------------------------
 * Same computations on buffers already resident in HBM of the CURRENT device; used by the
 * one-process-per-GPU driver (bench.py under torchrun) and by the kernel-only timings.
 * All pointers are DEVICE pointers; `stream` is a cudaStream_t (NULL = default stream).
 * Calls are asynchronous with respect to the host.
 *
 * Stage 1 writes raw Xi for the rows l1 in [row_lo, row_hi) (absolute l, lmin <= row_lo):
 *     X[(l1-lmin)*ldX + (l2-lmin)],  l2 = l1..lmax      (other entries untouched)
 * i.e. row l1 of the upper triangle is contiguous -- it is column l1 of the final
 * column-major matrix, so a band of rows is ONE contiguous slab that can be sent to
 * rank 0 as is.  For MCM kinds X holds Xi; for covariance blocks X holds C[l1,l2].
 * Stage 2 (psb200_finish_dev, on the rank that owns the whole matrix) fills both
 * triangles in place: scale = 1 applies the (2l+1) factors of the MCM, scale = 0 copies.
 */
int psb200_mcm_dev(int kind, int lmin, int lmax, const double* dV, int nV,
                   double* dX, long ldX, double* dX2,
                   int row_lo, int row_hi, void* stream);

/* dX: host array of the five device outputs in the order M00, M02_TP, M02_PT, Mpp, Mmm */
int psb200_mcm_master_dev(int lmin, int lmax, const double* dV_TT, const double* dV_TP, const double* dV_PT,
                          const double* dV_PP, int nV, double* const* dX, long ldX,
                          int row_lo, int row_hi, void* stream);

int psb200_cov_dev(int block, int lmin, int lmax,
                   const double* const* d_spectra, int nspec,   /* host array of device ptrs */
                   const double* const* d_ratios, int nratio,
                   const double* const* d_W, int nW, int lenW,
                   double* dX, long ldX,
                   int row_lo, int row_hi, void* stream);

/* The same over SEVERAL disjoint row bands in ONE launch: bands[2k], bands[2k+1] = [lo, hi) of band k, 1 <= nbands <= 4.
 * What a rank of the folded multi-GPU split passes -- a low band (many short l3 families) and a high band (few long
 * ones) merged into one heaviest-first tile list, so that every rank fills the tail of its long tiles with short ones. */
int psb200_mcm_dev_bands(int kind, int lmin, int lmax, const double* dV, int nV,
                         double* dX, long ldX, double* dX2,
                         const int* bands, int nbands, void* stream);
int psb200_cov_dev_bands(int block, int lmin, int lmax,
                         const double* const* d_spectra, int nspec,
                         const double* const* d_ratios, int nratio,
                         const double* const* d_W, int nW, int lenW,
                         double* dX, long ldX,
                         const int* bands, int nbands, void* stream);

int psb200_finish_dev(double* dX, long ldX, int lmin, int lmax, int scale, void* stream);

/* Work-balanced contiguous l1 bands: edges[0..nbands] with edges[0] = lmin, edges[nbands] = lmax+1.
 * Row cost = the l3 steps the kernel runs for that row: families truncated at the window length
 * lenW (pass the nV / lenW of the call); lenW <= 0 balances the reference's full-family count
 * (2 l1+1)(lmax-l1+1) instead. */
int psb200_band_edges(int lmin, int lmax, int lenW, int nbands, int* edges);

/* Bands of the HOST-level calls (psb200_mcm / psb200_cov / psb200_mcm_master with ngpus > 1; api/code as in
 * psb200_job_stats).  There every device also copies its own L-shaped region of the result to the caller's array, and
 * kernel time and copy time do not balance alike (low rows: cheap kernels, long columns), so these edges minimise the
 * largest max(kernel seconds, copy seconds) over the bands instead of the kernel cost alone. */
int psb200_host_band_edges(int api, int code, int lmin, int lmax, int lenW, int nbands, int* edges);

/* ---- QuickPol Xi matrix -------------------------------------------------
 * Replaces the pair loop of quickpolXi! (/root/reference/src/beam.jl:72-101) with Xisum (:16-28)
 * and the WignerF / wigner3j_f! calls it makes (:86-93):
 *   Xi[l'', l] = (-1)^(s1+s2+nu1+nu2) sum_{l'} W[l'] (l' l l''; s1+nu1, -s1, -nu1) (l' l l''; s2+nu2, -s2, -nu2)
 * for l'' = 2..lmax and l = max(2, l''-band_lo) .. min(lmax, l''+band_hi)   (specrowrange, :59-63).
 * W[0..lenW-1] = quickpolW(omega1, omega2) (:43-56; stays on the host).  Terms with l' > lenW-1 are
 * dropped (the reference reads W under @inbounds there); entries with |s| > l or |nu| > l'' are 0.
 * Supported domain: lmax <= 12287 (larger is rejected: the rescaling cadence of the sweep is analysed up to there);
 * validated spins |s| <= 4, |nu| <= 12 (l <= 5000) and single large spins up to |s| ~ l (rescaling path): worst 1 % of the
 * parity bound.  Pairs of families with BOTH |s| and |nu| of order l whose classical regions do not overlap are outside
 * it (absolute errors up to 3e-9 found by fuzzing) -- no use of the reference comes near (spins of CMB beams are <= 4).
 * Xb is the storage of the reference's BandedMatrix, parent(Xi).data (BandedMatrices.bandeddata): column-major
 * (band_lo+band_hi+1) x (lmax+1) with leading dimension ldb,
 *   Xi[l'', l]  at  Xb[(band_hi + l'' - l) + l*ldb].
 * Only the entries the reference loop visits are written; the caller applies the reference's final
 * `Xi .*= sgn` (:98-99) to the others if they are non-zero (the visited ones already carry it).
 * Columns are split over `ngpus` devices in cost-balanced bands; no exchange is needed. */
int psb200_quickpol_xi(int nu1, int nu2, int s1, int s2, int lmax, const double* W, int lenW,
                       int band_lo, int band_hi, double* Xb, long ldb, int ngpus);

/* Device-level form: dW, dXb are device pointers, columns l in [col_lo, col_hi) only, asynchronous. */
int psb200_quickpol_xi_dev(int nu1, int nu2, int s1, int s2, int lmax, const double* dW, int lenW,
                           int band_lo, int band_hi, double* dXb, long ldb,
                           int col_lo, int col_hi, void* stream);

/* Cost-balanced contiguous column bands of the Xi matrix: edges[0] = 0 .. edges[nbands] = lmax+1. */
int psb200_quickpol_edges(int lmax, int band_lo, int band_hi, int nbands, int* edges);

/* ---- decoupling on the device (optional, the reference's host solves keep working) ----------------
 * The mode-coupling matrix is computed on `ngpus` devices, assembled on ONE of them over NVLink (the pair kernels of
 * the other devices store their row bands straight into that device's memory), LU-factorised there (cuSOLVER getrf,
 * partial pivoting like the LAPACK call behind Julia's `lu`) and only the decoupled spectra return to the host:
 * the N^2 x 8 B matrix never crosses PCIe.
 *
 * psb200_mcm_solve replaces  `M = mcm(spec, ...); Cl = M \ pCl`   src/modecoupling.jl:359-362, src/blockspectralmatrix.jl:124-129
 *   system 0..3: the N x N matrix of psb200_mcm kind 0..3;  pCl, Cl: N x nrhs column-major (ldp, ldc >= N)
 *   system 4: [M++ M--; M-- M++] \ [pCl_EE; pCl_BB]   (M_EE_BB, src/modecoupling.jl:213-216, :365-371)
 *   system 5: [M++ -M--; -M-- M++] \ [pCl_EB; pCl_BE] (M_EB_BE, :220-223, :373-379);  pCl, Cl: 2N x nrhs (ld >= 2N)
 *   The block systems are block-circulant and split exactly into the two N x N systems of M++ + M-- and M++ - M--
 *   (8x fewer flops than the LU of the dense 2N x 2N hvcat that src/blockspectralmatrix.jl:89-122 factorises; same
 *   solution up to rounding). */
enum { PSB200_SYS_M00 = 0, PSB200_SYS_M02 = 1, PSB200_SYS_MPP = 2, PSB200_SYS_MMM = 3, PSB200_SYS_EE_BB = 4, PSB200_SYS_EB_BE = 5 };
int psb200_mcm_solve(int system, int lmin, int lmax, const double* V, int nV,
                     const double* pCl, long ldp, int nrhs, double* Cl, long ldc, int ngpus);

/* Everything maskedalm2spectra solves (src/modecoupling.jl:341-377) from ONE fused evaluation of the five matrices
 * (psb200_mcm_master) without moving any of them: pCl and Cl are N x 9 column-major, columns in the order
 *   0 TT  1 TE  2 ET  3 TB  4 BT  5 EE  6 BB  7 EB  8 BE
 * TT = M00(V_TT) \ pTT;  TE, TB = M02(V_TP) \ .;  ET, BT = M02(V_PT) \ .;  [EE; BB] = M_EE_BB \ [pEE; pBB];
 * [EB; BE] = M_EB_BE \ [pEB; pBE]. */
int psb200_master_solve(int lmin, int lmax, const double* V_TT, const double* V_TP, const double* V_PT,
                        const double* V_PP, int nV, const double* pCl, long ldp, double* Cl, long ldc, int ngpus);

/* decouple_covmat (src/covariance.jl:8-14): out = B1^-1 Y (B2^-1)^T through lu(B1'), lu(B2') as the reference does.
 * Host form: n x n column-major host matrices, computed on the current device.  Device form: device pointers, Y is
 * overwritten in place (B1, B2 untouched), asynchronous on `stream` except for the pivot-singularity check. */
int psb200_decouple_covmat(int n, const double* Y, long ldy, const double* B1, long ldb1, const double* B2, long ldb2,
                           double* out, long ldo);
int psb200_decouple_covmat_dev(int n, double* dY, long ldy, const double* dB1, long ldb1, const double* dB2, long ldb2,
                               void* stream);

/* ---- W-spectrum production, first slice -------------------------------------------------
 * The window spectra W of the covariance come from map2alm of mask products (effective_weight_alm!,
 * src/workspace.jl:141-171) and alm2cl of pairs of them (window_function_W!, :174-213).  The general HEALPix transform
 * stays on the host; for AZIMUTHALLY SYMMETRIC maps only m = 0 survives and map2alm is the Legendre quadrature
 *     alm[i][l] = sqrt(pi (2l+1)) sum_k w[k] fields[i][k] P_l(x[k]),   l = 0..lmax
 * over Gauss-Legendre nodes x (cos theta) and weights w, evaluated here for a batch of fields (rows of `fields`, leading
 * dimension ldf >= nnodes; rows of `alm`, leading dimension lda >= lmax+1) on the current device.  alm2cl of two zonal
 * maps is a_l0 b_l0 / (2l+1). */
int psb200_zonal_alm(int nfields, int nnodes, const double* x, const double* w, const double* fields, long ldf,
                     int lmax, double* alm, long lda);

/* ---- W-spectrum production: spin-0 HEALPix transforms ----------------------------------
 * Replaces, optionally, the Healpix.jl calls behind the window spectra:
 *   effective_weight_alm!  src/workspace.jl:141-171   map2alm(mask_i .* mask_j [.* sigma^2 .* Omega_pix]; lmax)   (:153-155, :159-163)
 *   window_function_W!     src/workspace.jl:174-213   alm2cl(w_X, w_Y)[0:lmax]                                      (:202-204)
 *   map2alm(mask) in front of mcm / master            src/modecoupling.jl:250-256, :328-329
 * Maps are HEALPix RING-ordered, 12 nside^2 doubles (parent(HealpixMap{Float64,RingOrder})), nside a power of two
 * <= 2048, lmax <= 4 nside - 1.  alm are complex128 stored as interleaved (re, im) doubles in Healpix.jl's Alm order,
 * mmax = lmax: index(l, m) = m (2 lmax + 1 - m)/2 + l, (lmax+1)(lmax+2)/2 coefficients.
 * psb200_map2alm: alm = map2alm(scale * factors[0] .* ... .* factors[nfactors-1]; lmax, niter), 1 <= nfactors <= 3 --
 *   uniform pixel weights 4 pi/npix and `niter` Jacobi iterations alm += A(map - S alm) (Healpix.jl default niter = 3).
 * psb200_alm2map: map = alm2map(alm, nside).   psb200_alm2cl: cl[l] = [a_l0 b_l0 + 2 sum_{m>0} Re(a_lm conj b_lm)]/(2l+1),
 * l = 0..lmax (alm1 == alm2 allowed).  Host forms take host buffers and run on the current device; the _dev forms take
 * device buffers and are asynchronous on `stream` (dmap is not modified).  No CPU fallback. */
int psb200_map2alm(int nside, int lmax, int niter, int nfactors, const double* const* factors, double scale, double* alm);
/* Every effective weight of a workspace in one call: alm[k] = map2alm(scale[k] * maps[idx[3k]] .* maps[idx[3k+1]] .* maps[idx[3k+2]])
 * for k < nprod (idx entries -1 = no factor; the first must be given).  The nmaps unique host maps are uploaded ONCE per
 * device instead of once per product; ngpus > 1 deals the products round-robin to that many devices (the transforms of
 * different products are independent: no exchange), 0 = all visible, 1 = the current device. */
int psb200_map2alm_many(int nside, int lmax, int niter, int nmaps, const double* const* maps, int nprod, const int* idx,
                        const double* scale, double* const* alm, int ngpus);
int psb200_alm2map(int nside, int lmax, const double* alm, double* map);
int psb200_alm2cl(int lmax, const double* alm1, const double* alm2, double* cl);
int psb200_map2alm_dev(int nside, int lmax, int niter, const void* dmap, void* dalm, void* stream);
int psb200_alm2map_dev(int nside, int lmax, const void* dalm, void* dmap, void* stream);
int psb200_alm2cl_dev(int lmax, const void* dalm1, const void* dalm2, void* dcl, void* stream);
/* Work accounting of ONE Legendre pass (analysis or synthesis; map2alm with niter iterations runs 2 niter + 1) as the
 * kernel tiles it (host arithmetic): out[0] executed (l, m, ring pair) steps, out[1] steps of the rings the transform
 * starts, out[2] warps, out[3] ring pairs per lane, out[4] ring-pair chunks, out[5] l steps per pass (out: 6 entries).
 * 4 FP64 instructions per step (2 of the un-normalised recurrence + 2 accumulate). */
int psb200_sht_stats(int nside, int lmax, long long* out);
/* The transforms keep their tables and work buffers per device between calls (one (nside, lmax) at a time; 12 GB at
 * nside 2048, lmax 6143); this frees them on every device. */
int psb200_sht_release(void);

/* 3j terms (full families, as the reference evaluates them) of one call on rows [row_lo,row_hi). */
long long psb200_terms(int families, int lmax, int row_lo, int row_hi);

/* Work accounting of ONE pair-kernel launch as the tuned kernel tiles it (host arithmetic, no device needed);
 * what bench.py turns into the executed-work roofline fraction.  api/code: 0/kind = psb200_mcm kinds,
 * 1/block = psb200_cov blocks, 2/0 = psb200_mcm_master.  lenW = window length of the call.
 *   out[0] executed pair-steps: lockstep l3 steps of every warp x its 32 R pair slots (dead slots included)
 *   out[1] live pair-steps: every pair over its own l3 range (truncated at lenW-1, one parity where the job
 *          steps l3 by 2) -- the part of out[0] that is needed
 *   out[2] warps launched   out[3] pairs per thread R   out[4] rows per warp   out[5] l3 stride
 * Rows l1 < 2 of the spin-2 jobs are evaluated by a separate tiny kernel and not counted. */
int psb200_job_stats(int api, int code, int lmax, int lenW, int row_lo, int row_hi, long long* out);

/* FP64 pipe microbenchmark: dependent-free DFMA streams on every SM for `iters` iterations;
 * returns achieved FLOP/s (2 per DFMA) on the current device, <0 on error. */
double psb200_dfma_peak(int iters);

/* ---- host result buffers -------------------------------------------------------------------------------
 * The reference allocates its result arrays itself (spectralzeros, src/modecoupling.jl:199 etc. -- pageable memory, which
 * the host calls accept as they are).  A caller that owns the allocation gets the full PCIe rate from a page-locked
 * array, and on a multi-socket box with several GPUs from one whose pages are spread over the sockets: every GPU writes
 * its own region of the result by DMA, and an array that lives on one socket makes half the GPUs write across the
 * socket link.
 * psb200_host_alloc: page-locked (cudaHostRegister, portable), zero-filled, 2 MB aligned.  policy 0 = pages on the NUMA
 *   node of the calling thread; policy 1 = 2 MB pieces alternating between the NUMA nodes the process may run on
 *   (mbind(MPOL_INTERLEAVE), or first touch by node-bound threads where the container filters mbind; same as 0 on a
 *   one-node host).  NULL on failure (psb200_last_error).  On a host without a CUDA device the memory is returned
 *   unregistered (placement stays testable); the compute calls still fail with code 5 there.
 * psb200_host_free: releases such a block (0, or 1 for a pointer that is not one).
 * psb200_host_placement: counts[k] = sampled pieces of the block found on node k < maxnodes; returns the number of
 *   samples, 0 when the kernel does not tell (get_mempolicy filtered), -1 for a bad pointer.
 * psb200_host_numa_nodes: nodes policy 1 would use. */
void* psb200_host_alloc(size_t bytes, int policy);
/* PAGEABLE result arrays -- what the reference itself allocates -- are served by a staged delivery inside the host
 * calls: the regions of the result are DMA'd into a ring of page-locked chunks of the library and scattered into the
 * caller's array by a few host threads, instead of the CUDA runtime's one-thread bounce copy.  Environment:
 * PSB200_STAGED=0 turns it off, PSB200_STAGE_THREADS (default min(8, cores / (2 ngpus))) and PSB200_STAGE_CHUNK_MB
 * (default 8) tune it.  PSB200_MIRROR=1 (off by default) selects the mirror delivery: only the block columns of the
 * result cross PCIe and the scatter threads write the symmetric side from the same bytes (same IEEE products as the
 * device: bit-identical) -- for hosts whose DMA ingest rate, not the GPUs, bounds a several-GPU call.
 * psb200_selftest_delivery is the test hook of all three (mode 0 direct, 1 staged, 2 mirror): host memory only, no
 * device (csrc/psb200.cu). */
int psb200_selftest_delivery(int lmin, int lmax, int a, int b, int nsub, int nout, int mode, int scale, int chunk_kb, int nch,
                             int nthreads, double* const* out, long ldo);
int psb200_host_free(void* p);
int psb200_host_placement(const void* p, int* counts, int maxnodes);
int psb200_host_numa_nodes(void);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_H */
