// psb200_sht.inl -- host side of the spin-0 HEALPix transforms (psb200_sht.cuh).
// Included by psb200.cu inside its anonymous namespace.

struct ShtPlan {                      // per device and (nside, lmax): tables + work buffers, kept between calls
    int nside = 0, lmax = 0, R = 0, C = 0;
    psb::ShtDims D{};
    double2* coef = nullptr;          // recurrence coefficients (d_j, Q_{j+1}), sht_coef_base layout
    double* cm = nullptr;             // log2 |lambda_mm| prefactors
    int* cmin = nullptr;              // first active chunk per m
    double4* Phi = nullptr;           // ring-pair phases, (lmax+1) x nrp x 32 B
    double* partial = nullptr;        // per-chunk partial alm of the analysis
    double* resid = nullptr;          // residual map of the Jacobi iterations
    double* work = nullptr;           // product map / host-call staging
    double* alm = nullptr;            // host-call staging
    double* alm2 = nullptr;
    size_t cap_work = 0;
};
ShtPlan g_sht[16];

void sht_free(ShtPlan& P)
{
    cudaFree(P.coef); cudaFree(P.cm); cudaFree(P.cmin); cudaFree(P.Phi); cudaFree(P.partial); cudaFree(P.resid);
    cudaFree(P.work); cudaFree(P.alm); cudaFree(P.alm2);
    P = ShtPlan{};
}

int sht_R()
{
    const char* e = getenv("PSB200_SHT_R");
    const int r = e ? atoi(e) : 4;
    return (r == 2 || r == 4 || r == 8) ? r : 4;
}

int sht_C() { return psb::SHT_C; }

// l steps per reduction of the analysis kernel: 16 (default), or 8 = two reductions per pass with 32 registers less and a
// fifth block per SM -- measured 30.5 vs 29.8 ms per analysis pass at nside 2048: the extra warps buy nothing
int sht_V()
{
    const char* e = getenv("PSB200_SHT_V");
    return (e && atoi(e) == 8) ? 8 : 16;
}

int sht_check(int nside, int lmax)
{
    if (nside < 1 || nside > 2048 || (nside & (nside - 1)))
        return fail(ERR_ARG, "sht: nside must be a power of two in [1, 2048] (got %d)", nside);
    if (lmax < 0 || lmax > 4 * nside - 1) return fail(ERR_ARG, "sht: lmax must lie in [0, 4 nside - 1] (got %d)", lmax);
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    return OK;
}

int sht_plan(int dev, int nside, int lmax, cudaStream_t st, ShtPlan** out)
{
    ShtPlan& P = g_sht[dev];
    const int R = sht_R();
    if (P.nside == nside && P.lmax == lmax && P.R == R) { P.C = sht_C(); *out = &P; return OK; }
    CUDA_TRY(cudaDeviceSynchronize());
    sht_free(P);
    psb::ShtDims D;
    D.nside = nside; D.lmax = lmax; D.nrp = 2 * nside;
    D.nchunks = (D.nrp + 32 * R - 1) / (32 * R);
    D.npix = 12LL * nside * nside;
    D.nalm = (long long)(lmax + 1) * (lmax + 2) / 2;
    CUDA_TRY(cudaMalloc(&P.coef, (size_t)psb::sht_coef_size(lmax) * sizeof(double2)));
    CUDA_TRY(cudaMalloc(&P.cm, (size_t)(lmax + 1) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&P.cmin, (size_t)(lmax + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&P.Phi, (size_t)(lmax + 1) * D.nrp * sizeof(double4)));
    CUDA_TRY(cudaMalloc(&P.partial, (size_t)D.nchunks * 2 * D.nalm * sizeof(double)));
    CUDA_TRY(cudaMalloc(&P.resid, (size_t)D.npix * sizeof(double)));
    CUDA_TRY(cudaMalloc(&P.work, (size_t)D.npix * sizeof(double)));
    CUDA_TRY(cudaMalloc(&P.alm, (size_t)2 * D.nalm * sizeof(double)));
    CUDA_TRY(cudaMalloc(&P.alm2, (size_t)2 * D.nalm * sizeof(double)));
    // log2 |lambda_mm| = 1/2 log2[(2m+1)/(4 pi) prod_{k<=m} (2k-1)/(2k)], long double on the host
    std::vector<double> cm(lmax + 1);
    long double acc = 0.0L;
    for (int m = 0; m <= lmax; ++m) {
        if (m) acc += log2l((2.0L * m - 1.0L) / (2.0L * m));
        cm[m] = (double)(0.5L * (log2l((2.0L * m + 1.0L) / (4.0L * 3.14159265358979323846264338327950288L)) + acc));
    }
    CUDA_TRY(cudaMemcpyAsync(P.cm, cm.data(), cm.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));           // cm is a local
    psb::sht_coef_kernel<<<dim3((unsigned)((lmax / psb::SHT_C + 64) / 64), (unsigned)(lmax + 1)), 64, 0, st>>>(lmax, P.coef);
    CUDA_TRY(cudaGetLastError());
    psb::sht_cmin_kernel<<<(lmax + 128) / 128, 128, 0, st>>>(D, R, P.cmin);
    CUDA_TRY(cudaGetLastError());
    const int hmax = 2 * nside;
    const int smem = (3 * hmax + 1 + 512) * (int)sizeof(double2);
    if (smem > 48 * 1024) {
        CUDA_TRY(cudaFuncSetAttribute(psb::sht_ring_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CUDA_TRY(cudaFuncSetAttribute(psb::sht_ring_synthesis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    P.nside = nside; P.lmax = lmax; P.R = R; P.C = sht_C(); P.D = D;
    *out = &P;
    return OK;
}

// ring kernels are launched per size class so that the short polar rings do not pay for the shared memory of the belt
template <class F> int sht_ring_launches(const ShtPlan& P, F launch)
{
    const int N = P.nside, nrp = P.D.nrp;
    const int edges[4] = {0, std::min(nrp, 256), std::min(nrp, 1024), nrp};        // h = 2 min(p+1, N)
    for (int c = 0; c < 3; ++c) {
        const int lo = edges[c], hi = edges[c + 1];
        if (hi <= lo) continue;
        const int hmax = 2 * std::min(hi, N);
        const int threads = hmax >= 2048 ? 512 : hmax >= 512 ? 256 : 64;
        if (int rc = launch(lo, hi - lo, threads, (size_t)(3 * hmax + 1 + threads) * sizeof(double2))) return rc;
    }
    return OK;
}

int sht_analysis(ShtPlan& P, cudaStream_t st, const double* dmap, double* dalm, int accumulate)
{
    const psb::ShtDims D = P.D;
    if (int rc = sht_ring_launches(P, [&](int lo, int cnt, int threads, size_t smem) {
            psb::sht_ring_analysis_kernel<<<dim3((unsigned)cnt, 2), threads, smem, st>>>(D, lo, dmap, (double2*)P.Phi);
            CUDA_TRY(cudaGetLastError());
            return (int)OK;
        })) return rc;
    const int bpm = (D.nchunks + psb::SHT_WARPS - 1) / psb::SHT_WARPS;
    const unsigned grid = (unsigned)((D.lmax + 1) * bpm);
#define PSB_SHT_ANA(RR, VV) psb::sht_leg_analysis_kernel<RR, VV><<<grid, 32 * psb::SHT_WARPS, 0, st>>>(D, P.Phi, P.coef, P.cm, P.cmin, P.partial)
    switch (P.R * 100 + sht_V()) {
        case 208: PSB_SHT_ANA(2, 8); break;
        case 216: PSB_SHT_ANA(2, 16); break;
        case 408: PSB_SHT_ANA(4, 8); break;
        case 808: PSB_SHT_ANA(8, 8); break;
        case 816: PSB_SHT_ANA(8, 16); break;
        default: PSB_SHT_ANA(4, 16);
    }
#undef PSB_SHT_ANA
    CUDA_TRY(cudaGetLastError());
    psb::sht_analysis_finish_kernel<<<dim3((unsigned)((2 * (D.lmax + 1) + 127) / 128), (unsigned)(D.lmax + 1)), 128, 0, st>>>(
        D, P.partial, P.cmin, accumulate, dalm);
    CUDA_TRY(cudaGetLastError());
    return OK;
}

// map_out = S(alm), or ref - S(alm) when ref is given
int sht_synthesis(ShtPlan& P, cudaStream_t st, const double* dalm, const double* ref, double* dmap)
{
    const psb::ShtDims D = P.D;
    const int bpm = (D.nchunks + psb::SHT_WARPS - 1) / psb::SHT_WARPS;
    const unsigned grid = (unsigned)((D.lmax + 1) * bpm);
#define PSB_SHT_SYN(RR) psb::sht_leg_synthesis_kernel<RR><<<grid, 32 * psb::SHT_WARPS, 0, st>>>(D, (const double2*)dalm, P.coef, P.cm, P.cmin, P.Phi)
    switch (P.R) {
        case 2: PSB_SHT_SYN(2); break;
        case 8: PSB_SHT_SYN(8); break;
        default: PSB_SHT_SYN(4);
    }
#undef PSB_SHT_SYN
    CUDA_TRY(cudaGetLastError());
    return sht_ring_launches(P, [&](int lo, int cnt, int threads, size_t smem) {
        psb::sht_ring_synthesis_kernel<<<dim3((unsigned)cnt, 2), threads, smem, st>>>(D, lo, (const double2*)P.Phi, ref, dmap);
        CUDA_TRY(cudaGetLastError());
        return (int)OK;
    });
}

// Healpix.jl map2alm(map; lmax, niter): analysis + niter Jacobi iterations, all on the device
int sht_map2alm(ShtPlan& P, cudaStream_t st, const double* dmap, double* dalm, int niter)
{
    if (int rc = sht_analysis(P, st, dmap, dalm, 0)) return rc;
    for (int it = 0; it < niter; ++it) {
        if (int rc = sht_synthesis(P, st, dalm, dmap, P.resid)) return rc;
        if (int rc = sht_analysis(P, st, P.resid, dalm, 1)) return rc;
    }
    return OK;
}

struct ClScratch { double* p = nullptr; size_t cap = 0; };
ClScratch g_cl[16];                   // per-device partial sums of psb200_alm2cl_dev (its own buffer: the call is asynchronous)

int cl_scratch(int dev, size_t n, double** out)
{
    ClScratch& c = g_cl[dev];
    if (c.cap < n) {
        if (c.p) { CUDA_TRY(cudaDeviceSynchronize()); cudaFree(c.p); c.p = nullptr; c.cap = 0; }
        CUDA_TRY(cudaMalloc(&c.p, n * sizeof(double)));
        c.cap = n;
    }
    *out = c.p;
    return OK;
}

// alm2cl on device buffers; scratch: ceil((lmax+1)/SHT_CL_SEG) x (lmax+1) doubles
int sht_alm2cl(int lmax, const double2* a, const double2* b, double* scratch, double* cl, cudaStream_t st)
{
    const int nseg = (lmax + psb::SHT_CL_SEG) / psb::SHT_CL_SEG;
    psb::sht_alm2cl_partial_kernel<<<dim3((unsigned)((lmax + 128) / 128), (unsigned)nseg), 128, 0, st>>>(lmax, a, b, scratch);
    CUDA_TRY(cudaGetLastError());
    psb::sht_alm2cl_finish_kernel<<<(lmax + 128) / 128, 128, 0, st>>>(lmax, nseg, scratch, cl);
    CUDA_TRY(cudaGetLastError());
    return OK;
}

int sht_enter(int nside, int lmax, int* dev)
{
    if (int rc = sht_check(nside, lmax)) return rc;
    CUDA_TRY(cudaGetDevice(dev));
    if (*dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", *dev);
    return scratch_reserve(*dev, 5, 16);          // creates the per-device streams on first use
}
