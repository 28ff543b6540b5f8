// psb200_sht_abi.inl -- C entry points of the spin-0 HEALPix transforms (include/psb200.h, "W-spectrum production").
// Included by psb200.cu inside its extern "C" block.

int psb200_map2alm_dev(int nside, int lmax, int niter, const void* dmap, void* dalm, void* stream)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!dmap || !dalm || niter < 0) return fail(ERR_ARG, "map2alm: bad arguments");
    int dev = 0;
    if (int rc = sht_enter(nside, lmax, &dev)) return rc;
    ShtPlan* P = nullptr;
    if (int rc = sht_plan(dev, nside, lmax, (cudaStream_t)stream, &P)) return rc;
    return sht_map2alm(*P, (cudaStream_t)stream, (const double*)dmap, (double*)dalm, niter);
}

int psb200_alm2map_dev(int nside, int lmax, const void* dalm, void* dmap, void* stream)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!dmap || !dalm) return fail(ERR_ARG, "alm2map: bad arguments");
    int dev = 0;
    if (int rc = sht_enter(nside, lmax, &dev)) return rc;
    ShtPlan* P = nullptr;
    if (int rc = sht_plan(dev, nside, lmax, (cudaStream_t)stream, &P)) return rc;
    return sht_synthesis(*P, (cudaStream_t)stream, (const double*)dalm, nullptr, (double*)dmap);
}

int psb200_alm2cl_dev(int lmax, const void* dalm1, const void* dalm2, void* dcl, void* stream)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (lmax < 0 || lmax > 8191 || !dalm1 || !dalm2 || !dcl) return fail(ERR_ARG, "alm2cl: bad arguments");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", dev);
    const size_t nseg = (size_t)(lmax + psb::SHT_CL_SEG) / psb::SHT_CL_SEG;
    double* part = nullptr;
    if (int rc = cl_scratch(dev, nseg * ((size_t)lmax + 1), &part)) return rc;       // calls on one device must be stream-ordered by the caller
    return sht_alm2cl(lmax, (const double2*)dalm1, (const double2*)dalm2, part, (double*)dcl, (cudaStream_t)stream);
}

int psb200_map2alm(int nside, int lmax, int niter, int nfactors, const double* const* factors, double scale, double* alm)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (nfactors < 1 || nfactors > 3 || !factors || !alm || niter < 0) return fail(ERR_ARG, "map2alm: bad arguments");
    for (int i = 0; i < nfactors; ++i)
        if (!factors[i]) return fail(ERR_ARG, "map2alm: null map");
    int dev = 0;
    if (int rc = sht_enter(nside, lmax, &dev)) return rc;
    cudaStream_t st = g_scratch[dev].stream;
    ShtPlan* P = nullptr;
    if (int rc = sht_plan(dev, nside, lmax, st, &P)) return rc;
    const size_t nb = (size_t)P->D.npix * sizeof(double);
    // factors staged in: work (first), resid (second), Phi (third: free until the first analysis)
    double* slot[3] = {P->work, P->resid, (double*)P->Phi};
    if (nfactors == 3 && (size_t)(lmax + 1) * P->D.nrp * sizeof(double4) < nb) {
        if (int rc = scratch_reserve(dev, 0, (size_t)P->D.npix)) return rc;
        slot[2] = g_scratch[dev].X[0];
    }
    for (int i = 0; i < nfactors; ++i)
        if (int rc = upload_async(dev, slot[i], factors[i], nb, st, 1)) return rc;      // pageable maps: staged (psb200.cu "Delivery")
    if (nfactors > 1 || scale != 1.0) {
        psb::sht_product_kernel<<<1184, 256, 0, st>>>(P->D.npix, slot[0], nfactors > 1 ? slot[1] : nullptr,
                                                     nfactors > 2 ? slot[2] : nullptr, scale, P->work);
        CUDA_TRY(cudaGetLastError());
    }
    if (int rc = sht_map2alm(*P, st, P->work, P->alm, niter)) return rc;
    return download_blocking(dev, alm, P->alm, (size_t)2 * P->D.nalm * sizeof(double), st, 1);
}

int psb200_alm2map(int nside, int lmax, const double* alm, double* map)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (!alm || !map) return fail(ERR_ARG, "alm2map: bad arguments");
    int dev = 0;
    if (int rc = sht_enter(nside, lmax, &dev)) return rc;
    cudaStream_t st = g_scratch[dev].stream;
    ShtPlan* P = nullptr;
    if (int rc = sht_plan(dev, nside, lmax, st, &P)) return rc;
    if (int rc = upload_async(dev, P->alm, alm, (size_t)2 * P->D.nalm * sizeof(double), st, 1)) return rc;
    if (int rc = sht_synthesis(*P, st, P->alm, nullptr, P->work)) return rc;
    return download_blocking(dev, map, P->work, (size_t)P->D.npix * sizeof(double), st, 1);
}

int psb200_alm2cl(int lmax, const double* alm1, const double* alm2, double* cl)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (lmax < 0 || lmax > 8191 || !alm1 || !alm2 || !cl) return fail(ERR_ARG, "alm2cl: bad arguments");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", dev);
    const size_t na = (size_t)(lmax + 1) * (lmax + 2);           // doubles per alm
    const size_t nseg = (size_t)(lmax + psb::SHT_CL_SEG) / psb::SHT_CL_SEG;
    if (int rc = scratch_reserve(dev, 0, 2 * na + (nseg + 1) * ((size_t)lmax + 1))) return rc;
    DeviceScratch& s = g_scratch[dev];
    double *a = s.X[0], *b = a + na, *c = b + na, *part = c + lmax + 1;
    CUDA_TRY(cudaMemcpyAsync(a, alm1, na * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    if (alm2 != alm1) CUDA_TRY(cudaMemcpyAsync(b, alm2, na * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    if (int rc = sht_alm2cl(lmax, (const double2*)a, (const double2*)(alm2 != alm1 ? b : a), part, c, s.stream)) return rc;
    CUDA_TRY(cudaMemcpyAsync(cl, c, ((size_t)lmax + 1) * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return OK;
}

// products dealt to device `dev`: k = first, first + stride, ...; every unique map the device needs is uploaded once
static int map2alm_many_on_device(int dev, int first, int stride, int nside, int lmax, int niter, int nmaps,
                                  const double* const* maps, int nprod, const int* idx, const double* scale,
                                  double* const* alm, std::string* err, int ngpus_in_call)
{
    auto run = [&]() -> int {
        CUDA_TRY(cudaSetDevice(dev));
        if (int rc = scratch_reserve(dev, 5, 16)) return rc;
        cudaStream_t st = g_scratch[dev].stream;
        ShtPlan* P = nullptr;
        if (int rc = sht_plan(dev, nside, lmax, st, &P)) return rc;
        const size_t nb = (size_t)P->D.npix * sizeof(double);
        std::vector<double*> dmap(nmaps, nullptr);
        int rc = OK;
        for (int k = first; k < nprod && rc == OK; k += stride)
            for (int f = 0; f < 3 && rc == OK; ++f) {
                const int i = idx[3 * k + f];
                if (i < 0 || dmap[i]) continue;
                const cudaError_t e = cudaMalloc(&dmap[i], nb);
                if (e != cudaSuccess) { dmap[i] = nullptr; rc = fail(ERR_OOM, "map2alm_many: device copy of map %d: %s", i, cudaGetErrorString(e)); break; }
                rc = upload_async(dev, dmap[i], maps[i], nb, st, ngpus_in_call);
            }
        // The alm of product n goes to the host on the copy stream while product n + 1 is computed (two alm buffers).  A
        // download into pageable memory blocks this thread until the bytes have arrived, so the kernels of product n + 1
        // are queued BEFORE product n is fetched; buffer b is free again once its download has returned.
        cudaStream_t cs = g_scratch[dev].copy_stream;
        cudaEvent_t* ev = g_scratch[dev].ev;                 // ev[b]: alm in buffer b computed
        const size_t alm_bytes = (size_t)2 * P->D.nalm * sizeof(double);
        auto fetch = [&](int k, int b) -> int {
            if (cudaStreamWaitEvent(cs, ev[b], 0) != cudaSuccess) return fail(ERR_CUDA, "map2alm_many: event wait");
            return download_blocking(dev, alm[k], b ? P->alm2 : P->alm, alm_bytes, cs, ngpus_in_call);
        };
        int n = 0, kprev = -1;
        for (int k = first; k < nprod && rc == OK; k += stride, ++n) {
            const int* ix = idx + 3 * k;
            const int b = n & 1;
            double* dalm = b ? P->alm2 : P->alm;
            psb::sht_product_kernel<<<1184, 256, 0, st>>>(P->D.npix, dmap[ix[0]], ix[1] >= 0 ? dmap[ix[1]] : nullptr,
                                                         ix[2] >= 0 ? dmap[ix[2]] : nullptr, scale[k], P->work);
            if (cudaGetLastError() != cudaSuccess) { rc = fail(ERR_CUDA, "map2alm_many: product kernel"); break; }
            rc = sht_map2alm(*P, st, P->work, dalm, niter);
            if (rc != OK) break;
            if (cudaEventRecord(ev[b], st) != cudaSuccess) { rc = fail(ERR_CUDA, "map2alm_many: event record"); break; }
            if (kprev >= 0) rc = fetch(kprev, b ^ 1);
            kprev = k;
        }
        if (rc == OK && kprev >= 0) rc = fetch(kprev, (n - 1) & 1);
        const cudaError_t es = cudaStreamSynchronize(st), ec = cudaStreamSynchronize(cs);
        if (rc == OK && (es != cudaSuccess || ec != cudaSuccess)) rc = fail(ERR_CUDA, "map2alm_many: %s", cudaGetErrorString(es != cudaSuccess ? es : ec));
        for (double* p : dmap) cudaFree(p);
        return rc;
    };
    const int rc = run();
    if (rc != OK && err) *err = g_err;
    return rc;
}

int psb200_map2alm_many(int nside, int lmax, int niter, int nmaps, const double* const* maps, int nprod, const int* idx,
                        const double* scale, double* const* alm, int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (nmaps < 1 || nprod < 1 || !maps || !idx || !scale || !alm || niter < 0) return fail(ERR_ARG, "map2alm_many: bad arguments");
    for (int i = 0; i < nmaps; ++i)
        if (!maps[i]) return fail(ERR_ARG, "map2alm_many: null map %d", i);
    for (int k = 0; k < nprod; ++k) {
        if (!alm[k]) return fail(ERR_ARG, "map2alm_many: null output %d", k);
        if (idx[3 * k] < 0 || idx[3 * k] >= nmaps) return fail(ERR_ARG, "map2alm_many: product %d has no first factor", k);
        for (int f = 1; f < 3; ++f)
            if (idx[3 * k + f] >= nmaps || idx[3 * k + f] < -1) return fail(ERR_ARG, "map2alm_many: product %d names map %d", k, idx[3 * k + f]);
        if (idx[3 * k + 1] < 0 && idx[3 * k + 2] >= 0) return fail(ERR_ARG, "map2alm_many: product %d skips its second factor", k);
    }
    if (int rc = sht_check(nside, lmax)) return rc;
    int ng = 0;
    if (int rc = resolve_ngpus(ngpus, &ng)) return rc;
    if (ng > nprod) ng = nprod;
    int cur = 0;
    CUDA_TRY(cudaGetDevice(&cur));
    if (ng == 1) {                                            // the caller's current device
        if (cur >= 16) return fail(ERR_ARG, "device index %d above the supported 15", cur);
        return map2alm_many_on_device(cur, 0, 1, nside, lmax, niter, nmaps, maps, nprod, idx, scale, alm, nullptr, 1);
    }
    std::vector<int> rcs(ng, OK);
    std::vector<std::string> errs(ng);
    std::vector<std::thread> th;
    for (int g = 1; g < ng; ++g)
        th.emplace_back([&, g] { rcs[g] = map2alm_many_on_device(g, g, ng, nside, lmax, niter, nmaps, maps, nprod, idx, scale, alm, &errs[g], ng); });
    rcs[0] = map2alm_many_on_device(0, 0, ng, nside, lmax, niter, nmaps, maps, nprod, idx, scale, alm, &errs[0], ng);
    for (auto& t : th) t.join();
    cudaSetDevice(cur);
    for (int g = 0; g < ng; ++g)
        if (rcs[g] != OK) { g_err = "device " + std::to_string(g) + ": " + errs[g]; return rcs[g]; }
    return OK;
}

/* Frees the tables and work buffers the transforms keep per device between calls (12 GB at nside 2048, lmax 6143). */
int psb200_sht_release(void)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    const int n = device_count();
    if (n <= 0) return OK;
    int cur = 0;
    CUDA_TRY(cudaGetDevice(&cur));
    for (int d = 0; d < n && d < 16; ++d) {
        if (!g_sht[d].nside && !g_cl[d].p) continue;
        CUDA_TRY(cudaSetDevice(d));
        CUDA_TRY(cudaDeviceSynchronize());
        sht_free(g_sht[d]);
        cudaFree(g_cl[d].p);
        g_cl[d] = ClScratch{};
    }
    CUDA_TRY(cudaSetDevice(cur));
    return OK;
}

/* Work accounting of one Legendre pass (analysis or synthesis) as the kernel tiles it; host arithmetic.
 *   out[0] executed (l, m, ring-pair slot) steps: every started warp x its 16-step passes x 32 R slots
 *   out[1] steps of rings the transform starts (m <= mlim), l = m..lmax
 *   out[2] warps started   out[3] R   out[4] chunks   out[5] l steps per pass */
int psb200_sht_stats(int nside, int lmax, long long* out)
{
    if (nside < 1 || nside > 2048 || (nside & (nside - 1)) || lmax < 0 || lmax > 4 * nside - 1 || !out)
        return fail(ERR_ARG, "sht_stats: bad arguments");
    const int R = sht_R(), C = sht_C(), nrp = 2 * nside, nchunks = (nrp + 32 * R - 1) / (32 * R);
    std::vector<double> ml(nrp);
    for (int p = 0; p < nrp; ++p) ml[p] = psb::sht_mlim(lmax, psb::sht_ring(nside, p).s);
    long long exec = 0, live = 0, warps = 0;
    for (int m = 0; m <= lmax; ++m) {
        const long long passes = (lmax - m + C) / C;
        for (int c = 0; c < nchunks; ++c) {
            const int plast = std::min((c + 1) * 32 * R, nrp) - 1;
            if ((double)m <= ml[plast]) { exec += passes * C * 32 * R; ++warps; }
        }
        for (int p = 0; p < nrp; ++p)
            if ((double)m <= ml[p]) live += lmax - m + 1;
    }
    out[0] = exec; out[1] = live; out[2] = warps; out[3] = R; out[4] = nchunks; out[5] = C;
    return OK;
}
