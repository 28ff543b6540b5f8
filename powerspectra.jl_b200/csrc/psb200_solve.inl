// psb200_solve.inl -- decoupling on the device.  Included by psb200.cu inside its anonymous namespace.
//
// What it replaces on the host side of the reference (which stays available):
//   M \ pCl                         /root/reference/src/blockspectralmatrix.jl:124-129 (lu + solve of parent(M))
//   M_EE_BB \ [pCl_EE; pCl_BB]      src/blockspectralmatrix.jl:89-122 (lu of the dense 2N x 2N hvcat), with the block
//                                   matrices of src/modecoupling.jl:213-223: [M++ M--; M-- M++] and [M++ -M--; -M-- M++];
//                                   here split exactly into the two N x N systems of M++ + M-- and M++ - M--
//   maskedalm2spectra's solves      src/modecoupling.jl:348-377
//   decouple_covmat(Y, B1, B2)      src/covariance.jl:8-14:  B1^-1 Y (B2^-1)^T through lu(B1'), lu(B2')
//
// The mode-coupling matrix never leaves the GPUs: every device computes its l1 row band and, with peer access, its pair
// kernel STORES the band straight into the root device's matrix over NVLink (no gather step, no staging copy; without
// peer access the band is computed locally and moved with one cudaMemcpyPeerAsync).  The root fills both triangles
// (finish_kernel), factorises with cuSOLVER getrf (partial pivoting, what Julia's `lu` calls in LAPACK) and solves; only
// the spectra -- O(N) doubles -- cross PCIe.  cuSOLVER is loaded with dlopen on first use, so libpsb200.so itself
// depends on nothing but the CUDA runtime; the LU is not the north-star hot path (SURVEY.md 8f-2 names getrf/getrs).

// (psb200.cu includes <cusolverDn.h> and <dlfcn.h> at file scope: declarations inside an unnamed namespace get internal linkage)

struct CuSolver {
    void* so = nullptr;
    decltype(&cusolverDnCreate) create = nullptr;
    decltype(&cusolverDnSetStream) set_stream = nullptr;
    decltype(&cusolverDnDgetrf_bufferSize) getrf_buf = nullptr;
    decltype(&cusolverDnDgetrf) getrf = nullptr;
    decltype(&cusolverDnDgetrs) getrs = nullptr;
    cusolverDnHandle_t handle[16] = {};
};
CuSolver g_cs;

int cusolver_load()
{
    if (g_cs.so) return OK;
    const char* names[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so",
                           "/usr/local/cuda/lib64/libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so"};
    void* so = nullptr;
    for (const char* n : names) { so = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (so) break; }
    if (!so) return fail(ERR_CUDA, "cuSOLVER not found (dlopen libcusolver.so.11): %s", dlerror());
#define PSB_SYM(field, name)                                                                   \
    g_cs.field = reinterpret_cast<decltype(g_cs.field)>(dlsym(so, name));                       \
    if (!g_cs.field) { dlclose(so); return fail(ERR_CUDA, "cuSOLVER symbol %s missing", name); }
    PSB_SYM(create, "cusolverDnCreate")
    PSB_SYM(set_stream, "cusolverDnSetStream")
    PSB_SYM(getrf_buf, "cusolverDnDgetrf_bufferSize")
    PSB_SYM(getrf, "cusolverDnDgetrf")
    PSB_SYM(getrs, "cusolverDnDgetrs")
#undef PSB_SYM
    g_cs.so = so;
    return OK;
}

#define CS_TRY(expr)                                                                                        \
    do {                                                                                                    \
        cusolverStatus_t s_ = (expr);                                                                       \
        if (s_ != CUSOLVER_STATUS_SUCCESS)                                                                  \
            return fail(s_ == CUSOLVER_STATUS_ALLOC_FAILED ? ERR_OOM : ERR_CUDA, "%s: cuSOLVER status %d (%s:%d)", \
                        #expr, (int)s_, __FILE__, __LINE__);                                                \
    } while (0)

// per-device scratch of the solves (grown on demand, kept between calls)
struct SolveScratch {
    double* A = nullptr;  size_t capA = 0;      // the matrix being factorised (N x N or 2N x 2N)
    double* B = nullptr;  size_t capB = 0;      // right-hand sides / transposed operand
    double* work = nullptr; size_t capW = 0;    // getrf workspace
    int* ipiv = nullptr;  size_t capP = 0;
    int* info = nullptr;
};
SolveScratch g_solve[16];

template <class T>
int grow(T** p, size_t* cap, size_t n)
{
    if (*cap >= n) return OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CUDA_TRY(cudaMalloc(p, n * sizeof(T)));
    *cap = n;
    return OK;
}

// out[j*ldo + i] = in[i*ldi + j]  (n x n, 32x32 tiles through shared memory)
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, long ldi, double* __restrict__ out,
                                                        long ldo, int n)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
    for (int r = ty; r < 32; r += 8) {
        const int i = bi + r, j = bj + tx;
        tile[r][tx] = (i < n && j < n) ? in[(long)i * ldi + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bj + r, i = bi + tx;
        if (i < n && j < n) out[(long)j * ldo + i] = tile[tx][r];
    }
}

// LU of the n x n matrix dA (in place) on the current device + solve of nrhs right-hand sides dB (in place).
// trans = 0: A X = B;  1: A^T X = B.  Returns ERR_SINGULAR with the LAPACK info when a pivot is exactly zero
// (Julia's lu throws SingularException there).
int lu_factor(int dev, cudaStream_t st, double* dA, int n, long lda)
{
    if (int rc = cusolver_load()) return rc;
    if (!g_cs.handle[dev]) CS_TRY(g_cs.create(&g_cs.handle[dev]));
    cusolverDnHandle_t h = g_cs.handle[dev];
    CS_TRY(g_cs.set_stream(h, st));
    SolveScratch& s = g_solve[dev];
    int lwork = 0;
    CS_TRY(g_cs.getrf_buf(h, n, n, dA, (int)lda, &lwork));
    if (int rc = grow(&s.work, &s.capW, (size_t)std::max(lwork, 1))) return rc;
    if (int rc = grow(&s.ipiv, &s.capP, (size_t)n)) return rc;
    if (!s.info) CUDA_TRY(cudaMalloc(&s.info, 2 * sizeof(int)));
    CS_TRY(g_cs.getrf(h, n, n, dA, (int)lda, s.work, s.ipiv, s.info));
    int info = 0;
    CUDA_TRY(cudaMemcpyAsync(&info, s.info, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (info > 0) return fail(ERR_SINGULAR, "matrix is singular: U(%d,%d) is exactly zero", info, info);
    if (info < 0) return fail(ERR_CUDA, "getrf: illegal argument %d", -info);
    return OK;
}
int lu_solve(int dev, cudaStream_t st, const double* dA, int n, long lda, double* dB, long ldb, int nrhs, int trans)
{
    cusolverDnHandle_t h = g_cs.handle[dev];
    CS_TRY(g_cs.set_stream(h, st));
    SolveScratch& s = g_solve[dev];
    CS_TRY(g_cs.getrs(h, trans ? CUBLAS_OP_T : CUBLAS_OP_N, n, nrhs, dA, (int)lda, s.ipiv, dB, (int)ldb, s.info));
    return OK;
}

// ---------------------------------------------------------------------------------------
// All outputs of a host job as FULL matrices (both triangles, (2l+1) factors applied) in the ROOT device's X scratch.
// ngpus == 1: root = the caller's current device.  ngpus > 1: root = device 0, bands of devices g > 0 arrive through
// peer stores of their pair kernels (or one peer copy per output when peer access is unavailable).
// ---------------------------------------------------------------------------------------
bool g_peer_on[16][16] = {};

int enable_peer(int from, int to, bool* direct)
{
    *direct = false;
    if (from == to) { *direct = true; return OK; }
    if (getenv("PSB200_NO_PEER_STORES")) return OK;            // test hook: force the explicit peer-copy path
    if (g_peer_on[from][to]) { *direct = true; return OK; }
    int can = 0;
    CUDA_TRY(cudaDeviceCanAccessPeer(&can, from, to));
    if (!can) return OK;
    cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);          // current device == from
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    if (e != cudaSuccess) { cudaGetLastError(); return OK; }     // fall back to explicit peer copies
    g_peer_on[from][to] = true;
    *direct = true;
    return OK;
}

int full_on_root(const HostJob& hj, int ngpus, int* root_out)
{
    int cur = 0;
    cudaGetDevice(&cur);
    const int root = ngpus == 1 ? cur : 0;
    *root_out = root;
    const int N = hj.lmax - hj.lmin + 1;
    CUDA_TRY(cudaSetDevice(root));
    for (int o = 0; o < hj.nout; ++o)
        if (int rc = scratch_reserve(root, o, (size_t)N * N)) return rc;
    DeviceScratch& R = g_scratch[root];
    std::vector<int> edges(ngpus + 1);
    if (ngpus == 1) { edges[0] = hj.lmin; edges[1] = hj.lmax + 1; }
    else psb200_band_edges(hj.lmin, hj.lmax, hj.lenW, ngpus, edges.data());
    std::vector<int> rcs(ngpus, OK);
    std::vector<std::string> errs(ngpus);
    auto band = [&](int slot, int g) -> int {
        const int a = edges[slot], b = edges[slot + 1];
        if (b <= a) return OK;
        CUDA_TRY(cudaSetDevice(g));
        bool direct = false;
        if (int rc = enable_peer(g, root, &direct)) return rc;
        double* Xs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (direct) {
            for (int o = 0; o < hj.nout; ++o) Xs[o] = R.X[o];                 // kernels index rows from lmin
        } else {
            for (int o = 0; o < hj.nout; ++o) {
                if (int rc = scratch_reserve(g, o, (size_t)(b - a) * N)) return rc;
                Xs[o] = g_scratch[g].X[o] - (long)(a - hj.lmin) * N;
            }
        }
        if (int rc = run_on_device(hj, g, a, b, Xs, N)) return rc;             // also creates g's streams
        DeviceScratch& s = g_scratch[g];
        if (!direct)
            for (int o = 0; o < hj.nout; ++o) {
                const cudaError_t ce = cudaMemcpyPeerAsync(R.X[o] + (size_t)(a - hj.lmin) * N, root, s.X[o], g,
                                                           (size_t)(b - a) * N * sizeof(double), s.stream);
                if (ce != cudaSuccess)
                    return fail(ERR_COLL, "band of device %d could not be moved to device %d: %s", g, root, cudaGetErrorString(ce));
            }
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        return OK;
    };
    if (ngpus == 1) {
        if (int rc = band(0, root)) return rc;
    } else {
        // the root needs its streams before any peer writes into its scratch
        if (int rc = scratch_reserve(root, 5, 4)) return rc;
        std::vector<std::thread> th;
        for (int g = 1; g < ngpus; ++g)
            th.emplace_back([&, g] { rcs[g] = band(g, g); if (rcs[g] != OK) errs[g] = g_err; });
        rcs[0] = band(0, 0);
        if (rcs[0] != OK) errs[0] = g_err;
        for (auto& t : th) t.join();
        for (int g = 0; g < ngpus; ++g)
            if (rcs[g] != OK) { g_err = "device " + std::to_string(g) + ": " + errs[g]; cudaSetDevice(cur); return rcs[g]; }
    }
    CUDA_TRY(cudaSetDevice(root));
    const int nt = (N + 31) / 32;
    for (int o = 0; o < hj.nout; ++o) {
        finish_kernel<<<dim3(nt, nt), 256, 0, R.stream>>>(R.X[o], N, hj.lmin, N, hj.scale, 0);
        CUDA_TRY(cudaGetLastError());
    }
    return OK;
}

// One right-hand side / solution column of a system of nb stacked N-blocks: host pointers of its nb pieces.
struct RhsCol { const double* in[2]; double* out[2]; };

// Solve `A \ rhs` on the root for a matrix already in device memory (destroyed: LU in place), n = nb N rows.
int solve_on_root(int root, double* dA, int N, int nb, const std::vector<RhsCol>& cols)
{
    DeviceScratch& R = g_scratch[root];
    SolveScratch& s = g_solve[root];
    const int n = nb * N, nrhs = (int)cols.size();
    if (nrhs == 0) return OK;
    if (int rc = grow(&s.B, &s.capB, (size_t)n * nrhs)) return rc;
    for (int k = 0; k < nrhs; ++k)
        for (int b = 0; b < nb; ++b)
            CUDA_TRY(cudaMemcpyAsync(s.B + (size_t)k * n + (size_t)b * N, cols[k].in[b], (size_t)N * sizeof(double),
                                     cudaMemcpyHostToDevice, R.stream));
    if (int rc = lu_factor(root, R.stream, dA, n, n)) return rc;
    if (int rc = lu_solve(root, R.stream, dA, n, n, s.B, n, nrhs, 0)) return rc;
    for (int k = 0; k < nrhs; ++k)
        for (int b = 0; b < nb; ++b)
            CUDA_TRY(cudaMemcpyAsync(cols[k].out[b], s.B + (size_t)k * n + (size_t)b * N, (size_t)N * sizeof(double),
                                     cudaMemcpyDeviceToHost, R.stream));
    CUDA_TRY(cudaStreamSynchronize(R.stream));
    return OK;
}

// S = P + Q, D = P - Q  (N x N, ld = N)
__global__ void __launch_bounds__(256) sumdiff_kernel(const double* __restrict__ P, const double* __restrict__ Q,
                                                      double* __restrict__ S, double* __restrict__ D, size_t n2)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const double p = P[i], q = Q[i];
    S[i] = p + q;
    D[i] = p - q;
}

// The 2N x 2N block systems of src/modecoupling.jl:213-223,
//     [P Q; Q P] [x; y] = [p; q]   (M_EE_BB)        [P -Q; -Q P] [x; y] = [p; q]   (M_EB_BE),
// are block-circulant: with S = P + Q, D = P - Q they split exactly into two N x N systems,
//     EE_BB:  S (x+y) = p+q,  D (x-y) = p-q          EB_BE:  D (x+y) = p+q,  S (x-y) = p-q,
// so ONE LU of S and ONE of D serve both systems: 2 x (2/3) N^3 flops instead of the 2 x (2/3)(2N)^3 of factorising the
// dense hvcat twice (what src/blockspectralmatrix.jl:89-122 does on the host) -- 8x less work, same solution up to
// rounding (the split is an orthogonal similarity; checked against the dense host LU in tests/test_gpu_solve.py).
// plus / minus: right-hand-side columns of the EE_BB / EB_BE system (two N-pieces each).
int solve_block_systems_on_root(int root, const double* dP, const double* dQ, int N, const std::vector<RhsCol>& plus,
                                const std::vector<RhsCol>& minus)
{
    DeviceScratch& R = g_scratch[root];
    SolveScratch& s = g_solve[root];
    const size_t n2 = (size_t)N * N;
    if (int rc = grow(&s.A, &s.capA, 2 * n2)) return rc;
    double *dS = s.A, *dD = s.A + n2;
    sumdiff_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, R.stream>>>(dP, dQ, dS, dD, n2);
    CUDA_TRY(cudaGetLastError());
    const size_t np = plus.size(), nm = minus.size(), nc = np + nm;
    if (nc == 0) return OK;
    // host staging: column k of `hs` goes to S, of `hd` to D
    std::vector<double> hs(nc * N), hd(nc * N), xs(nc * N), xd(nc * N);
    for (size_t k = 0; k < nc; ++k) {
        const RhsCol& c = k < np ? plus[k] : minus[k - np];
        double* sum = (k < np ? hs.data() : hd.data()) + k * N;      // p + q
        double* dif = (k < np ? hd.data() : hs.data()) + k * N;      // p - q
        for (int i = 0; i < N; ++i) { sum[i] = c.in[0][i] + c.in[1][i]; dif[i] = c.in[0][i] - c.in[1][i]; }
    }
    std::vector<RhsCol> cs(nc), cd(nc);
    for (size_t k = 0; k < nc; ++k) {
        cs[k].in[0] = hs.data() + k * N; cs[k].out[0] = xs.data() + k * N;
        cd[k].in[0] = hd.data() + k * N; cd[k].out[0] = xd.data() + k * N;
    }
    if (int rc = solve_on_root(root, dS, N, 1, cs)) return rc;
    if (int rc = solve_on_root(root, dD, N, 1, cd)) return rc;
    for (size_t k = 0; k < nc; ++k) {
        const RhsCol& c = k < np ? plus[k] : minus[k - np];
        const double* a = (k < np ? xs.data() : xd.data()) + k * N;  // x + y
        const double* b = (k < np ? xd.data() : xs.data()) + k * N;  // x - y
        for (int i = 0; i < N; ++i) { c.out[0][i] = 0.5 * (a[i] + b[i]); c.out[1][i] = 0.5 * (a[i] - b[i]); }
    }
    return OK;
}

// B1^-1 Y (B2^-1)^T on the current device, Y in place; all N x N column-major device matrices.
int decouple_on_device(int dev, cudaStream_t st, int N, double* dY, long ldy, const double* dB1, long ld1,
                       const double* dB2, long ld2)
{
    SolveScratch& s = g_solve[dev];
    if (int rc = grow(&s.A, &s.capA, (size_t)N * N)) return rc;
    if (int rc = grow(&s.B, &s.capB, (size_t)N * N)) return rc;
    const dim3 grid((N + 31) / 32, (N + 31) / 32);
    // rdiv!(C', lu(B1')):  C <- B1^-1 C, through the LU of B1^T and a transposed solve
    transpose_kernel<<<grid, 256, 0, st>>>(dB1, ld1, s.A, N, N);
    CUDA_TRY(cudaGetLastError());
    if (int rc = lu_factor(dev, st, s.A, N, N)) return rc;
    if (int rc = lu_solve(dev, st, s.A, N, N, dY, ldy, N, 1)) return rc;
    // rdiv!(C, lu(B2')):  C <- C (B2^T)^-1  <=>  (B2^T)^T C^T = ... solved on the transposed operand
    transpose_kernel<<<grid, 256, 0, st>>>(dB2, ld2, s.A, N, N);
    CUDA_TRY(cudaGetLastError());
    if (int rc = lu_factor(dev, st, s.A, N, N)) return rc;
    transpose_kernel<<<grid, 256, 0, st>>>(dY, ldy, s.B, N, N);
    CUDA_TRY(cudaGetLastError());
    if (int rc = lu_solve(dev, st, s.A, N, N, s.B, N, N, 1)) return rc;
    transpose_kernel<<<grid, 256, 0, st>>>(s.B, N, dY, ldy, N);
    CUDA_TRY(cudaGetLastError());
    return OK;
}
