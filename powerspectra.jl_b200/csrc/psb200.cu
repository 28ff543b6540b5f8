// psb200.cu -- C ABI (include/psb200.h) + host plumbing of libpsb200.so.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU path in this file: every compute entry point launches CUDA kernels and
// fails with PSB200_ERR_NODEVICE when no device is present.
#include "../../include/psb200.h"
#include "psb200_common.cuh"
#include "psb200_pair_v1.cuh"
#include "psb200_pair_v2.cuh"
#include "psb200_pair_v3.cuh"
#include "psb200_pair_v4.cuh"
#include "psb200_lowrows.cuh"
#include "psb200_quickpol.cuh"
#include "psb200_zonal.cuh"
#include "psb200_sht.cuh"

#include <cusolverDn.h>      // types + prototypes only: the library is loaded with dlopen (psb200_solve.inl)
#include <dlfcn.h>
#include <emmintrin.h>       // SSE2 streaming stores of the staged delivery (host side)

#include <algorithm>
#include <condition_variable>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <tuple>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

enum { OK = 0, ERR_ARG = 1, ERR_CUDA = 2, ERR_COLL = 3, ERR_OOM = 4, ERR_NODEVICE = 5, ERR_SINGULAR = 6 };

thread_local std::string g_err;
std::mutex g_mutex;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(e_ == cudaErrorMemoryAllocation ? ERR_OOM : ERR_CUDA, "%s: %s (%s:%d)", \
                        #expr, cudaGetErrorString(e_), __FILE__, __LINE__);                  \
    } while (0)

int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---------------------------------------------------------------------------------------
// finish kernel: X holds, for l1 <= l2, x = X[(l1-lmin)*ld + (l2-lmin)] (row l1 of the upper
// triangle contiguous = column l1 of the column-major result).  Writes, in place,
//   A[l2,l1] (same address)            = scale ? (2 l1+1) x : x
//   A[l1,l2] (address l1 + l2*ld)      = scale ? (2 l2+1) x : x
// 32x32 tiles through shared memory so both sides are coalesced.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finish_kernel(double* __restrict__ X, long ld, int lmin, int N, int scale,
                                                     int tile_lo)
{
    __shared__ double tile[32][33];
    const int bi = blockIdx.y + tile_lo, bj = blockIdx.x;      // tile row (l1) and tile column (l2)
    if (bi > bj) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;          // i = l1-lmin, j = l2-lmin
        double v = 0.0;
        if (i < N && j < N && i <= j) v = X[(long)i * ld + j];
        tile[r][tx] = v;
    }
    __syncthreads();
    // lower-triangle side: same addresses, factor (2 l1 + 1)
    if (scale) {
        for (int r = ty; r < 32; r += 8) {
            const int i = bi * 32 + r, j = bj * 32 + tx;
            if (i < N && j < N && i <= j) X[(long)i * ld + j] = (double)(2 * (i + lmin) + 1) * tile[r][tx];
        }
    }
    // upper-triangle side: transposed addresses, factor (2 l2 + 1); skip the diagonal (done above)
    for (int r = ty; r < 32; r += 8) {
        const int j = bj * 32 + r, i = bi * 32 + tx;          // write A[i, j] at i + j*ld, contiguous in i
        if (i < N && j < N && i < j) {
            const double v = tile[tx][r];
            X[(long)j * ld + i] = scale ? (double)(2 * (j + lmin) + 1) * v : v;
        }
    }
}

// Band variant of stage 2 for a device that owns rows [a, a+nb) only (slab X, ld doubles per row,
// X[(l1-a)*ld + (l2-lmin)]).  For the columns to the right of the band's diagonal block
// (c0 <= l2-lmin < N) it writes T[(l2-lmin-c0)*nb + (l1-a)] = s2 * x  (the block ROW of the result,
// column-major, ready for a pitched D2H) and, in place, X = s1 * x (the block COLUMN part);
// s1 = 2 l1+1, s2 = 2 l2+1 for MCM jobs (scale = 1), both 1 for covariance blocks.
__global__ void __launch_bounds__(256) band_transpose_kernel(double* __restrict__ X, long ld, double* __restrict__ T,
                                                             int nb, int a, int lmin, int c0, int N, int scale)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 32, cc0 = c0 + blockIdx.x * 32;     // band-relative row, matrix-relative column
    for (int r = ty; r < 32; r += 8) {
        const int i = r0 + r, j = cc0 + tx;
        double v = 0.0;
        if (i < nb && j < N) {
            v = X[(long)i * ld + j];
            if (scale) X[(long)i * ld + j] = (double)(2 * (a + i) + 1) * v;
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = cc0 + r, i = r0 + tx;
        if (i < nb && j < N) {
            const double v = tile[tx][r];
            T[(long)(j - c0) * nb + i] = scale ? (double)(2 * (lmin + j) + 1) * v : v;
        }
    }
}

// FP64 pipe microbenchmark: 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------------------
// launch of one job on the current device
// ---------------------------------------------------------------------------------------
int kernel_version()
{
    // read on every call so a test can switch kernels inside one process
    // v1: simple kernel (sqrt/divide recurrence, sum normalisation); v2: tuned recurrence kernel; v3: closed-form
    // kernel with per-chunk table staging; default v4: closed-form kernel with asynchronous ring staging.  v1 and v2
    // are on-device cross-checks of the closed forms by independent methods, v3 the A/B partner of v4; none is a fallback.
    const char* e = getenv("PSB200_KERNEL");
    if (e && strcmp(e, "v1") == 0) return 1;
    if (e && strcmp(e, "v2") == 0) return 2;
    if (e && strcmp(e, "v3") == 0) return 3;
    return 4;
}

// PSB200_TRACE=1: wall-clock of the host-level calls per band on stderr (syncs only where the call syncs anyway);
// PSB200_TRACE=2: also every phase of every launch (adds a stream sync per phase: serialises the sub-bands, so the
// totals of that mode say nothing about overlap)
int trace_level()
{
    const char* e = getenv("PSB200_TRACE");
    return (e && *e) ? atoi(e) : 0;
}
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(int level = 1) : on(trace_level() >= level), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what, int dev, cudaStream_t st)
    {
        if (!on) return;
        cudaSetDevice(dev);
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[psb200] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---------------------------------------------------------------------------------------
// v2 support: per-device constant tables, cached block lists, per-call W' buffer
// ---------------------------------------------------------------------------------------
struct DevTables {
    int lmax = -1;
    int nS = 0;
    double *S = nullptr, *IS = nullptr, *INV = nullptr, *gam = nullptr, *igam = nullptr;
    double* seq[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};    // v4: G0 G1 G2 H0 H1, each nS + V4_PAD doubles
};
DevTables g_tables[16];
std::mutex g_tab_mutex;

int ensure_tables(int dev, int lmax, cudaStream_t st, DevTables** out)
{
    std::lock_guard<std::mutex> lk(g_tab_mutex);
    DevTables& t = g_tables[dev];
    if (t.lmax >= lmax) { *out = &t; return OK; }
    const int want = std::max(lmax, 1024);
    const int nS = 4 * want + 4096;
    const int ng = nS;                          // the closed-form kernel reads g(n) up to n ~ 2 lmax + a chunk
    std::vector<double> S(nS), IS(nS), INV(nS), G(ng), IG(ng);
    for (int n = 0; n < nS; ++n) {
        const long double r = sqrtl((long double)n);
        S[n] = (double)r;
        IS[n] = n ? (double)(1.0L / r) : 0.0;
        INV[n] = n ? (double)(1.0L / (long double)n) : 0.0;
    }
    long double g = 1.0L;                       // binom(2n,n)/4^n = prod (2i-1)/(2i)
    G[0] = 1.0; IG[0] = 1.0;
    for (int n = 1; n < ng; ++n) {
        g *= (long double)(2 * n - 1) / (long double)(2 * n);
        G[n] = (double)g;
        IG[n] = (double)(1.0L / g);
    }
    // v4 sequences (psb200_pair_v4.cuh), V4_PAD zeros in front: g, (2n+1) g, 2n g, 1/((2n+1) g), (2n+2)/((2n+1) g)
    std::vector<double> Q[5];
    for (auto& q : Q) q.assign((size_t)nS + psb::V4_PAD, 0.0);
    g = 1.0L;
    for (int n = 0; n < nS; ++n) {
        if (n) g *= (long double)(2 * n - 1) / (long double)(2 * n);
        const size_t i = (size_t)n + psb::V4_PAD;
        const long double h0 = 1.0L / ((long double)(2 * n + 1) * g);
        Q[0][i] = (double)g; Q[1][i] = (double)((long double)(2 * n + 1) * g); Q[2][i] = (double)((long double)(2 * n) * g);
        Q[3][i] = (double)h0; Q[4][i] = (double)((long double)(2 * n + 2) * h0);
    }
    if (t.S) {                                   // growing: nothing may still be reading the old tables
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFree(t.S); cudaFree(t.IS); cudaFree(t.INV); cudaFree(t.gam); cudaFree(t.igam);
        for (double* q : t.seq) cudaFree(q);
        t = DevTables{};
    }
    CUDA_TRY(cudaMalloc(&t.S, nS * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.IS, nS * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.INV, nS * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.gam, ng * sizeof(double)));
    CUDA_TRY(cudaMalloc(&t.igam, ng * sizeof(double)));
    // Uploads go on the LAUNCHING stream (the kernels that read the tables run on non-blocking streams, which
    // do not order against the legacy default stream a plain cudaMemcpy uses), and the stream is drained
    // before the pageable host vectors die.
    CUDA_TRY(cudaMemcpyAsync(t.S, S.data(), nS * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t.IS, IS.data(), nS * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t.INV, INV.data(), nS * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t.gam, G.data(), ng * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t.igam, IG.data(), ng * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 5; ++k) {
        CUDA_TRY(cudaMalloc(&t.seq[k], Q[k].size() * sizeof(double)));
        CUDA_TRY(cudaMemcpyAsync(t.seq[k], Q[k].data(), Q[k].size() * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    t.lmax = want; t.nS = nS;
    *out = &t;
    return OK;
}

// Block list of one launch: (first l1, d_lo) tiles of the band's upper triangle, heaviest first
// (longest-processing-time order keeps the 148 SMs balanced to the last wave).  A tile is nr consecutive rows
// x (32/nr) r pairs of each: consecutive d (ds = 1) or d of one parity (ds = 2, (0,0,0)-recurrence jobs).
struct BlockList { int4* d = nullptr; int n = 0; unsigned long stamp = 0; };
typedef std::tuple<int, int, int, std::vector<int>, int, int, int> BlockKey;     // dev, lmax, lenW, bands (lo, hi, lo, hi ...), d stride, pairs per thread, rows per warp
std::map<BlockKey, BlockList> g_blocks;
unsigned long g_block_stamp = 0;

// Tiling of rows [row_lo, row_hi) as the tuned kernel runs it.  fn(l1_first, d_lo, steps) per warp.
template <class F>
void for_each_tile(int lmax, int lenW, int row_lo, int row_hi, int ds, int r, int nr, F fn)
{
    const int span = (32 / nr) * r;
    for (int l1 = row_lo; l1 < row_hi; l1 += nr) {
        const int l1_last = std::min(l1 + nr, row_hi) - 1;
        const int nd = lmax - l1 + 1;
        for (int base = 0; base < nd; base += ds * span) {
            for (int par = 0; par < ds; ++par) {
                const int d_lo = base + par;
                if (d_lo >= nd) continue;
                const long last = (long)lenW - 1 - d_lo;
                const long steps = last < 0 ? 0 : std::min<long>(span - 1 + (2 * l1_last) / ds, last / ds) + 1;
                fn(l1, d_lo, steps);
            }
        }
    }
}

int ensure_blocks(int dev, const psb::PairArgs& A, int ds, int r, int nr, cudaStream_t st, BlockList* out)
{
    std::lock_guard<std::mutex> lk(g_tab_mutex);
    std::vector<int> bands{A.row_lo, A.row_hi};
    for (int k = 0; k < A.nxb; ++k) { bands.push_back(A.xb[2 * k]); bands.push_back(A.xb[2 * k + 1]); }
    const BlockKey key(dev, A.lmax, A.lenW, bands, ds, r, nr);
    auto it = g_blocks.find(key);
    if (it != g_blocks.end()) { it->second.stamp = ++g_block_stamp; *out = it->second; return OK; }
    // tiles of all bands of the launch in ONE list, heaviest first: a rank that owns a low and a high band (folded
    // split) fills the tail of its long tiles with its short ones
    std::vector<std::pair<long, int4>> v;
    for (size_t b = 0; b + 1 < bands.size(); b += 2)
        for_each_tile(A.lmax, A.lenW, bands[b], bands[b + 1], ds, r, nr,
                      [&](int l1, int d_lo, long steps) { v.push_back({steps, make_int4(l1, d_lo, bands[b + 1], 0)}); });
    std::stable_sort(v.begin(), v.end(), [](const std::pair<long, int4>& a, const std::pair<long, int4>& b) { return a.first > b.first; });
    std::vector<int4> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = v[i].second;
    // LRU per device (a list is only ever evicted by a call on the device that owns it, after that device has
    // drained: lists of other devices may be in use by their own worker threads)
    size_t mine = 0;
    for (auto& kv : g_blocks) if (std::get<0>(kv.first) == dev) ++mine;
    if (mine >= 1024) {            // 16 sub-bands x a dozen jobs x a few sizes; a list is a few hundred KB at most
        auto old = g_blocks.end();
        for (auto jt = g_blocks.begin(); jt != g_blocks.end(); ++jt)
            if (std::get<0>(jt->first) == dev && (old == g_blocks.end() || jt->second.stamp < old->second.stamp)) old = jt;
        CUDA_TRY(cudaDeviceSynchronize());       // current device == dev (launch_job asked cudaGetDevice)
        cudaFree(old->second.d);
        g_blocks.erase(old);
    }
    BlockList bl;
    bl.n = (int)h.size();
    bl.stamp = ++g_block_stamp;
    if (bl.n) {
        CUDA_TRY(cudaMalloc(&bl.d, h.size() * sizeof(int4)));
        // on the launching stream, drained before the pageable host vector dies (see ensure_tables)
        CUDA_TRY(cudaMemcpyAsync(bl.d, h.data(), h.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    g_blocks[key] = bl;
    *out = bl;
    return OK;
}

// W' scratch: one persistent buffer per (device, stream); reuse on one stream is stream-ordered
// and therefore safe without synchronisation (and avoids a cudaMallocAsync per call).
struct WpBuf { double* p = nullptr; size_t cap = 0; unsigned long stamp = 0; };
std::map<std::pair<int, cudaStream_t>, WpBuf> g_wp;

int wp_reserve(int dev, cudaStream_t st, size_t n, double** out)
{
    std::lock_guard<std::mutex> lk(g_tab_mutex);
    WpBuf& b = g_wp[{dev, st}];
    b.stamp = ++g_block_stamp;
    if (b.cap < n) {
        if (b.p) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
        if (g_wp.size() > 32) {                  // forget the stalest stream's buffer
            auto old = g_wp.end();
            for (auto it = g_wp.begin(); it != g_wp.end(); ++it)
                if (&it->second != &b && (old == g_wp.end() || it->second.stamp < old->second.stamp)) old = it;
            if (old != g_wp.end()) { CUDA_TRY(cudaDeviceSynchronize()); cudaFree(old->second.p); g_wp.erase(old); }
        }
        const size_t want = n + n / 4;
        CUDA_TRY(cudaMalloc(&b.p, want * sizeof(double)));
        b.cap = want;
    }
    *out = b.p;
    return OK;
}

template <int JOB>
int launch_job(const psb::PairArgs& A_in, cudaStream_t st)
{
    // the bands of the launch: [row_lo, row_hi) and the extra ones, empty ones dropped
    std::vector<int> bands;
    auto add_band = [&](int lo, int hi) {
        if constexpr (psb::job_has_spin2(JOB)) {
            // rows l1 < 2 of the spin-2 jobs (true symbol 0; what the reference's family routine yields): psb200_lowrows.cuh
            if (lo < 2 && hi > lo) {
                psb::PairArgs L = A_in;
                L.row_lo = lo; L.row_hi = hi; L.nxb = 0;
                const int nlow = std::min(hi, 2) - lo;
                dim3 grid((L.lmax - lo + 1 + 127) / 128, nlow);
                psb::low_rows_kernel<JOB><<<grid, 128, 0, st>>>(L);
                lo = std::min(hi, 2);
            }
        }
        if (hi > lo) { bands.push_back(lo); bands.push_back(hi); }
    };
    add_band(A_in.row_lo, A_in.row_hi);
    for (int k = 0; k < A_in.nxb; ++k) add_band(A_in.xb[2 * k], A_in.xb[2 * k + 1]);
    CUDA_TRY(cudaGetLastError());
    if (bands.empty()) return OK;
    psb::PairArgs A = A_in;
    A.row_lo = bands[0]; A.row_hi = bands[1];
    A.nxb = (int)bands.size() / 2 - 1;
    for (size_t k = 2; k < bands.size(); ++k) A.xb[k - 2] = bands[k];
    if (kernel_version() == 1) {
        for (size_t b = 0; b + 1 < bands.size(); b += 2) {          // the simple kernel runs one grid per band
            psb::PairArgs B = A;
            B.row_lo = bands[b]; B.row_hi = bands[b + 1]; B.nxb = 0;
            const int maxcols = B.lmax - B.row_lo + 1;
            dim3 grid((maxcols + psb::V1_THREADS - 1) / psb::V1_THREADS, B.row_hi - B.row_lo);
            psb::pair_kernel_v1<JOB><<<grid, psb::V1_THREADS, 0, st>>>(B);
            CUDA_TRY(cudaGetLastError());
        }
        return OK;
    }
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", dev);
    Trace tr(2);
    tr.mark("  (inputs uploaded)", dev, st);
    DevTables* t = nullptr;
    if (int rc = ensure_tables(dev, A.lmax, st, &t)) return rc;
    const int kv = kernel_version();
    const bool v3 = kv >= 3;                    // v3 and v4 share tiling, W' layout and work accounting
    BlockList bl;
    if (v3) {
        if (int rc = ensure_blocks(dev, A, 2, psb::v3_r(JOB), psb::v3_nr(JOB), st, &bl)) return rc;
    } else {
        if (int rc = ensure_blocks(dev, A, psb::v2_family(JOB) == psb::FAM_00 ? 2 : 1, psb::v2_r(JOB), psb::v2_nr(JOB), st, &bl)) return rc;
    }
    // W'[j][q] = (2j+1) W_q[j] / 4pi, zero-padded so staging never reads past the end
    const int nqp = v3 ? psb::v3_nqp(JOB) : psb::v2_nqp(JOB);
    const int rows_w = A.lenW + 2 * (psb::V2_TC_MAX + psb::V2_PB_MAX) + 2;
    double* Wp = nullptr;
    tr.mark("  tables + block list", dev, st);
    if (int rc = wp_reserve(dev, st, (size_t)rows_w * nqp, &Wp)) return rc;
    if (v3)
        psb::v3_prep_w<JOB><<<(rows_w + 255) / 256, 256, 0, st>>>(Wp, rows_w, A.lenW,
            A.W[0], A.W[1], A.W[2], A.W[3], A.W[4], A.W[5], A.W[6], A.W[7]);
    else
        psb::v2_prep_w<<<(rows_w + 255) / 256, 256, 0, st>>>(Wp, rows_w, nqp, psb::job_nw(JOB), A.lenW,
            A.W[0], A.W[1], A.W[2], A.W[3], A.W[4], A.W[5], A.W[6], A.W[7]);
    CUDA_TRY(cudaGetLastError());
    tr.mark("  prep W' kernel", dev, st);
    int e = 0;
    if (kv == 4) {
        psb::V4Tables T{};
        T.G0 = t->seq[0] + psb::V4_PAD; T.G1 = t->seq[1] + psb::V4_PAD; T.G2 = t->seq[2] + psb::V4_PAD;
        T.H0 = t->seq[3] + psb::V4_PAD; T.H1 = t->seq[4] + psb::V4_PAD; T.nS = t->nS;
        T.blocks = bl.d; T.Wp = Wp;
        e = psb::launch_pair_v4<JOB>(A, T, bl.n, st);
    } else if (v3) {
        psb::V3Tables T{};
        T.gam = t->gam; T.igam = t->igam; T.INV = t->INV; T.nS = t->nS;
        T.blocks = bl.d; T.Wp = Wp;
        e = psb::launch_pair_v3<JOB>(A, T, bl.n, st);
    } else {
        psb::V2Tables T{};
        T.S = t->S; T.IS = t->IS; T.INV = t->INV; T.gam = t->gam; T.nS = t->nS;
        T.blocks = bl.d; T.Wp = Wp;
        e = psb::launch_pair_v2<JOB>(A, T, bl.n, st);
    }
    if (e != 0) return fail(ERR_CUDA, "pair kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    tr.mark("  pair kernel", dev, st);
    return OK;
}

// tiling parameters of a job as the tuned kernel runs it: l3 stride, pairs per thread, rows per warp
template <int JOB>
void job_tiling(int* ds, int* r, int* nr)
{
    if (kernel_version() == 2) { *ds = psb::v2_family(JOB) == psb::FAM_00 ? 2 : 1; *r = psb::v2_r(JOB); *nr = psb::v2_nr(JOB); }
    else { *ds = 2; *r = psb::v3_r(JOB); *nr = psb::v3_nr(JOB); }
}
int job_tiling_any(int job, int* ds, int* r, int* nr)
{
    using namespace psb;
    switch (job) {
        case JOB_M00: job_tiling<JOB_M00>(ds, r, nr); return OK;
        case JOB_M02: job_tiling<JOB_M02>(ds, r, nr); return OK;
        case JOB_MPP: job_tiling<JOB_MPP>(ds, r, nr); return OK;
        case JOB_MMM: job_tiling<JOB_MMM>(ds, r, nr); return OK;
        case JOB_MPPMMM: job_tiling<JOB_MPPMMM>(ds, r, nr); return OK;
        case JOB_TTTT: job_tiling<JOB_TTTT>(ds, r, nr); return OK;
        case JOB_EEEE: job_tiling<JOB_EEEE>(ds, r, nr); return OK;
        case JOB_TTTE: job_tiling<JOB_TTTE>(ds, r, nr); return OK;
        case JOB_TETE: job_tiling<JOB_TETE>(ds, r, nr); return OK;
        case JOB_TEEEP: job_tiling<JOB_TEEEP>(ds, r, nr); return OK;
        case JOB_TEEE: job_tiling<JOB_TEEE>(ds, r, nr); return OK;
        case JOB_TTEE: job_tiling<JOB_TTEE>(ds, r, nr); return OK;
        case JOB_MASTER: job_tiling<JOB_MASTER>(ds, r, nr); return OK;
    }
    return fail(ERR_ARG, "unknown job %d", job);
}

int launch_any(int job, const psb::PairArgs& A, cudaStream_t st)
{
    using namespace psb;
    switch (job) {
        case JOB_M00: return launch_job<JOB_M00>(A, st);
        case JOB_M02: return launch_job<JOB_M02>(A, st);
        case JOB_MPP: return launch_job<JOB_MPP>(A, st);
        case JOB_MMM: return launch_job<JOB_MMM>(A, st);
        case JOB_MPPMMM: return launch_job<JOB_MPPMMM>(A, st);
        case JOB_TTTT: return launch_job<JOB_TTTT>(A, st);
        case JOB_EEEE: return launch_job<JOB_EEEE>(A, st);
        case JOB_TTTE: return launch_job<JOB_TTTE>(A, st);
        case JOB_TETE: return launch_job<JOB_TETE>(A, st);
        case JOB_TEEEP: return launch_job<JOB_TEEEP>(A, st);
        case JOB_TEEE: return launch_job<JOB_TEEE>(A, st);
        case JOB_TTEE: return launch_job<JOB_TTEE>(A, st);
        case JOB_MASTER: return launch_job<JOB_MASTER>(A, st);
    }
    return fail(ERR_ARG, "unknown job %d", job);
}

const int kMcmJob[5] = {psb::JOB_M00, psb::JOB_M02, psb::JOB_MPP, psb::JOB_MMM, psb::JOB_MPPMMM};
const int kCovJob[7] = {psb::JOB_TTTT, psb::JOB_EEEE, psb::JOB_TTTE, psb::JOB_TETE,
                        psb::JOB_TEEEP, psb::JOB_TEEE, psb::JOB_TTEE};
const int kCovNeedSp[7] = {4, 4, 4, 4, 4, 4, 4};
const int kCovNeedRt[7] = {4, 4, 2, 2, 2, 2, 0};
const int kCovNeedW[7] = {8, 8, 4, 5, 4, 4, 2};

int check_common(int lmin, int lmax, long ld, int row_lo, int row_hi)
{
    if (lmin < 0 || lmax < lmin) return fail(ERR_ARG, "need 0 <= lmin <= lmax (got %d, %d)", lmin, lmax);
    if (lmax > 32767) return fail(ERR_ARG, "lmax %d above the supported 32767", lmax);
    if (ld < (long)(lmax - lmin + 1)) return fail(ERR_ARG, "leading dimension %ld < N=%d", ld, lmax - lmin + 1);
    if (row_lo < lmin || row_hi > lmax + 1 || row_lo > row_hi)
        return fail(ERR_ARG, "row band [%d,%d) outside [%d,%d]", row_lo, row_hi, lmin, lmax);
    return OK;
}

// ---------------------------------------------------------------------------------------
// per-device scratch for the host-level calls (grown on demand, kept between calls)
// ---------------------------------------------------------------------------------------
struct DeviceScratch {
    double* X[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t capX[5] = {0, 0, 0, 0, 0};
    double* vec = nullptr;      // packed input vectors
    size_t capVec = 0;
    double* T = nullptr;        // transposed block row of a band (multi-GPU host path)
    size_t capT = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;         // odd sub-bands of a host-level call (consecutive sub-bands overlap their tails)
    cudaStream_t copy_stream = nullptr;     // D2H of finished column bands, overlapped with compute
    cudaEvent_t ev_in = nullptr;            // inputs of the call are on the device
    cudaEvent_t ev[32] = {};
    char* ring = nullptr;                   // page-locked staging chunks of the pageable-destination delivery
    size_t ring_bytes = 0;
    cudaEvent_t ring_ev[32] = {};           // blocking-sync events, one per chunk
};
DeviceScratch g_scratch[16];

int scratch_reserve(int dev, int which, size_t n)
{
    DeviceScratch& s = g_scratch[dev];
    if (!s.stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        for (auto& e : s.ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (which < 5) {
        if (s.capX[which] < n) {
            if (s.X[which]) cudaFree(s.X[which]);
            s.X[which] = nullptr; s.capX[which] = 0;
            CUDA_TRY(cudaMalloc(&s.X[which], n * sizeof(double)));
            s.capX[which] = n;
        }
    } else if (which == 5) {
        if (s.capVec < n) {
            if (s.vec) cudaFree(s.vec);
            s.vec = nullptr; s.capVec = 0;
            CUDA_TRY(cudaMalloc(&s.vec, n * sizeof(double)));
            s.capVec = n;
        }
    } else {
        if (s.capT < n) {
            if (s.T) cudaFree(s.T);
            s.T = nullptr; s.capT = 0;
            CUDA_TRY(cudaMalloc(&s.T, n * sizeof(double)));
            s.capT = n;
        }
    }
    return OK;
}

int resolve_ngpus(int ngpus, int* out)
{
    const int have = device_count();
    if (have <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    if (ngpus < 0) return fail(ERR_ARG, "ngpus must be >= 0");
    if (ngpus == 0) ngpus = have;
    if (ngpus > have) return fail(ERR_ARG, "ngpus=%d but only %d device(s) visible", ngpus, have);
    if (ngpus > 16) return fail(ERR_ARG, "ngpus=%d above the supported 16", ngpus);
    *out = ngpus;
    return OK;
}

// Shared body of psb200_mcm / psb200_cov on host buffers.
//   vecs[k] / lens[k]: input vectors to upload, in the order W..., spectra..., ratios...
struct HostJob {
    int job;
    int lmin, lmax, lenW;
    int nW, nsp, nrt;
    const double* vecs[16];
    size_t lens[16];
    double* out[5];
    long ldo;
    int nout;
    int scale;
};

int run_on_device(const HostJob& hj, int dev, int row_lo, int row_hi, double* const* X, long ldX,
                  psb::PairArgs* keep = nullptr)
{
    // Upload the (small) inputs and launch stage 1 for one band on device `dev`.
    CUDA_TRY(cudaSetDevice(dev));
    DeviceScratch& s = g_scratch[dev];
    size_t tot = 0;
    const int nv = hj.nW + hj.nsp + hj.nrt;
    for (int k = 0; k < nv; ++k) tot += (hj.lens[k] + 3) & ~size_t(3);
    if (int rc = scratch_reserve(dev, 5, tot)) return rc;
    psb::PairArgs A{};
    A.lmin = hj.lmin; A.lmax = hj.lmax; A.lenW = hj.lenW;
    A.row_lo = row_lo; A.row_hi = row_hi; A.ld = ldX;
    A.out0 = X[0]; A.out1 = X[1]; A.out2 = X[2]; A.out3 = X[3]; A.out4 = X[4];
    size_t off = 0;
    for (int k = 0; k < nv; ++k) {
        CUDA_TRY(cudaMemcpyAsync(s.vec + off, hj.vecs[k], hj.lens[k] * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        const double* p = s.vec + off;
        if (k < hj.nW) A.W[k] = p;
        else if (k < hj.nW + hj.nsp) A.sp[k - hj.nW] = p;
        else A.rt[k - hj.nW - hj.nsp] = p;
        off += (hj.lens[k] + 3) & ~size_t(3);
    }
    if (keep) { *keep = A; return OK; }          // caller launches sub-bands itself
    return launch_any(hj.job, A, s.stream);
}

// Cost of row l1 as the default kernel executes it: one warp per NR rows x SPAN pairs of one parity of d
// (SPAN = (32/NR) R), stepping l3 by 2 from its first d to min(d + 2 l1, lenW-1) in lockstep -- SPAN-1 + l1 + 1 steps,
// cut at the window length.  A warp of s steps is charged s (1 + QUAD s) + OVH: OVH for the ring prologue and the
// epilogue, QUAD because long tiles pack worse into the last waves of a launch.  Both from a least-squares fit of
// t = a (sum s + OVH tiles + QUAD sum s^2) + c over the fourteen per-rank pair-kernel times of the 2-, 4- and 8-GPU
// runs of round 2 (profiles/r02_bench_n{2,4,8}.json: rms residual 0.11 ms of 11.4-46.2 ms; c = 0.5 ms per rank and step
// is the launch ramp of the five kernels and does not move the edges).  The tiling of the 8-accumulator covariance
// jobs is used for every job (they dominate a step).
static long double row_cost_raw(int l1, int lmax, int lenW)
{
    const long n = 2L * l1 + 1, D = lmax - l1;                  // family length, last d
    if (lenW <= 0) return (long double)n * (D + 1);              // full families (reference term count)
    constexpr long NRH = psb::v3_nr(psb::JOB_TTTT);
    constexpr long SPAN = psb::v3_span(psb::JOB_TTTT);           // the covariance jobs dominate a step
    constexpr long SKEW = SPAN - 1, OVH = 20;
    constexpr long double QUAD = 2.35e-5L;
    long double c = 0;
    for (long base = 0; base <= D; base += 2 * SPAN) {
        for (long par = 0; par < 2; ++par) {
            const long d_lo = base + par;
            if (d_lo > D) continue;
            const long last = (long)lenW - 1 - d_lo;
            const long steps = last < 0 ? 0 : std::min<long>(SKEW + l1, last / 2) + 1;
            c += ((long double)steps * (1.0L + QUAD * steps) + OVH) / NRH;   // a warp is shared by NR rows
        }
    }
    return c;
}

// Row costs are asked for on the critical path of every host-level call (sub-band split, band edges): 2 x 6144 rows x
// ~50 tiles of long-double arithmetic were 1.4 ms per call at lmax 6143 -- 7 ms of the 10 ms a benchmark step lost
// end to end against the resident kernels.  One vector per (lmax, lenW), computed once.
static const std::vector<long double>& row_costs(int lmax, int lenW)
{
    static std::mutex m;
    static std::map<std::pair<int, int>, std::vector<long double>> cache;
    std::lock_guard<std::mutex> lk(m);
    auto it = cache.find({lmax, lenW});
    if (it != cache.end()) return it->second;
    if (cache.size() > 64) cache.clear();
    std::vector<long double> c(lmax + 1);
    for (int l = 0; l <= lmax; ++l) c[l] = row_cost_raw(l, lmax, lenW);
    return cache.emplace(std::make_pair(lmax, lenW), std::move(c)).first->second;
}

// cost-balanced split of rows [a, b) into at most nsub consecutive pieces
static std::vector<int> split_rows(int a, int b, int lmax, int lenW, int nsub)
{
    std::vector<int> e{a};
    if (nsub <= 1 || b - a < 2 * nsub) { e.push_back(b); return e; }
    const std::vector<long double>& rc = row_costs(lmax, lenW);
    long double total = 0;
    for (int l = a; l < b; ++l) total += rc[l];
    long double run = 0;
    int k = 1;
    for (int l = a; l < b && k < nsub; ++l) {
        run += rc[l];
        if (run >= total * k / nsub) { if (l + 1 > e.back() && l + 1 < b) e.push_back(l + 1); ++k; }
    }
    e.push_back(b);
    return e;
}

// ---------------------------------------------------------------------------------------
// Delivery of a band's result to the caller's host array.
//
// A band hands over a list of 2-D copies (block columns straight from the slab, block rows from their transposed
// copies), each tied to the sub-band whose kernels produce it.  Into PAGE-LOCKED destinations they are issued as they
// are: cudaMemcpy2DAsync on the copy stream, DMA straight into the caller's array (55 GB/s on the B200 boxes).  Into
// PAGEABLE destinations -- what the reference allocates (spectralzeros, src/modecoupling.jl:199, src/covariance.jl:47)
// -- the CUDA runtime stages every copy through its own bounce buffer with one host thread (20 GB/s measured: 15 ms for
// one 302 MB matrix against 4.8 ms of TT kernels).  The staged path below does the bouncing itself: pieces of <= 8 MB are
// DMA'd densely into a ring of page-locked chunks and nthreads host workers scatter finished chunks into the caller's
// array (piece i belongs to worker i mod nthreads; a chunk is reused once its piece is scattered), so the DMA engine
// and several memcpy streams run side by side and behind the kernels of the later sub-bands.
// PSB200_STAGED=0 restores the runtime's own staging; PSB200_STAGE_THREADS / PSB200_STAGE_CHUNK_MB tune it.
// ---------------------------------------------------------------------------------------
struct Copy2D {
    double* dst; size_t dpitch;             // host; bytes between consecutive rows of the copy
    const double* src; size_t spitch;       // device (a host stand-in under the CPU test hook)
    size_t width, h;                        // bytes per row, rows
    int k;                                  // sub-band whose event the copy waits for
    // mirror delivery (see scatter_mirror): the copy is a block column holding RAW values below its diagonal block
    int mirror = 0, scale = 0, lmin = 0;
    int direct = 0;                         // mirror delivery of a symmetric copy into a page-locked array: DMA straight to dst
    double* base = nullptr; size_t ld = 0;  // the caller's matrix and its leading dimension (elements)
    size_t col0 = 0, row0 = 0, mrow = 0;    // matrix column of copy row 0, matrix row of element 0, first row below the diagonal block
};

// the copies of one band, in the order the sub-bands finish (shared by the direct and the staged delivery)
static std::vector<Copy2D> band_copies(int N, int lmin, int a, const std::vector<int>& sub, int nout, double* const* out, long ldo,
                                       double* const* X, long ldX, const double* T, const std::vector<size_t>& toff,
                                       bool mirror = false, int scale = 0)
{
    std::vector<Copy2D> cp;
    const int ns = (int)sub.size() - 1;
    for (int k = 0; k < ns; ++k) {
        const int sa = sub[k], sb = sub[k + 1], nbs = sb - sa;
        const int c0 = sb - lmin;                                  // first column right of this diagonal block
        const size_t r0 = (size_t)(sa - lmin);
        for (int o = 0; o < nout; ++o) {
            const double* slab = X[o] + (size_t)(sa - a) * ldX;     // row sa of the slab
            // block column: nbs columns of (N - r0) rows each
            cp.push_back({out[o] + r0 * ldo + r0, (size_t)ldo * sizeof(double), slab + r0, (size_t)ldX * sizeof(double),
                          (size_t)(N - r0) * sizeof(double), (size_t)nbs, k});
            if (mirror) {                    // the block row is written by the host from the same bytes
                Copy2D& c = cp.back();
                c.mirror = 1; c.scale = scale; c.lmin = lmin; c.base = out[o]; c.ld = (size_t)ldo;
                c.col0 = r0; c.row0 = r0; c.mrow = (size_t)c0;
                continue;
            }
            // block row: (N - c0) columns of nbs rows each
            if (c0 < N) {
                const double* Tk = T + toff[k] + (size_t)o * nbs * (N - c0);
                cp.push_back({out[o] + (size_t)c0 * ldo + r0, (size_t)ldo * sizeof(double), Tk, (size_t)nbs * sizeof(double),
                              (size_t)nbs * sizeof(double), (size_t)(N - c0), k});
            }
        }
    }
    return cp;
}

// cut every copy into pieces of at most chunk_bytes (whole rows)
static std::vector<Copy2D> split_copies(const std::vector<Copy2D>& cp, size_t chunk_bytes)
{
    std::vector<Copy2D> out;
    for (const Copy2D& c : cp) {
        if (c.width == 0 || c.h == 0) continue;
        const size_t rows = std::max<size_t>(1, chunk_bytes / c.width);
        for (size_t r = 0; r < c.h; r += rows) {
            Copy2D p = c;
            p.dst = (double*)((char*)c.dst + r * c.dpitch);
            p.src = (const double*)((const char*)c.src + r * c.spitch);
            p.h = std::min(rows, c.h - r);
            p.col0 = c.col0 + r;
            out.push_back(p);
        }
    }
    return out;
}

// One row of a piece into the caller's array with streaming (non-temporal) stores: the destination is written once and
// not read again by this call, so there is no point in reading its cache lines first (a plain memcpy of a 3-49 KB row
// does: read-for-ownership, a third more memory traffic).  PSB200_STAGE_NT=0 selects memcpy.
static inline void copy_row_nt(char* dst, const char* src, size_t n)
{
    size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    if (head > n) head = n;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
        _mm_stream_si128((__m128i*)(dst + i), a);
        _mm_stream_si128((__m128i*)(dst + i + 16), b);
        _mm_stream_si128((__m128i*)(dst + i + 32), c);
        _mm_stream_si128((__m128i*)(dst + i + 48), d);
    }
    for (; i + 16 <= n; i += 16) _mm_stream_si128((__m128i*)(dst + i), _mm_loadu_si128((const __m128i*)(src + i)));
    if (i < n) memcpy(dst + i, src + i, n - i);
}

// MIRROR DELIVERY (PSB200_MIRROR=1; off by default).  Every result of this path is symmetric up to a column scaling:
// C[l2,l1] = C[l1,l2] (src/covariance.jl:119), M[l1,l2] = (2 l2 + 1) Xi, M[l2,l1] = (2 l1 + 1) Xi (src/modecoupling.jl:90-91).
// The standard delivery sends both triangles over PCIe (block column + transposed block row of every sub-band).  When
// several GPUs write into one host, what bounds a call is the rate at which the host absorbs DMA writes (65-73 GB/s for
// eight B200s), so here only the block columns cross PCIe -- with the entries below the diagonal
// block left RAW (x = Xi or C) -- and the scatter workers write both sides from the same bytes:
//     A[i, j] = s1 x,  s1 = 2 (lmin + j) + 1     (the block column, as band_transpose_kernel scales it in place)
//     A[j, i] = s2 x,  s2 = 2 (lmin + i) + 1     (the block row,    as band_transpose_kernel writes it to T)
// (s1 = s2 = 1 for covariance blocks): the same IEEE products the device forms, so the result is bit-identical.  The rows
// of the diagonal block arrive finished (finish_kernel) and are copied as they are.  The transposed side is written as
// contiguous runs: 8 matrix columns i at a time, each a run over the piece's h consecutive j, staged in a small buffer.
// For covariance blocks (no scaling) going into a PAGE-LOCKED array the block column needs no host thread at all: it is
// DMA'd straight to its place (p.direct) and the workers read it back from there for the symmetric side only.
static void scatter_mirror(const Copy2D& p, const char* chunk, bool nt)
{
    const size_t W = p.width / sizeof(double);                  // matrix rows row0 .. row0 + W - 1 of every column
    const size_t na = p.mrow > p.row0 ? std::min(W, p.mrow - p.row0) : 0;     // rows of the diagonal block
    const size_t h = p.h;
    const size_t pitch = p.direct ? p.dpitch : p.width;         // bytes between consecutive matrix columns of the source
    if (p.direct) chunk = (const char*)p.dst;
    // the block column
    for (size_t jj = 0; jj < h && !p.direct; ++jj) {
        const double* src = (const double*)(chunk + jj * p.width);
        double* dst = p.base + p.row0 + (p.col0 + jj) * p.ld;
        if (!p.scale) {
            if (nt) copy_row_nt((char*)dst, (const char*)src, p.width); else memcpy(dst, src, p.width);
        } else {
            if (na) memcpy(dst, src, na * sizeof(double));
            const double s1 = (double)(2 * ((long)p.lmin + (long)(p.col0 + jj)) + 1);
            if (!nt) {
                for (size_t t = na; t < W; ++t) dst[t] = s1 * src[t];
            } else {                              // scaled through a small buffer so that the stores can stream
                double buf[512];
                for (size_t t = na; t < W; t += 512) {
                    const size_t n = std::min<size_t>(512, W - t);
                    for (size_t u = 0; u < n; ++u) buf[u] = s1 * src[t + u];
                    copy_row_nt((char*)(dst + t), (const char*)buf, n * sizeof(double));
                }
            }
        }
    }
    // the block row: A[col0 + jj, i] for i = mrow .. row0 + W - 1.  Tiles of TI matrix columns i x TJ rows jj through a
    // 64 KB buffer: the source is read as TI consecutive doubles per jj (the lines of the next jj are prefetched: at a
    // stride of W doubles the hardware prefetchers see nothing), the destination is written as TI runs of TJ doubles.
    // MEASURED (host microbenchmark, one thread, 8 MB piece): 1.5 GB/s with buffer rows exactly TJ doubles apart (2 KB: the
    // TI store streams of the gather fall into two L1 sets), 4.8 GB/s with the rows 8 doubles further apart; block column 8-11 GB/s
    constexpr size_t TI = 32, TJ = 256, TS = TJ + 8;
    double tmp[TI * TS];
    for (size_t j0 = 0; j0 < h; j0 += TJ) {
        const size_t nj = std::min(TJ, h - j0);
        for (size_t t0 = na; t0 < W; t0 += TI) {
            const size_t nd = std::min(TI, W - t0);
            double s2[TI];
            for (size_t d = 0; d < nd; ++d) s2[d] = p.scale ? (double)(2 * ((long)p.lmin + (long)(p.row0 + t0 + d)) + 1) : 1.0;
            for (size_t jj = 0; jj < nj; ++jj) {
                const double* src = (const double*)(chunk + (j0 + jj) * pitch) + t0;
                if (jj + 1 < nj) {
                    const char* nx = (const char*)src + pitch;
                    for (size_t b = 0; b < nd * sizeof(double); b += 64) _mm_prefetch(nx + b, _MM_HINT_T0);
                }
                for (size_t d = 0; d < nd; ++d) tmp[d * TS + jj] = s2[d] * src[d];
            }
            for (size_t d = 0; d < nd; ++d) {
                double* dst = p.base + (p.col0 + j0) + (p.row0 + t0 + d) * p.ld;
                if (nt) copy_row_nt((char*)dst, (const char*)&tmp[d * TS], nj * sizeof(double));
                else memcpy(dst, &tmp[d * TS], nj * sizeof(double));
            }
        }
    }
    if (nt) _mm_sfence();
}

// The pipeline: issue(i, chunk, c) starts the dense copy of piece i into chunk c and marks its completion, wait(c) blocks
// until that copy has landed.  Both return 0 or an error code (message in the calling thread's g_err).
template <class Issue, class Wait>
static int deliver_staged(const std::vector<Copy2D>& pieces, char* ring, size_t chunk_bytes, int nch, int nthreads,
                          Issue issue, Wait wait)
{
    const int P = (int)pieces.size();
    if (P == 0) return OK;
    nthreads = std::max(1, std::min(nthreads, P));
    std::mutex m;
    std::condition_variable cv;
    int issued = 0, rc_shared = OK;
    bool stop = false;
    std::string err;
    std::vector<char> busy(nch, 0);                    // chunk holds a piece that is not scattered yet
    const char* nt_env = getenv("PSB200_STAGE_NT");
    const bool nt = !(nt_env && nt_env[0] == '0');
    auto give_up = [&](int rc) {
        std::lock_guard<std::mutex> lk(m);
        if (rc_shared == OK) { rc_shared = rc; err = g_err; }
        stop = true;
        cv.notify_all();
    };
    auto worker = [&](int t) {
        for (int i = t; i < P; i += nthreads) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [&] { return issued > i || stop; });
                if (issued <= i) return;               // stopped before piece i was issued
            }
            const int c = i % nch;
            if (int rc = wait(c)) { give_up(rc); return; }
            const Copy2D& p = pieces[i];
            const char* src = ring + (size_t)c * chunk_bytes;
            if (p.mirror) {
                scatter_mirror(p, src, nt);
            } else if (nt) {
                for (size_t r = 0; r < p.h; ++r) copy_row_nt((char*)p.dst + r * p.dpitch, src + r * p.width, p.width);
                _mm_sfence();
            } else if (p.dpitch == p.width) memcpy(p.dst, src, p.width * p.h);
            else for (size_t r = 0; r < p.h; ++r) memcpy((char*)p.dst + r * p.dpitch, src + r * p.width, p.width);
            {
                std::lock_guard<std::mutex> lk(m);
                busy[c] = 0;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
    for (int i = 0; i < P; ++i) {
        const int c = i % nch;
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return !busy[c] || stop; });
            if (stop) break;
            busy[c] = 1;
        }
        if (int rc = issue(i, ring + (size_t)c * chunk_bytes, c)) { give_up(rc); break; }
        {
            std::lock_guard<std::mutex> lk(m);
            issued = i + 1;
        }
        cv.notify_all();
    }
    for (auto& t : th) t.join();
    if (rc_shared != OK) g_err = err;
    return rc_shared;
}

static size_t stage_chunk_bytes()
{
    size_t mb = 8;
    if (const char* e = getenv("PSB200_STAGE_CHUNK_MB")) mb = (size_t)std::max(1, std::min(64, atoi(e)));
    return mb << 20;
}

static int stage_threads(int ngpus_in_call, bool mirror = false)
{
    if (const char* e = getenv("PSB200_STAGE_THREADS")) return std::max(1, std::min(28, atoi(e)));
    const int hw = (int)std::thread::hardware_concurrency();
    // MEASURED (1 B200, 16 host threads, lmax 6143, ms per call; page-locked destination 7.0 / 22.3, CUDA runtime's bounce
    // copies 33.3 / 70.0 for TT / fused EE-BB): 2 workers 36.9 / 75.3, 4: 20.2 / 38.0, 8: 13.4 / 27.9, 12: 11.8 / 24.5
    if (mirror) return std::max(2, std::min(12, hw / std::max(1, ngpus_in_call)));   // the workers do the transposing: every core
    return std::max(2, std::min(12, 3 * hw / (4 * std::max(1, ngpus_in_call))));
}

// is p ordinary (pageable) host memory?  Page-locked, registered and managed memory all answer otherwise.
static bool is_pageable(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// page-locked ring of device g: nch chunks of chunk_bytes and their events
static int stage_ring_reserve(int g, size_t chunk_bytes, int nch)
{
    DeviceScratch& s = g_scratch[g];
    const size_t need = chunk_bytes * (size_t)nch;
    if (s.ring_bytes < need) {
        if (s.ring) cudaFreeHost(s.ring);
        s.ring = nullptr; s.ring_bytes = 0;
        CUDA_TRY(cudaHostAlloc((void**)&s.ring, need, cudaHostAllocPortable));
        s.ring_bytes = need;
    }
    for (int c = 0; c < nch; ++c)
        if (!s.ring_ev[c]) CUDA_TRY(cudaEventCreateWithFlags(&s.ring_ev[c], cudaEventDisableTiming | cudaEventBlockingSync));
    return OK;
}

// A contiguous device -> host copy of `bytes` on copy stream `cs` of device g, blocking until the bytes are in `hdst`:
// staged (see above) when the destination is pageable and large, one cudaMemcpyAsync + synchronize otherwise.
static int download_blocking(int g, void* hdst, const void* dsrc, size_t bytes, cudaStream_t cs, int ngpus_in_call)
{
    const char* st_env = getenv("PSB200_STAGED");
    if ((st_env && st_env[0] == '0') || bytes < (size_t(4) << 20) || !is_pageable(hdst)) {
        CUDA_TRY(cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaStreamSynchronize(cs));
        return OK;
    }
    DeviceScratch& s = g_scratch[g];
    const size_t chunk = stage_chunk_bytes();
    const int nthr = stage_threads(ngpus_in_call), nch = std::min(32, 2 * nthr);
    if (int rc = stage_ring_reserve(g, chunk, nch)) return rc;
    std::vector<Copy2D> pieces;
    for (size_t off = 0; off < bytes; off += chunk) {
        const size_t w = std::min(chunk, bytes - off);
        pieces.push_back({(double*)((char*)hdst + off), w, (const double*)((const char*)dsrc + off), w, w, 1, 0});
    }
    auto issue = [&](int i, char* dst, int c) -> int {
        CUDA_TRY(cudaMemcpyAsync(dst, pieces[i].src, pieces[i].width, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaEventRecord(s.ring_ev[c], cs));
        return OK;
    };
    auto wait = [&](int c) -> int {
        CUDA_TRY(cudaSetDevice(g));
        CUDA_TRY(cudaEventSynchronize(s.ring_ev[c]));
        return OK;
    };
    return deliver_staged(pieces, s.ring, chunk, nch, nthr, issue, wait);
}

// A contiguous host -> device copy of `bytes`, queued on stream `st` of device g (the copies are in the stream when the
// call returns; the source may be reused only after the stream has passed them -- every caller synchronises before it
// returns to the user).  Pageable sources are gathered into the page-locked ring by nthreads workers, each of which owns
// two chunks and issues its own pieces: fill chunk, cudaMemcpyAsync, record the chunk's event, and wait for that event
// before the chunk is filled again.  (The CUDA runtime's own path for a pageable source is one thread and one bounce
// buffer: three 403 MB maps were about 110 ms of the 316 ms of a host-level map2alm at nside 2048.)
static int upload_async(int g, void* ddst, const void* hsrc, size_t bytes, cudaStream_t st, int ngpus_in_call)
{
    const char* st_env = getenv("PSB200_STAGED");
    if ((st_env && st_env[0] == '0') || bytes < (size_t(4) << 20) || !is_pageable(hsrc)) {
        CUDA_TRY(cudaMemcpyAsync(ddst, hsrc, bytes, cudaMemcpyHostToDevice, st));
        return OK;
    }
    DeviceScratch& s = g_scratch[g];
    const size_t chunk = stage_chunk_bytes();
    const int nthr = stage_threads(ngpus_in_call), nch = std::min(32, 2 * nthr);
    if (int rc = stage_ring_reserve(g, chunk, nch)) return rc;
    const int nw = nch / 2;                                       // worker t owns chunks t and t + nw
    const size_t P = (bytes + chunk - 1) / chunk;
    std::vector<int> rcs(nw, OK);
    std::vector<std::string> errs(nw);
    auto worker = [&](int t) {
        auto body = [&]() -> int {
            CUDA_TRY(cudaSetDevice(g));
            for (size_t i = t; i < P; i += nw) {
                const int c = (int)(i % (size_t)nch);               // = t or t + nw: i = t (mod nw)
                char* buf = s.ring + (size_t)c * chunk;
                const size_t off = i * chunk, w = std::min(chunk, bytes - off);
                CUDA_TRY(cudaEventSynchronize(s.ring_ev[c]));       // the chunk's previous piece has left (no-op on a fresh event)
                memcpy(buf, (const char*)hsrc + off, w);
                CUDA_TRY(cudaMemcpyAsync((char*)ddst + off, buf, w, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaEventRecord(s.ring_ev[c], st));
            }
            return OK;
        };
        rcs[t] = body();
        if (rcs[t] != OK) errs[t] = g_err;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nw; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& t : th) t.join();
    for (int t = 0; t < nw; ++t)
        if (rcs[t] != OK) { g_err = errs[t]; return rcs[t]; }
    return OK;
}

int run_band_on_device(const HostJob& hj, int g, int a, int b, std::string* err, int ngpus_in_call = 1)
{
    auto body = [&]() -> int {
        const int N = hj.lmax - hj.lmin + 1;
        const long ldX = N;
        const int nb = b - a;
        if (nb <= 0) return OK;
        CUDA_TRY(cudaSetDevice(g));
        for (int o = 0; o < hj.nout; ++o)
            if (int rc = scratch_reserve(g, o, (size_t)nb * N)) return rc;
        // sub-bands: each is a complete band of its own (L-shaped region), so its copies can start
        // while the next sub-band computes
        // MEASURED (1 GPU, lmax 6143, e2e ms per bench step over the 89.0 ms of the resident kernels): 4 sub-bands 104.1,
        // 8: 99.3, 16: 96.3 -- what a call exposes is the delivery of its last pieces, which shrinks with the piece
        int nsub = nb >= 4096 ? 16 : (nb >= 2048 ? 8 : (nb >= 512 ? 4 : (nb >= 128 ? 2 : 1)));
        if (const char* e = getenv("PSB200_NSUB")) nsub = std::max(1, std::min(32, atoi(e)));
        const std::vector<int> sub = split_rows(a, b, hj.lmax, hj.lenW, nsub);
        // (Measured and dropped: when one device owns the whole matrix, finishing both triangles in place sub-band by
        // sub-band and copying full columns as contiguous runs.  The L-shaped regions front-load the bytes -- the first
        // sub-bands carry the long columns -- while full columns leave a fifth of the matrix behind the last kernel:
        // 97.2 vs 94.6 ms per step end to end at lmax 6143.)
        const int ns = (int)sub.size() - 1;
        std::vector<size_t> toff(ns + 1, 0);             // one transposed block row per (sub-band, output)
        for (int k = 0; k < ns; ++k)
            toff[k + 1] = toff[k] + (size_t)(sub[k + 1] - sub[k]) * (size_t)(N - (sub[k + 1] - hj.lmin)) * hj.nout;
        // mirror delivery (scatter_mirror): no transposed block rows on the device, the host writes them
        const char* mir_env = getenv("PSB200_MIRROR");
        const bool mirror = mir_env && mir_env[0] == '1' && (size_t)nb * N * sizeof(double) >= (size_t(4) << 20);
        if (toff[ns] && !mirror) if (int rc = scratch_reserve(g, 6, toff[ns])) return rc;
        DeviceScratch& s = g_scratch[g];
        double* Xs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        const long rowoff = (long)(a - hj.lmin);
        for (int o = 0; o < hj.nout; ++o) Xs[o] = s.X[o] - rowoff * ldX;   // kernels index rows from lmin
        psb::PairArgs A{};
        if (int rc = run_on_device(hj, g, a, b, Xs, ldX, &A)) return rc;
        // Sub-bands alternate between two streams: they touch disjoint rows, so the blocks of sub-band k+1 can fill the SMs
        // that the last blocks of sub-band k leave idle (measured at lmax 6143 on one GPU: no difference, 99.4 vs 99.2 ms
        // per step -- the drains are short; kept because it costs nothing).
        CUDA_TRY(cudaEventRecord(s.ev_in, s.stream));
        CUDA_TRY(cudaStreamWaitEvent(s.stream2, s.ev_in, 0));
        Trace tr;
        for (int k = 0; k < ns; ++k) {
            const int sa = sub[k], sb = sub[k + 1], nbs = sb - sa;
            cudaStream_t sk = (k & 1) && !getenv("PSB200_ONE_STREAM") ? s.stream2 : s.stream;
            A.row_lo = sa; A.row_hi = sb;
            if (int rc = launch_any(hj.job, A, sk)) return rc;
            const int nt = (nbs + 31) / 32;
            const int c0 = sb - hj.lmin;                  // first column right of this diagonal block
            for (int o = 0; o < hj.nout; ++o) {
                double* slab = s.X[o] + (size_t)(sa - a) * ldX;         // row sa of the slab
                finish_kernel<<<dim3(nt, nt), 256, 0, sk>>>(slab + (sa - hj.lmin), ldX, sa, nbs, hj.scale, 0);
                CUDA_TRY(cudaGetLastError());
                if (c0 < N && !mirror) {
                    double* Tk = s.T + toff[k] + (size_t)o * nbs * (N - c0);
                    band_transpose_kernel<<<dim3((N - c0 + 31) / 32, nt), 256, 0, sk>>>(
                        slab, ldX, Tk, nbs, sa, hj.lmin, c0, N, hj.scale);
                    CUDA_TRY(cudaGetLastError());
                }
            }
            CUDA_TRY(cudaEventRecord(s.ev[k], sk));
        }

        // everything is queued; now the delivery on the copy stream (see "Delivery" above)
        const std::vector<Copy2D> copies = band_copies(N, hj.lmin, a, sub, hj.nout, hj.out, hj.ldo, s.X, ldX, s.T, toff,
                                                       mirror, hj.scale);
        size_t bytes = 0;
        for (const Copy2D& c : copies) bytes += c.width * c.h;
        const char* st_env = getenv("PSB200_STAGED");
        const bool staged = mirror || (!(st_env && st_env[0] == '0') && bytes >= (size_t(4) << 20) && is_pageable(hj.out[0]));
        if (!staged) {
            int waited = -1;
            for (const Copy2D& c : copies) {
                if (c.k != waited) { CUDA_TRY(cudaStreamWaitEvent(s.copy_stream, s.ev[c.k], 0)); waited = c.k; }
                CUDA_TRY(cudaMemcpy2DAsync(c.dst, c.dpitch, c.src, c.spitch, c.width, c.h, cudaMemcpyDeviceToHost, s.copy_stream));
            }
        } else {
            const size_t chunk = stage_chunk_bytes();
            const int nthr = stage_threads(ngpus_in_call, mirror), nch = std::min(32, nthr + 4);   // chunks: one per worker + 4 in flight
            if (int rc = stage_ring_reserve(g, chunk, nch)) return rc;
            std::vector<Copy2D> pieces = split_copies(copies, chunk);
            if (mirror && hj.scale == 0 && !is_pageable(hj.out[0]))
                for (Copy2D& p : pieces) p.direct = 1;
            int waited = -1;
            auto issue = [&](int i, char* dst, int c) -> int {
                const Copy2D& p = pieces[i];
                if (p.k != waited) { CUDA_TRY(cudaStreamWaitEvent(s.copy_stream, s.ev[p.k], 0)); waited = p.k; }
                if (p.direct)
                    CUDA_TRY(cudaMemcpy2DAsync(p.dst, p.dpitch, p.src, p.spitch, p.width, p.h, cudaMemcpyDeviceToHost, s.copy_stream));
                else
                CUDA_TRY(cudaMemcpy2DAsync(dst, p.width, p.src, p.spitch, p.width, p.h, cudaMemcpyDeviceToHost, s.copy_stream));
                CUDA_TRY(cudaEventRecord(s.ring_ev[c], s.copy_stream));
                return OK;
            };
            auto wait = [&](int c) -> int {                   // runs on the worker threads
                CUDA_TRY(cudaSetDevice(g));
                CUDA_TRY(cudaEventSynchronize(s.ring_ev[c]));
                return OK;
            };
            if (int rc = deliver_staged(pieces, s.ring, chunk, nch, nthr, issue, wait)) return rc;
        }
        tr.mark("   band: kernels + finish + transpose", g, s.stream);
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream2));
        CUDA_TRY(cudaStreamSynchronize(s.copy_stream));
        tr.mark("   band: tail of D2H", g, s.copy_stream);
        return OK;
    };
    const int rc = body();
    if (rc != OK && err) *err = g_err;        // g_err is thread-local: hand the message to the caller
    return rc;
}

// FP64 instructions per pair-step of the default kernel (profiles/r02_sass_fp64.json), for the time model below.
static double job_fp64_per_pair_step(int job)
{
    using namespace psb;
    switch (job) {
        case JOB_M00: return 2; case JOB_M02: return 5; case JOB_MPP: return 6; case JOB_MMM: return 5;
        case JOB_MPPMMM: return 11; case JOB_TTTT: return 9; case JOB_EEEE: return 13; case JOB_TTTE: return 5;
        case JOB_TETE: return 9; case JOB_TEEEP: return 9; case JOB_TEEE: return 8; case JOB_TTEE: return 3;
        case JOB_MASTER: return 14;
    }
    return 9;
}

// Bands of the HOST-level multi-GPU call.  Every device delivers its own L-shaped region of the result straight to
// the caller's array while it computes the next sub-band, so its time is max(kernel time, D2H time) -- and the two
// do not balance alike: the low rows are cheap to compute but long (band 0 of an 8-way kernel-balanced split owns
// 43 % of the bytes: 14 ms of copies against 11 ms of kernels per step at lmax 6143, the exposed D2H of the round-1
// 8-GPU line).  Rows carry two weights, kernel seconds (row_cost x the job's FP64 work per pair-step / the sustained
// FP64 issue rate) and copy seconds (bytes of the row's part of the L region / the PCIe rate); the split minimises the
// largest max(sum kernel + a quarter of sum copy, sum copy) over contiguous bands (bisection on that bound, greedy
// feasibility).
// PSB200_HOST_SPLIT=kernel restores the kernel-balanced edges of psb200_band_edges.
static void host_band_edges(const HostJob& hj, int nb, int* edges)
{
    const char* mode = getenv("PSB200_HOST_SPLIT");
    if (mode && strcmp(mode, "kernel") == 0) { psb200_band_edges(hj.lmin, hj.lmax, hj.lenW, nb, edges); return; }
    const int N = hj.lmax - hj.lmin + 1;
    constexpr double kIssue = 1.83e13 * 0.82;                 // FP64 lane-instructions per second, as measured
    constexpr double kPcie = 50e9;                            // bytes per second per device, pinned destination
    // row_cost sums to warp-steps of the TTTT tiling (a warp's steps are shared out over its NR rows): 32 R pair-steps each
    const double per_unit = 32.0 * psb::v3_r(psb::JOB_TTTT) * job_fp64_per_pair_step(hj.job) / kIssue;   // seconds per unit
    // the edges of a (job, shape, nb) never change: computed once
    static std::mutex cm;
    static std::map<std::tuple<int, int, int, int, int, int>, std::vector<int>> done;
    const auto key = std::make_tuple(hj.job, hj.nout, hj.lmin, hj.lmax, hj.lenW, nb);
    {
        std::lock_guard<std::mutex> lk(cm);
        auto it = done.find(key);
        if (it != done.end()) { for (int b = 0; b <= nb; ++b) edges[b] = it->second[b]; return; }
    }
    const std::vector<long double>& rcv = row_costs(hj.lmax, hj.lenW);
    std::vector<double> c(N), d(N);
    double sc = 0, sd = 0;
    for (int i = 0; i < N; ++i) {
        c[i] = (double)rcv[hj.lmin + i] * per_unit;
        d[i] = 8.0 * hj.nout * (2.0 * (N - i) - 1.0) / kPcie;
        sc += c[i]; sd += d[i];
    }
    auto bands_needed = [&](double lam, int* e) {
        int k = 0, i = 0;
        if (e) e[0] = hj.lmin;
        while (i < N) {
            double a = 0, b = 0;
            int j = i;
            // time of a band: its copies run behind its kernels sub-band by sub-band, the last quarter or so is exposed
            while (j < N && a + c[j] + 0.25 * (b + d[j]) <= lam && b + d[j] <= lam) { a += c[j]; b += d[j]; ++j; }
            if (j == i) ++j;                                  // a single row above the bound: its own band
            ++k;
            if (e && k <= nb) e[k] = hj.lmin + j;
            i = j;
        }
        return k;
    };
    double lo = std::max(sc, sd) / nb * 0.5, hi = std::max(sc, sd);
    for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (bands_needed(mid, nullptr) <= nb) hi = mid; else lo = mid;
    }
    std::vector<int> e(nb + 1, hj.lmax + 1);
    const int k = bands_needed(hi, e.data());
    for (int b = k + 1; b <= nb; ++b) e[b] = hj.lmax + 1;     // fewer bands than devices: the rest stay empty
    e[nb] = hj.lmax + 1;
    for (int b = 0; b <= nb; ++b) edges[b] = e[b];
    std::lock_guard<std::mutex> lk(cm);
    if (done.size() > 256) done.clear();
    done[key] = e;
}

int run_host_job(const HostJob& hj, int ngpus)
{
    int cur = 0;
    cudaGetDevice(&cur);
    if (ngpus == 1) {                                   // same scheme, one band, calling thread, the CALLER's current device
        const int rc = run_band_on_device(hj, cur, hj.lmin, hj.lmax + 1, nullptr);
        cudaSetDevice(cur);
        return rc;
    }
    Trace tr;
    std::vector<int> edges(ngpus + 1);
    host_band_edges(hj, ngpus, edges.data());
    if (tr.on) {
        fprintf(stderr, "[psb200] host bands:");
        for (int g = 0; g <= ngpus; ++g) fprintf(stderr, " %d", edges[g]);
        fprintf(stderr, "\n");
    }
    std::vector<int> rcs(ngpus, OK);
    std::vector<std::string> errs(ngpus);
    std::vector<std::thread> th;
    for (int g = 1; g < ngpus; ++g)
        th.emplace_back([&, g] { rcs[g] = run_band_on_device(hj, g, edges[g], edges[g + 1], &errs[g], ngpus); });
    rcs[0] = run_band_on_device(hj, 0, edges[0], edges[1], &errs[0], ngpus);
    for (auto& t : th) t.join();
    tr.mark("all bands delivered", 0, g_scratch[0].stream);
    cudaSetDevice(cur);
    for (int g = 0; g < ngpus; ++g)
        if (rcs[g] != OK) { g_err = "device " + std::to_string(g) + ": " + errs[g]; return rcs[g]; }
    return OK;
}


// ---------------------------------------------------------------------------------------
// QuickPol Xi (psb200_quickpol.cuh): columns l of the band storage are independent, so the multi-GPU
// split is by contiguous column bands of equal cost (a column of nb band rows costs ~ nb * l
// recurrence steps) and every device moves its own columns to and from the caller's array.
// ---------------------------------------------------------------------------------------
int check_quickpol(int lmax, int lenW, int band_lo, int band_hi, long ldb, int col_lo, int col_hi)
{
    // 12287: the bound up to which the rescaling cadence of the sweep was analysed (psb200_quickpol.cuh)
    if (lmax < 0 || lmax > 12287) return fail(ERR_ARG, "need 0 <= lmax <= 12287 (got %d)", lmax);
    if (lenW < 1) return fail(ERR_ARG, "empty scan spectrum W");
    // BandedMatrices accepts bandwidths beyond the matrix size (the extra storage rows are padding): so do we
    if (band_lo < 0 || band_hi < 0 || band_lo > (1 << 20) || band_hi > (1 << 20))
        return fail(ERR_ARG, "band widths (%d, %d) must be non-negative (and below 2^20)", band_lo, band_hi);
    if (ldb < (long)band_lo + band_hi + 1) return fail(ERR_ARG, "leading dimension %ld < band_lo+band_hi+1", ldb);
    if (col_lo < 0 || col_hi > lmax + 1 || col_lo > col_hi)
        return fail(ERR_ARG, "column band [%d,%d) outside [0,%d]", col_lo, col_hi, lmax);
    return OK;
}

// PSB200_QP=simple selects the table-free variant of the QuickPol kernel (one rsqrt per family and a
// reciprocal per step); default: the tabulated variant (one rsqrt per step for both families).
bool quickpol_tabulated()
{
    const char* e = getenv("PSB200_QP");
    return !(e && strcmp(e, "simple") == 0);
}

int launch_quickpol(psb::QpArgs A, cudaStream_t st)
{
    const int ncol = A.col_hi - A.col_lo;
    if (ncol <= 0) return OK;
    const int nb = A.band_lo + A.band_hi + 1;
#if PSB200_QP_FLAT
    dim3 grid((unsigned)(((long)ncol * nb + psb::QP_THREADS - 1) / psb::QP_THREADS));
#else
    dim3 grid(ncol, (nb + psb::QP_THREADS - 1) / psb::QP_THREADS);
#endif
    if (!quickpol_tabulated()) {
        psb::quickpol_kernel<false><<<grid, psb::QP_THREADS, 0, st>>>(A);
        CUDA_TRY(cudaGetLastError());
        return OK;
    }
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", dev);
    const int nT = 2 * A.lmax + 2;                     // j = 0 .. 2 lmax + 1
    double* buf = nullptr;
    if (int rc = wp_reserve(dev, st, (size_t)5 * nT, &buf)) return rc;
    psb::QpD2* bb0 = reinterpret_cast<psb::QpD2*>(buf);
    psb::QpD2* bb1 = bb0 + nT;
    double* ij2 = buf + (size_t)4 * nT;
    psb::quickpol_tables_kernel<<<(nT + 255) / 256, 256, 0, st>>>(ij2, bb0, bb1, nT, A.s1 + A.nu1, A.s2 + A.nu2);
    CUDA_TRY(cudaGetLastError());
    A.T.IJ2 = ij2; A.T.BB0 = bb0; A.T.BB1 = bb1;
    psb::quickpol_kernel<true><<<grid, psb::QP_THREADS, 0, st>>>(A);
    CUDA_TRY(cudaGetLastError());
    return OK;
}

struct QpHostJob {
    psb::QpArgs A;              // W / Xb filled per device
    const double* W;
    double* Xb;
};

int run_quickpol_on_device(const QpHostJob& hj, int g, int a, int b, std::string* err)
{
    auto body = [&]() -> int {
        if (b <= a) return OK;
        CUDA_TRY(cudaSetDevice(g));
        const int nb = hj.A.band_lo + hj.A.band_hi + 1;
        const size_t nW = ((size_t)hj.A.lenW + 3) & ~size_t(3);
        if (int rc = scratch_reserve(g, 5, nW)) return rc;
        if (int rc = scratch_reserve(g, 0, (size_t)nb * (b - a))) return rc;
        DeviceScratch& s = g_scratch[g];
        CUDA_TRY(cudaMemcpyAsync(s.vec, hj.W, (size_t)hj.A.lenW * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        // entries the reference loop does not visit (rows l'' < 2 inside the band) must keep the caller's
        // values, so the slab starts as a copy of the caller's columns
        double* host_cols = hj.Xb + (size_t)a * hj.A.ldb;
        CUDA_TRY(cudaMemcpy2DAsync(s.X[0], (size_t)nb * sizeof(double), host_cols, (size_t)hj.A.ldb * sizeof(double),
                                   (size_t)nb * sizeof(double), b - a, cudaMemcpyHostToDevice, s.stream));
        psb::QpArgs A = hj.A;
        A.W = s.vec;
        A.ldb = nb;
        A.Xb = s.X[0] - (long)a * nb;          // the kernel indexes columns from l = 0
        A.col_lo = a; A.col_hi = b;
        if (int rc = launch_quickpol(A, s.stream)) return rc;
        CUDA_TRY(cudaMemcpy2DAsync(host_cols, (size_t)hj.A.ldb * sizeof(double), s.X[0], (size_t)nb * sizeof(double),
                                   (size_t)nb * sizeof(double), b - a, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        return OK;
    };
    const int rc = body();
    if (rc != OK && err) *err = g_err;
    return rc;
}

#include "psb200_solve.inl"
#include "psb200_sht.inl"

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* psb200_last_error(void) { return g_err.c_str(); }
const char* psb200_version(void) { return "psb200 0.1 (sm_100a)"; }
int psb200_device_count(void) { return device_count(); }

int psb200_band_edges(int lmin, int lmax, int lenW, int nbands, int* edges)
{
    if (lmin < 0 || lmax < lmin || nbands < 1 || !edges) return fail(ERR_ARG, "band_edges: bad arguments");
    const std::vector<long double>& rc = row_costs(lmax, lenW);
    long double total = 0;
    for (int l = lmin; l <= lmax; ++l) total += rc[l];
    edges[0] = lmin;
    long double run = 0;
    int b = 1;
    for (int l = lmin; l <= lmax && b < nbands; ++l) {
        run += rc[l];
        while (b < nbands && run >= total * b / nbands) edges[b++] = l + 1;
    }
    while (b <= nbands) edges[b++] = lmax + 1;
    return OK;
}

int psb200_host_band_edges(int api, int code, int lmin, int lmax, int lenW, int nbands, int* edges)
{
    if (lmin < 0 || lmax < lmin || nbands < 1 || nbands > 16 || !edges) return fail(ERR_ARG, "host_band_edges: bad arguments");
    HostJob hj{};
    if (api == 0 && code >= 0 && code <= 4) { hj.job = kMcmJob[code]; hj.nout = code == 4 ? 2 : 1; }
    else if (api == 1 && code >= 0 && code <= 6) { hj.job = kCovJob[code]; hj.nout = 1; }
    else if (api == 2) { hj.job = psb::JOB_MASTER; hj.nout = 5; }
    else return fail(ERR_ARG, "host_band_edges: unknown api/code %d/%d", api, code);
    hj.lmin = lmin; hj.lmax = lmax; hj.lenW = lenW;
    host_band_edges(hj, nbands, edges);
    return OK;
}

long long psb200_terms(int families, int lmax, int row_lo, int row_hi)
{
    long long t = 0;
    for (int l = row_lo; l < row_hi; ++l) t += (long long)(2 * l + 1) * (lmax - l + 1);
    return t * families;
}

int psb200_job_stats(int api, int code, int lmax, int lenW, int row_lo, int row_hi, long long* out)
{
    if (!out || lmax < 0 || lenW < 1 || row_lo < 0 || row_hi > lmax + 1 || row_lo > row_hi)
        return fail(ERR_ARG, "job_stats: bad arguments");
    int job = -1;
    if (api == 0 && code >= 0 && code <= 4) job = kMcmJob[code];
    else if (api == 1 && code >= 0 && code <= 6) job = kCovJob[code];
    else if (api == 2) job = psb::JOB_MASTER;
    else return fail(ERR_ARG, "job_stats: unknown api/code %d/%d", api, code);
    int ds = 1, r = 1, nr = 1;
    if (int rc = job_tiling_any(job, &ds, &r, &nr)) return rc;
    int lo = row_lo;
    switch (job) {           // spin-2 jobs: rows l1 < 2 belong to low_rows_kernel
        case psb::JOB_M00: case psb::JOB_TTTT: case psb::JOB_TTTE: case psb::JOB_TTEE: break;
        default: lo = std::min(row_hi, std::max(row_lo, 2));
    }
    long long exec = 0, warps = 0;
    for_each_tile(lmax, lenW, lo, row_hi, ds, r, nr, [&](int, int, long steps) { exec += steps * 32LL * r; ++warps; });
    long long live = 0;
    for (int l1 = lo; l1 < row_hi; ++l1) {
        for (int d = 0; d <= lmax - l1; ++d) {
            const long jend = std::min<long>((long)d + 2L * l1, (long)lenW - 1);
            if (jend >= d) live += (jend - d) / ds + 1;
        }
    }
    out[0] = exec; out[1] = live; out[2] = warps; out[3] = r; out[4] = nr; out[5] = ds;
    return OK;
}

int psb200_mcm_dev(int kind, int lmin, int lmax, const double* dV, int nV, double* dX, long ldX,
                   double* dX2, int row_lo, int row_hi, void* stream)
{
    if (kind < 0 || kind > 4) return fail(ERR_ARG, "unknown mcm kind %d", kind);
    if (int rc = check_common(lmin, lmax, ldX, row_lo, row_hi)) return rc;
    if (!dV || nV < 1 || !dX) return fail(ERR_ARG, "null / empty buffer");
    if (kind == 4 && !dX2) return fail(ERR_ARG, "kind 4 needs a second output");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    psb::PairArgs A{};
    A.lmin = lmin; A.lmax = lmax; A.lenW = nV; A.row_lo = row_lo; A.row_hi = row_hi; A.ld = ldX;
    A.W[0] = dV; A.out0 = dX; A.out1 = dX2;
    return launch_any(kMcmJob[kind], A, (cudaStream_t)stream);
}

int psb200_mcm_master_dev(int lmin, int lmax, const double* dV_TT, const double* dV_TP, const double* dV_PT,
                          const double* dV_PP, int nV, double* const* dX, long ldX, int row_lo, int row_hi,
                          void* stream)
{
    if (int rc = check_common(lmin, lmax, ldX, row_lo, row_hi)) return rc;
    if (!dV_TT || !dV_TP || !dV_PT || !dV_PP || nV < 1 || !dX) return fail(ERR_ARG, "null / empty buffer");
    for (int o = 0; o < 5; ++o) if (!dX[o]) return fail(ERR_ARG, "output %d is null", o);
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    psb::PairArgs A{};
    A.lmin = lmin; A.lmax = lmax; A.lenW = nV; A.row_lo = row_lo; A.row_hi = row_hi; A.ld = ldX;
    A.W[0] = dV_TT; A.W[1] = dV_TP; A.W[2] = dV_PT; A.W[3] = dV_PP;
    A.out0 = dX[0]; A.out1 = dX[1]; A.out2 = dX[2]; A.out3 = dX[3]; A.out4 = dX[4];
    return launch_any(psb::JOB_MASTER, A, (cudaStream_t)stream);
}

int psb200_cov_dev(int block, int lmin, int lmax, const double* const* dsp, int nspec,
                   const double* const* drt, int nratio, const double* const* dW, int nW, int lenW,
                   double* dX, long ldX, int row_lo, int row_hi, void* stream)
{
    if (block < 0 || block > 6) return fail(ERR_ARG, "unknown covariance block %d", block);
    if (int rc = check_common(lmin, lmax, ldX, row_lo, row_hi)) return rc;
    if (nspec != kCovNeedSp[block] || nratio != kCovNeedRt[block] || nW != kCovNeedW[block])
        return fail(ERR_ARG, "block %d takes %d spectra, %d ratios, %d W (got %d, %d, %d)", block,
                    kCovNeedSp[block], kCovNeedRt[block], kCovNeedW[block], nspec, nratio, nW);
    if (!dX || lenW < 1 || !dsp || !dW || (nratio && !drt)) return fail(ERR_ARG, "null / empty buffer");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    psb::PairArgs A{};
    A.lmin = lmin; A.lmax = lmax; A.lenW = lenW; A.row_lo = row_lo; A.row_hi = row_hi; A.ld = ldX;
    for (int k = 0; k < nW; ++k) { if (!dW[k]) return fail(ERR_ARG, "W[%d] is null", k); A.W[k] = dW[k]; }
    for (int k = 0; k < nspec; ++k) { if (!dsp[k]) return fail(ERR_ARG, "spectra[%d] is null", k); A.sp[k] = dsp[k]; }
    for (int k = 0; k < nratio; ++k) { if (!drt[k]) return fail(ERR_ARG, "ratios[%d] is null", k); A.rt[k] = drt[k]; }
    A.out0 = dX;
    return launch_any(kCovJob[block], A, (cudaStream_t)stream);
}

// Device-level calls over SEVERAL row bands in one launch (bands[2k], bands[2k+1] = [lo, hi) of band k, nbands <= 4,
// disjoint): what a rank of the folded multi-GPU split passes -- a low and a high band, so that every rank owns short
// and long tiles.
static int fill_bands(psb::PairArgs& A, int lmin, int lmax, long ld, const int* bands, int nbands)
{
    if (!bands || nbands < 1 || nbands > 4) return fail(ERR_ARG, "need 1..4 row bands");
    for (int k = 0; k < nbands; ++k)
        if (int rc = check_common(lmin, lmax, ld, bands[2 * k], bands[2 * k + 1])) return rc;
    A.row_lo = bands[0]; A.row_hi = bands[1];
    A.nxb = nbands - 1;
    for (int k = 2; k < 2 * nbands; ++k) A.xb[k - 2] = bands[k];
    return OK;
}

int psb200_mcm_dev_bands(int kind, int lmin, int lmax, const double* dV, int nV, double* dX, long ldX, double* dX2,
                         const int* bands, int nbands, void* stream)
{
    if (kind < 0 || kind > 4) return fail(ERR_ARG, "unknown mcm kind %d", kind);
    if (!dV || nV < 1 || !dX) return fail(ERR_ARG, "null / empty buffer");
    if (kind == 4 && !dX2) return fail(ERR_ARG, "kind 4 needs a second output");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    psb::PairArgs A{};
    A.lmin = lmin; A.lmax = lmax; A.lenW = nV; A.ld = ldX;
    if (int rc = fill_bands(A, lmin, lmax, ldX, bands, nbands)) return rc;
    A.W[0] = dV; A.out0 = dX; A.out1 = dX2;
    return launch_any(kMcmJob[kind], A, (cudaStream_t)stream);
}

int psb200_cov_dev_bands(int block, int lmin, int lmax, const double* const* dsp, int nspec,
                         const double* const* drt, int nratio, const double* const* dW, int nW, int lenW,
                         double* dX, long ldX, const int* bands, int nbands, void* stream)
{
    if (block < 0 || block > 6) return fail(ERR_ARG, "unknown covariance block %d", block);
    if (nspec != kCovNeedSp[block] || nratio != kCovNeedRt[block] || nW != kCovNeedW[block])
        return fail(ERR_ARG, "block %d takes %d spectra, %d ratios, %d W (got %d, %d, %d)", block,
                    kCovNeedSp[block], kCovNeedRt[block], kCovNeedW[block], nspec, nratio, nW);
    if (!dX || lenW < 1 || !dsp || !dW || (nratio && !drt)) return fail(ERR_ARG, "null / empty buffer");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    psb::PairArgs A{};
    A.lmin = lmin; A.lmax = lmax; A.lenW = lenW; A.ld = ldX;
    if (int rc = fill_bands(A, lmin, lmax, ldX, bands, nbands)) return rc;
    for (int k = 0; k < nW; ++k) { if (!dW[k]) return fail(ERR_ARG, "W[%d] is null", k); A.W[k] = dW[k]; }
    for (int k = 0; k < nspec; ++k) { if (!dsp[k]) return fail(ERR_ARG, "spectra[%d] is null", k); A.sp[k] = dsp[k]; }
    for (int k = 0; k < nratio; ++k) { if (!drt[k]) return fail(ERR_ARG, "ratios[%d] is null", k); A.rt[k] = drt[k]; }
    A.out0 = dX;
    return launch_any(kCovJob[block], A, (cudaStream_t)stream);
}

int psb200_finish_dev(double* dX, long ldX, int lmin, int lmax, int scale, void* stream)
{
    if (int rc = check_common(lmin, lmax, ldX, lmin, lmax + 1)) return rc;
    if (!dX) return fail(ERR_ARG, "null buffer");
    const int N = lmax - lmin + 1;
    const int nt = (N + 31) / 32;
    finish_kernel<<<dim3(nt, nt), 256, 0, (cudaStream_t)stream>>>(dX, ldX, lmin, N, scale ? 1 : 0, 0);
    CUDA_TRY(cudaGetLastError());
    return OK;
}

int psb200_mcm(int kind, int lmin, int lmax, const double* V, int nV, double* M, long ldM,
               double* M2, int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (kind < 0 || kind > 4) return fail(ERR_ARG, "unknown mcm kind %d", kind);
    if (int rc = check_common(lmin, lmax, ldM, lmin, lmax + 1)) return rc;
    if (!V || nV < 1 || !M) return fail(ERR_ARG, "null / empty buffer");
    if (kind == 4 && !M2) return fail(ERR_ARG, "kind 4 needs a second output");
    int ng = 0;
    if (int rc = resolve_ngpus(ngpus, &ng)) return rc;
    HostJob hj{};
    hj.job = kMcmJob[kind];
    hj.lmin = lmin; hj.lmax = lmax; hj.lenW = nV;
    hj.nW = 1; hj.nsp = 0; hj.nrt = 0;
    hj.vecs[0] = V; hj.lens[0] = (size_t)nV;
    hj.out[0] = M; hj.out[1] = M2; hj.ldo = ldM; hj.nout = kind == 4 ? 2 : 1;
    hj.scale = 1;
    return run_host_job(hj, ng);
}

int psb200_mcm_master(int lmin, int lmax, const double* V_TT, const double* V_TP, const double* V_PT,
                      const double* V_PP, int nV, double* M00, double* M02_TP, double* M02_PT, double* Mpp,
                      double* Mmm, long ldM, int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (int rc = check_common(lmin, lmax, ldM, lmin, lmax + 1)) return rc;
    if (!V_TT || !V_TP || !V_PT || !V_PP || nV < 1 || !M00 || !M02_TP || !M02_PT || !Mpp || !Mmm)
        return fail(ERR_ARG, "null / empty buffer");
    int ng = 0;
    if (int rc = resolve_ngpus(ngpus, &ng)) return rc;
    HostJob hj{};
    hj.job = psb::JOB_MASTER;
    hj.lmin = lmin; hj.lmax = lmax; hj.lenW = nV;
    hj.nW = 4; hj.nsp = 0; hj.nrt = 0;
    const double* Vs[4] = {V_TT, V_TP, V_PT, V_PP};
    for (int k = 0; k < 4; ++k) { hj.vecs[k] = Vs[k]; hj.lens[k] = (size_t)nV; }
    hj.out[0] = M00; hj.out[1] = M02_TP; hj.out[2] = M02_PT; hj.out[3] = Mpp; hj.out[4] = Mmm;
    hj.ldo = ldM; hj.nout = 5; hj.scale = 1;
    return run_host_job(hj, ng);
}

int psb200_cov(int block, int lmin, int lmax, const double* const* spectra, int nspec,
               const double* const* ratios, int nratio, const double* const* W, int nW, int lenW,
               double* C, long ldC, int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (block < 0 || block > 6) return fail(ERR_ARG, "unknown covariance block %d", block);
    if (int rc = check_common(lmin, lmax, ldC, lmin, lmax + 1)) return rc;
    if (nspec != kCovNeedSp[block] || nratio != kCovNeedRt[block] || nW != kCovNeedW[block])
        return fail(ERR_ARG, "block %d takes %d spectra, %d ratios, %d W (got %d, %d, %d)", block,
                    kCovNeedSp[block], kCovNeedRt[block], kCovNeedW[block], nspec, nratio, nW);
    if (!C || lenW < 1 || !spectra || !W || (nratio && !ratios)) return fail(ERR_ARG, "null / empty buffer");
    int ng = 0;
    if (int rc = resolve_ngpus(ngpus, &ng)) return rc;
    HostJob hj{};
    hj.job = kCovJob[block];
    hj.lmin = lmin; hj.lmax = lmax; hj.lenW = lenW;
    hj.nW = nW; hj.nsp = nspec; hj.nrt = nratio;
    int k = 0;
    for (int i = 0; i < nW; ++i, ++k) { if (!W[i]) return fail(ERR_ARG, "W[%d] is null", i); hj.vecs[k] = W[i]; hj.lens[k] = (size_t)lenW; }
    for (int i = 0; i < nspec; ++i, ++k) { if (!spectra[i]) return fail(ERR_ARG, "spectra[%d] is null", i); hj.vecs[k] = spectra[i]; hj.lens[k] = (size_t)lmax + 1; }
    for (int i = 0; i < nratio; ++i, ++k) { if (!ratios[i]) return fail(ERR_ARG, "ratios[%d] is null", i); hj.vecs[k] = ratios[i]; hj.lens[k] = (size_t)lmax + 1; }
    hj.out[0] = C; hj.out[1] = nullptr; hj.ldo = ldC; hj.nout = 1;
    hj.scale = 0;
    return run_host_job(hj, ng);
}

int psb200_quickpol_edges(int lmax, int band_lo, int band_hi, int nbands, int* edges)
{
    if (lmax < 0 || band_lo < 0 || band_hi < 0 || nbands < 1 || !edges) return fail(ERR_ARG, "quickpol_edges: bad arguments");
    auto cost = [&](int l) -> long double {           // sum over the band rows of column l of min(l, l'')
        if (l < 2) return 0;
        const long lo = std::max(2, l - band_hi), hi = std::min(lmax, l + band_lo);
        const long below = l - lo;                     // rows with l'' < l: min = l''
        return (long double)below * (lo + l - 1) / 2 + (long double)(hi - l + 1) * l + 64.0L * (hi - lo + 1);
    };
    long double total = 0;
    for (int l = 0; l <= lmax; ++l) total += cost(l);
    edges[0] = 0;
    long double run = 0;
    int b = 1;
    for (int l = 0; l <= lmax && b < nbands; ++l) {
        run += cost(l);
        while (b < nbands && run >= total * b / nbands) edges[b++] = l + 1;
    }
    while (b <= nbands) edges[b++] = lmax + 1;
    return OK;
}

int psb200_quickpol_xi_dev(int nu1, int nu2, int s1, int s2, int lmax, const double* dW, int lenW,
                           int band_lo, int band_hi, double* dXb, long ldb, int col_lo, int col_hi, void* stream)
{
    if (int rc = check_quickpol(lmax, lenW, band_lo, band_hi, ldb, col_lo, col_hi)) return rc;
    if (!dW || !dXb) return fail(ERR_ARG, "null buffer");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    psb::QpArgs A{};
    A.nu1 = nu1; A.nu2 = nu2; A.s1 = s1; A.s2 = s2; A.lmax = lmax; A.lenW = lenW;
    A.band_lo = band_lo; A.band_hi = band_hi; A.col_lo = col_lo; A.col_hi = col_hi; A.ldb = ldb;
    A.W = dW; A.Xb = dXb;
    return launch_quickpol(A, (cudaStream_t)stream);
}

int psb200_quickpol_xi(int nu1, int nu2, int s1, int s2, int lmax, const double* W, int lenW,
                       int band_lo, int band_hi, double* Xb, long ldb, int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (int rc = check_quickpol(lmax, lenW, band_lo, band_hi, ldb, 0, lmax + 1)) return rc;
    if (!W || !Xb) return fail(ERR_ARG, "null buffer");
    int ng = 0;
    if (int rc = resolve_ngpus(ngpus, &ng)) return rc;
    QpHostJob hj{};
    hj.A.nu1 = nu1; hj.A.nu2 = nu2; hj.A.s1 = s1; hj.A.s2 = s2; hj.A.lmax = lmax; hj.A.lenW = lenW;
    hj.A.band_lo = band_lo; hj.A.band_hi = band_hi; hj.A.ldb = ldb;
    hj.W = W; hj.Xb = Xb;
    int cur = 0;
    cudaGetDevice(&cur);
    std::vector<int> edges(ng + 1);
    psb200_quickpol_edges(lmax, band_lo, band_hi, ng, edges.data());
    std::vector<int> rcs(ng, OK);
    std::vector<std::string> errs(ng);
    std::vector<std::thread> th;
    for (int g = 1; g < ng; ++g)
        th.emplace_back([&, g] { rcs[g] = run_quickpol_on_device(hj, g, edges[g], edges[g + 1], &errs[g]); });
    rcs[0] = run_quickpol_on_device(hj, 0, edges[0], edges[1], &errs[0]);
    for (auto& t : th) t.join();
    cudaSetDevice(cur);
    for (int g = 0; g < ng; ++g)
        if (rcs[g] != OK) { g_err = "device " + std::to_string(g) + ": " + errs[g]; return rcs[g]; }
    return OK;
}


// ---- decoupling on the device (psb200_solve.inl) ------------------------------------------------
int psb200_mcm_solve(int system, int lmin, int lmax, const double* V, int nV, const double* pCl, long ldp, int nrhs,
                     double* Cl, long ldc, int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (system < 0 || system > 5) return fail(ERR_ARG, "unknown system %d", system);
    const int N = lmax - lmin + 1;
    if (int rc = check_common(lmin, lmax, N, lmin, lmax + 1)) return rc;
    const int nb = system >= 4 ? 2 : 1;
    if (!V || nV < 1 || !pCl || !Cl || nrhs < 1) return fail(ERR_ARG, "null / empty buffer");
    if (ldp < (long)nb * N || ldc < (long)nb * N) return fail(ERR_ARG, "leading dimension of pCl / Cl below %d", nb * N);
    int ng = 0;
    if (int rc = resolve_ngpus(ngpus, &ng)) return rc;
    HostJob hj{};
    hj.job = system >= 4 ? (int)psb::JOB_MPPMMM : kMcmJob[system];
    hj.lmin = lmin; hj.lmax = lmax; hj.lenW = nV;
    hj.nW = 1; hj.vecs[0] = V; hj.lens[0] = (size_t)nV;
    hj.nout = nb; hj.scale = 1;
    int cur = 0, root = 0;
    cudaGetDevice(&cur);
    int rc = full_on_root(hj, ng, &root);
    if (rc == OK) {
        std::vector<RhsCol> cols(nrhs);
        for (int k = 0; k < nrhs; ++k) {
            cols[k].in[0] = pCl + (size_t)k * ldp; cols[k].in[1] = cols[k].in[0] + N;
            cols[k].out[0] = Cl + (size_t)k * ldc; cols[k].out[1] = cols[k].out[0] + N;
        }
        DeviceScratch& R = g_scratch[root];
        if (nb == 1) rc = solve_on_root(root, R.X[0], N, 1, cols);
        else if (system == 4) rc = solve_block_systems_on_root(root, R.X[0], R.X[1], N, cols, {});
        else rc = solve_block_systems_on_root(root, R.X[0], R.X[1], N, {}, cols);
    }
    cudaSetDevice(cur);
    return rc;
}

int psb200_master_solve(int lmin, int lmax, const double* V_TT, const double* V_TP, const double* V_PT,
                        const double* V_PP, int nV, const double* pCl, long ldp, double* Cl, long ldc, int ngpus)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    const int N = lmax - lmin + 1;
    if (int rc = check_common(lmin, lmax, N, lmin, lmax + 1)) return rc;
    if (!V_TT || !V_TP || !V_PT || !V_PP || nV < 1 || !pCl || !Cl) return fail(ERR_ARG, "null / empty buffer");
    if (ldp < N || ldc < N) return fail(ERR_ARG, "leading dimension of pCl / Cl below N=%d", N);
    int ng = 0;
    if (int rc = resolve_ngpus(ngpus, &ng)) return rc;
    HostJob hj{};
    hj.job = psb::JOB_MASTER;
    hj.lmin = lmin; hj.lmax = lmax; hj.lenW = nV;
    hj.nW = 4;
    const double* Vs[4] = {V_TT, V_TP, V_PT, V_PP};
    for (int k = 0; k < 4; ++k) { hj.vecs[k] = Vs[k]; hj.lens[k] = (size_t)nV; }
    hj.nout = 5; hj.scale = 1;
    int cur = 0, root = 0;
    cudaGetDevice(&cur);
    int rc = full_on_root(hj, ng, &root);
    // columns: 0 TT, 1 TE, 2 ET, 3 TB, 4 BT, 5 EE, 6 BB, 7 EB, 8 BE   (src/modecoupling.jl:348-377)
    auto col = [&](int a) { RhsCol c{}; c.in[0] = pCl + (size_t)a * ldp; c.out[0] = Cl + (size_t)a * ldc; return c; };
    auto col2 = [&](int a, int b) { RhsCol c = col(a); c.in[1] = pCl + (size_t)b * ldp; c.out[1] = Cl + (size_t)b * ldc; return c; };
    DeviceScratch& R = g_scratch[root];
    // both block systems from one LU of M++ + M-- and one of M++ - M--
    if (rc == OK) rc = solve_block_systems_on_root(root, R.X[3], R.X[4], N, {col2(5, 6)}, {col2(7, 8)});
    if (rc == OK) rc = solve_on_root(root, R.X[0], N, 1, {col(0)});
    if (rc == OK) rc = solve_on_root(root, R.X[1], N, 1, {col(1), col(3)});     // TE and TB share mcm(:TE, maskT1, maskP2)
    if (rc == OK) rc = solve_on_root(root, R.X[2], N, 1, {col(2), col(4)});     // ET and BT share mcm(:ET, maskP1, maskT2)
    cudaSetDevice(cur);
    return rc;
}

int psb200_decouple_covmat_dev(int n, double* dY, long ldy, const double* dB1, long ld1, const double* dB2, long ld2,
                               void* stream)
{
    if (n < 1 || !dY || !dB1 || !dB2 || ldy < n || ld1 < n || ld2 < n) return fail(ERR_ARG, "decouple_covmat: bad arguments");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", dev);
    return decouple_on_device(dev, (cudaStream_t)stream, n, dY, ldy, dB1, ld1, dB2, ld2);
}

int psb200_decouple_covmat(int n, const double* Y, long ldy, const double* B1, long ld1, const double* B2, long ld2,
                           double* out, long ldo)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (n < 1 || !Y || !B1 || !B2 || !out || ldy < n || ld1 < n || ld2 < n || ldo < n)
        return fail(ERR_ARG, "decouple_covmat: bad arguments");
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", dev);
    for (int o = 0; o < 3; ++o)
        if (int rc = scratch_reserve(dev, o, (size_t)n * n)) return rc;
    DeviceScratch& s = g_scratch[dev];
    const size_t w = (size_t)n * sizeof(double);
    CUDA_TRY(cudaMemcpy2DAsync(s.X[0], w, Y, (size_t)ldy * sizeof(double), w, n, cudaMemcpyHostToDevice, s.stream));
    CUDA_TRY(cudaMemcpy2DAsync(s.X[1], w, B1, (size_t)ld1 * sizeof(double), w, n, cudaMemcpyHostToDevice, s.stream));
    CUDA_TRY(cudaMemcpy2DAsync(s.X[2], w, B2, (size_t)ld2 * sizeof(double), w, n, cudaMemcpyHostToDevice, s.stream));
    if (int rc = decouple_on_device(dev, s.stream, n, s.X[0], n, s.X[1], n, s.X[2], n)) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)ldo * sizeof(double), s.X[0], w, w, n, cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return OK;
}


// ---- W-spectrum production, first slice: zonal map2alm (psb200_zonal.cuh) -----------------------
int psb200_zonal_alm(int nfields, int nnodes, const double* x, const double* w, const double* fields, long ldf,
                     int lmax, double* alm, long lda)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    if (nfields < 1 || nnodes < 1 || lmax < 0 || !x || !w || !fields || !alm || ldf < nnodes || lda < (long)lmax + 1)
        return fail(ERR_ARG, "zonal_alm: bad arguments");
    if (lmax > 65535) return fail(ERR_ARG, "zonal_alm: lmax %d above the supported 65535", lmax);
    if (device_count() <= 0) return fail(ERR_NODEVICE, "no CUDA device visible: libpsb200 has no CPU fallback");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 16) return fail(ERR_ARG, "device index %d above the supported 15", dev);
    const int nblocks = (nnodes + psb::ZN_THREADS - 1) / psb::ZN_THREADS;
    const size_t L = (size_t)lmax + 1;
    for (int f0 = 0; f0 < nfields; f0 += psb::ZN_FMAX) {
        const int nf = std::min(psb::ZN_FMAX, nfields - f0);
        // scratch: [x | w | fields nf x n | alm nf x L] in slot 0, partial sums in slot 1
        const size_t nin = 2 * (size_t)nnodes + (size_t)nf * nnodes + (size_t)nf * L;
        if (int rc = scratch_reserve(dev, 0, nin)) return rc;
        if (int rc = scratch_reserve(dev, 1, (size_t)nblocks * nf * L)) return rc;
        DeviceScratch& s = g_scratch[dev];
        double *dx = s.X[0], *dw = dx + nnodes, *df = dw + nnodes, *da = df + (size_t)nf * nnodes;
        CUDA_TRY(cudaMemcpyAsync(dx, x, nnodes * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(dw, w, nnodes * sizeof(double), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpy2DAsync(df, nnodes * sizeof(double), fields + (size_t)f0 * ldf, ldf * sizeof(double),
                                   nnodes * sizeof(double), nf, cudaMemcpyHostToDevice, s.stream));
        psb::zonal_partial_kernel<<<nblocks, psb::ZN_THREADS, 0, s.stream>>>(dx, dw, df, nnodes, nf, nnodes, lmax, s.X[1]);
        CUDA_TRY(cudaGetLastError());
        psb::zonal_finish_kernel<<<dim3((unsigned)((L + 127) / 128), nf), 128, 0, s.stream>>>(s.X[1], nblocks, nf, lmax, da, (long)L);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy2DAsync(alm + (size_t)f0 * lda, lda * sizeof(double), da, L * sizeof(double), L * sizeof(double), nf,
                                   cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaStreamSynchronize(s.stream));
    }
    return OK;
}

#include "psb200_sht_abi.inl"

/* Test hook (host memory only, no device): the delivery of one band [a, b) of an N x N result cut into nsub sub-bands,
 * executed on stand-in "device" slabs.  The slabs start with distinct RAW values x(l1, l2), l2 >= l1; the hook then does on
 * the host what the device kernels do per sub-band (finish_kernel on the diagonal block; band_transpose_kernel for the
 * columns to its right unless mode = 2) and delivers: mode 0 directly (the 2-D copies as the page-locked path issues
 * them), mode 1 through the staged pipeline (ring of nch chunks of chunk_kb KB, nthreads scatter workers), mode 2 by
 * mirror delivery (block columns only, the block rows written by the scatter workers), mode 3 by mirror delivery with the
 * block columns copied straight to their place (scale 0 only: what a page-locked array gets for a covariance block).  All
 * must leave the same bytes in out[0..nout-1]; tests/test_host.py compares them.  scale: 1 = MCM factors (2l+1), 0 = symmetric copy. */
int psb200_selftest_delivery(int lmin, int lmax, int a, int b, int nsub, int nout, int mode, int scale, int chunk_kb, int nch,
                             int nthreads, double* const* out, long ldo)
{
    const int N = lmax - lmin + 1, nb = b - a;
    if (lmin < 0 || a < lmin || b > lmax + 1 || nb <= 0 || nout < 1 || nout > 5 || !out || ldo < N || chunk_kb < 1 || nch < 1 ||
        nch > 32 || nthreads < 1 || mode < 0 || mode > 3 || scale < 0 || scale > 1 || (mode == 3 && scale != 0))
        return fail(ERR_ARG, "selftest_delivery: bad arguments");
    const std::vector<int> sub = split_rows(a, b, lmax, lmax + 1, nsub);
    const int ns = (int)sub.size() - 1;
    const bool mirror = mode >= 2;
    std::vector<size_t> toff(ns + 1, 0);
    for (int k = 0; k < ns; ++k)
        toff[k + 1] = toff[k] + (size_t)(sub[k + 1] - sub[k]) * (size_t)(N - (sub[k + 1] - lmin)) * nout;
    std::vector<std::vector<double>> Xs(nout, std::vector<double>((size_t)nb * N));
    std::vector<double> T(toff[ns] + 1, 0.0);
    double* X[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const long ld = N;
    auto fac = [&](long l) { return scale ? (double)(2 * l + 1) : 1.0; };
    for (int o = 0; o < nout; ++o) {
        X[o] = Xs[o].data();
        // raw values: row l1 - a, column l2 - lmin >= l1 - lmin (entries left of the diagonal are never read: poison them)
        for (int i = 0; i < nb; ++i)
            for (int j = 0; j < N; ++j)
                X[o][(size_t)i * ld + j] = j >= (a - lmin) + i ? (double)(o + 1) + 1e-3 * (a + i) + 1e-7 * (lmin + j) : -7.77e77;
        for (int k = 0; k < ns; ++k) {
            const int sa = sub[k], sb = sub[k + 1], nbs = sb - sa, c0 = sb - lmin;
            double* slab = X[o] + (size_t)(sa - a) * ld;
            double* D = slab + (sa - lmin);                       // diagonal block, as finish_kernel sees it
            for (int i = 0; i < nbs; ++i)
                for (int j = i; j < nbs; ++j) {
                    const double x = D[(size_t)i * ld + j];
                    D[(size_t)i * ld + j] = fac(sa + i) * x;
                    if (j > i) D[(size_t)j * ld + i] = fac(sa + j) * x;
                }
            if (c0 < N && !mirror) {                              // band_transpose_kernel
                double* Tk = T.data() + toff[k] + (size_t)o * nbs * (N - c0);
                for (int i = 0; i < nbs; ++i)
                    for (int j = c0; j < N; ++j) {
                        const double x = slab[(size_t)i * ld + j];
                        slab[(size_t)i * ld + j] = fac(sa + i) * x;
                        Tk[(size_t)(j - c0) * nbs + i] = fac(lmin + j) * x;
                    }
            }
        }
    }
    const std::vector<Copy2D> copies = band_copies(N, lmin, a, sub, nout, out, ldo, X, ld, T.data(), toff, mirror, scale);
    auto copy2d = [](char* dst, size_t dpitch, const Copy2D& c) {
        for (size_t r = 0; r < c.h; ++r) memcpy(dst + r * dpitch, (const char*)c.src + r * c.spitch, c.width);
    };
    if (mode == 0) {
        for (const Copy2D& c : copies) copy2d((char*)c.dst, c.dpitch, c);
        return OK;
    }
    const size_t chunk = (size_t)chunk_kb << 10;
    for (const Copy2D& c : copies)
        if (c.width > chunk) return fail(ERR_ARG, "selftest_delivery: a row of %zu bytes does not fit a chunk", c.width);
    std::vector<Copy2D> pieces = split_copies(copies, chunk);
    if (mode == 3)
        for (Copy2D& p : pieces) p.direct = 1;
    std::vector<char> ring(chunk * (size_t)nch);
    auto issue = [&](int i, char* dst, int) -> int {
        if (pieces[i].direct) copy2d((char*)pieces[i].dst, pieces[i].dpitch, pieces[i]);
        else copy2d(dst, pieces[i].width, pieces[i]);
        return OK;
    };
    auto wait = [](int) -> int { return OK; };
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = deliver_staged(pieces, ring.data(), chunk, nch, nthreads, issue, wait);
    if (trace_level() >= 1)
        fprintf(stderr, "[psb200] selftest delivery mode %d, %d workers: %.3f ms\n", mode, nthreads,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return rc;
}

double psb200_dfma_peak(int iters)
{
    if (device_count() <= 0) { fail(ERR_NODEVICE, "no CUDA device visible"); return -1.0; }
    if (iters < 1) iters = 1;
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { fail(ERR_CUDA, "cudaGetDeviceProperties failed"); return -1.0; }
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) { fail(ERR_OOM, "cudaMalloc failed"); return -1.0; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<<<blocks, threads>>>(out, iters / 4 + 1, 1.0);   // warm-up
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) { fail(ERR_CUDA, "dfma kernel: %s", cudaGetErrorString(e)); return -1.0; }
    const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;
    return flops / (ms * 1e-3);
}

}  // extern "C"

#include "psb200_hostmem.inl"
