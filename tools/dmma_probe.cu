// dmma_probe.cu -- does the FP64 tensor path (mma.sync m8n8k4 f64, "DMMA") of B200 run beside the FP64 vector
// pipe (DFMA), or do they share one datapath?  Decides whether a tensor-assisted Xi accumulation
// (Xi[pair, q] += g[pair, j] W'[j, q] is an M x 8 x K product) could lift the covariance kernels above the
// DFMA roofline they sit at today.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/dmma_probe tools/dmma_probe.cu
// Prints one JSON line.  Measurement tool, not part of the library.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// m16n8k8 (sm_90+ shape): 16 x 8 x 8 = 1024 FMA per warp instruction
__device__ __forceinline__ void dmma16(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// big shape alone (mode 0) or interleaved with DFMA in every warp (mode 1)
template <int MODE>
__global__ void __launch_bounds__(256) probe16(double* out, int iters, double seed)
{
    double f[8], c[4][4];
    const double m = 1.0000001, k = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = seed + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = seed * i + j;
    const double a[4] = {1.0 + 1e-9 * threadIdx.x, 1.0, 1.0 - 1e-9, 1.0 + 2e-9}, b[2] = {1.0 - 1e-9 * threadIdx.x, 1.0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fma(f[i], m, k);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma16(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static float run16(double* out, int blocks, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe16<MODE><<<blocks, 256>>>(out, iters / 4 + 1, 1.0);
    cudaEventRecord(e0);
    probe16<MODE><<<blocks, 256>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

// mode 0: DFMA only; 1: DMMA only; 2: both in every warp (interleaved); 3: even warps DFMA, odd warps DMMA
template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed)
{
    double f[8], c[8][2];
    const double m = 1.0000001, k = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = seed + threadIdx.x + i; c[i][0] = seed * i; c[i][1] = seed + i; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    const bool odd = (threadIdx.x >> 5) & 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0 || MODE == 2 || (MODE == 3 && !odd)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fma(f[i], m, k);
            }
            if (MODE == 1 || MODE == 2 || (MODE == 3 && odd)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i] + c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static float run(double* out, int blocks, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(out, iters / 4 + 1, 1.0);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    const int blocks = p.multiProcessorCount * 8, iters = 4000;
    double* out = nullptr;
    cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double));
    const double threads = (double)blocks * 256, warps = threads / 32;
    const double dfma_flops = 2.0 * 8 * 4 * iters * threads;              // per kernel that runs DFMA in every warp
    const double dmma_flops = 2.0 * 8 * 8 * 4 * 8 * 4 * iters * warps;    // 8 mma x (8x8x4 FMA) x 4 unroll, per warp
    const float t0 = run<0>(out, blocks, iters), t1 = run<1>(out, blocks, iters);
    const float t2 = run<2>(out, blocks, iters), t3 = run<3>(out, blocks, iters);
    const float t4 = run16<0>(out, blocks, iters), t5 = run16<1>(out, blocks, iters);
    const double dmma16_flops = 2.0 * 4 * 16 * 8 * 8 * 4 * iters * warps;  // 4 mma x (16x8x8 FMA) x 4 unroll, per warp
    const cudaError_t e = cudaDeviceSynchronize();
    printf("{\"what\": \"DFMA vs DMMA (mma.sync m8n8k4 f64) on %s\", \"cuda_error\": \"%s\", "
           "\"dfma_only_tflops\": %.2f, \"dmma_only_tflops\": %.2f, "
           "\"interleaved_ms\": %.3f, \"interleaved_dfma_tflops\": %.2f, \"interleaved_dmma_tflops\": %.2f, "
           "\"split_warps_ms\": %.3f, \"split_dfma_tflops\": %.2f, \"split_dmma_tflops\": %.2f, "
           "\"dfma_only_ms\": %.3f, \"dmma_only_ms\": %.3f, "
           "\"dmma_m16n8k8_only_tflops\": %.2f, \"m16n8k8_interleaved_dmma_tflops\": %.2f, \"m16n8k8_interleaved_dfma_tflops\": %.2f}\n",
           p.name, cudaGetErrorString(e), dfma_flops / t0 * 1e-9, dmma_flops / t1 * 1e-9,
           t2, dfma_flops / t2 * 1e-9, dmma_flops / t2 * 1e-9,
           t3, 0.5 * dfma_flops / t3 * 1e-9, 0.5 * dmma_flops / t3 * 1e-9, t0, t1,
           dmma16_flops / t4 * 1e-9, dmma16_flops / t5 * 1e-9, dfma_flops / t5 * 1e-9);
    cudaFree(out);
    return 0;
}
