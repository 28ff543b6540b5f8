// psb200_zonal.cuh -- W-spectrum production, first slice: the spin-0 map2alm of ZONAL maps.
//
// The reference builds every window spectrum W from spherical-harmonic transforms of mask products
// (effective_weight_alm!, /root/reference/src/workspace.jl:141-171: map2alm of mask_i .* mask_j [.* sigma^2 Omega_pix];
// window_function_W!, :174-213: averages of alm2cl of two such alms).  A general HEALPix map2alm is a different
// roofline (FFT + associated-Legendre synthesis) and stays on the host.  For azimuthally symmetric maps only m = 0
// survives and map2alm collapses to a Legendre quadrature,
//     a_l0 = sqrt(pi (2l+1)) sum_k w_k f(x_k) P_l(x_k),     x_k, w_k: Gauss-Legendre nodes / weights,
// which is what this kernel evaluates for a batch of fields at once (the synthetic skies of the benchmark and the tests
// are zonal; powerspectra.jl_b200/synthetic.py::ZonalSky.al0 is the host statement of the same sum).
// One thread per node runs the three-term recurrence of P_l along l; per l the nf products are reduced over the block
// (warp shuffles + a small shared array); per-block partial sums are added in block order by a second kernel, so the
// result does not depend on scheduling.  HBM: O(n nf) inputs, O(nblocks nf lmax) partials -- nothing of note.
#pragma once
#include <cuda_runtime.h>

namespace psb {

constexpr int ZN_THREADS = 128;     // nodes per block
constexpr int ZN_FMAX = 32;         // fields per launch

__global__ void __launch_bounds__(ZN_THREADS) zonal_partial_kernel(const double* __restrict__ x, const double* __restrict__ w,
                                                                   const double* __restrict__ f, long ldf, int nf, int n,
                                                                   int lmax, double* __restrict__ partial)
{
    __shared__ double red[ZN_THREADS / 32][ZN_FMAX];
    const int k = blockIdx.x * ZN_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double xk = k < n ? x[k] : 0.0;
    double wf[ZN_FMAX];
#pragma unroll
    for (int i = 0; i < ZN_FMAX; ++i) wf[i] = (i < nf && k < n) ? w[k] * f[(long)i * ldf + k] : 0.0;
    double p0 = 1.0, p1 = xk;        // P_{l-1}, P_l while l >= 1
    for (int l = 0; l <= lmax; ++l) {
        double p;
        if (l == 0) p = 1.0;
        else if (l == 1) p = xk;
        else {
            p = ((double)(2 * l - 1) * xk * p1 - (double)(l - 1) * p0) / (double)l;
            p0 = p1; p1 = p;
        }
#pragma unroll
        for (int i = 0; i < ZN_FMAX; ++i) {
            if (i < nf) {                                   // uniform
                double v = wf[i] * p;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[wid][i] = v;
            }
        }
        __syncthreads();
        if (threadIdx.x < nf) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < ZN_THREADS / 32; ++q) s += red[q][threadIdx.x];
            partial[((size_t)blockIdx.x * nf + threadIdx.x) * (size_t)(lmax + 1) + l] = s;
        }
        __syncthreads();
    }
}

// alm[i][l] = sqrt(pi (2l+1)) sum_b partial[b][i][l], blocks in order
__global__ void zonal_finish_kernel(const double* __restrict__ partial, int nblocks, int nf, int lmax, double* __restrict__ alm, long lda)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (l > lmax) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[((size_t)b * nf + i) * (size_t)(lmax + 1) + l];
    alm[(long)i * lda + l] = s * sqrt(3.14159265358979323846264338327950288 * (double)(2 * l + 1));
}

}  // namespace psb
