// TEST INFRASTRUCTURE ONLY.  Compiles the __host__ __device__ pair function of
// powerspectra.jl_b200/csrc/psb200_quickpol.cuh with g++ so that the CPU test suite can check the
// arithmetic of the CUDA kernel against the oracle without a GPU.  Never part of libpsb200.so, never
// imported by the product package.
#include "../../powerspectra.jl_b200/csrc/psb200_quickpol.cuh"

extern "C" int qp_host_xi(int nu1, int nu2, int s1, int s2, int lmax, const double* W, int lenW,
                          int band_lo, int band_hi, double* Xb, long ldb)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int l = 2; l <= lmax; ++l) {
        for (int r = 0; r < band_lo + band_hi + 1; ++r) {
            const int lpp = l + r - band_hi;
            if (lpp < 2 || lpp > lmax) continue;
            Xb[(long)r + (long)l * ldb] = psb::quickpol_pair(l, lpp, nu1, nu2, s1, s2, W, lenW);
        }
    }
    return 0;
}

extern "C" double qp_host_pair(int l, int lpp, int nu1, int nu2, int s1, int s2, const double* W, int lenW)
{
    return psb::quickpol_pair(l, lpp, nu1, nu2, s1, s2, W, lenW);
}
