// TEST INFRASTRUCTURE ONLY.  Compiles the __host__ __device__ pair function of
// powerspectra.jl_b200/csrc/psb200_quickpol.cuh with g++ so that the CPU test suite can check the
// arithmetic of the CUDA kernel against the oracle without a GPU.  Never part of libpsb200.so, never
// imported by the product package.
#include "../../powerspectra.jl_b200/csrc/psb200_quickpol.cuh"

#include <vector>

namespace {
struct HostTabs {
    std::vector<double> ij2;
    std::vector<psb::QpD2> bb0, bb1;
    psb::QpTabs T;
    HostTabs(int jmax, int m1a, int m1b) : ij2(jmax + 2), bb0(jmax + 2), bb1(jmax + 2)
    {
        for (int j = 0; j < jmax + 2; ++j) psb::qp_tab_entry(j, m1a, m1b, &ij2[j], &bb0[j], &bb1[j]);   // as quickpol_tables_kernel
        T.IJ2 = ij2.data(); T.BB0 = bb0.data(); T.BB1 = bb1.data();
    }
};
}  // namespace

// variant: 0 = simple, 1 = tabulated (the two instantiations of the CUDA kernel)
extern "C" int qp_host_xi(int variant, int nu1, int nu2, int s1, int s2, int lmax, const double* W, int lenW,
                          int band_lo, int band_hi, double* Xb, long ldb)
{
    const HostTabs H(2 * lmax, s1 + nu1, s2 + nu2);
#pragma omp parallel for schedule(dynamic, 1)
    for (int l = 2; l <= lmax; ++l) {
        for (int r = 0; r < band_lo + band_hi + 1; ++r) {
            const int lpp = l + r - band_hi;
            if (lpp < 2 || lpp > lmax) continue;
            Xb[(long)r + (long)l * ldb] = variant ? psb::quickpol_pair_t<true>(l, lpp, nu1, nu2, s1, s2, W, lenW, H.T)
                                                  : psb::quickpol_pair_t<false>(l, lpp, nu1, nu2, s1, s2, W, lenW, H.T);
        }
    }
    return 0;
}

extern "C" double qp_host_pair(int variant, int l, int lpp, int nu1, int nu2, int s1, int s2, const double* W, int lenW)
{
    const HostTabs H(l + lpp, s1 + nu1, s2 + nu2);
    return variant ? psb::quickpol_pair_t<true>(l, lpp, nu1, nu2, s1, s2, W, lenW, H.T)
                   : psb::quickpol_pair_t<false>(l, lpp, nu1, nu2, s1, s2, W, lenW, H.T);
}
