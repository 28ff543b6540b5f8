// TEST INFRASTRUCTURE ONLY.  Compiles the __host__ __device__ parts of powerspectra.jl_b200/csrc/psb200_sht.cuh with
// g++ (ring geometry, mixed-radix ring FFTs with the real-input packing and the aliasing rule, the scaled lambda_lm
// recurrence with its start values, rescaling cadence and mlim skip) and strings them together with plain loops in the
// order the CUDA kernels use, so the CPU suite can check the arithmetic of the device path against the oracle without
// a GPU.  What only exists on the device (warp butterfly, staging, launch geometry) is not covered here.
// Never part of libpsb200.so, never imported by the product package.
#include "../../powerspectra.jl_b200/csrc/psb200_sht.cuh"

#include <cmath>
#include <vector>

namespace {

struct HostCtx {
    int tid = 0, nthr = 1;
    void sync() {}
};

struct Dims { int nside, lmax, nrp; long long npix, nalm; };

Dims dims(int nside, int lmax)
{
    return Dims{nside, lmax, 2 * nside, 12LL * nside * nside, (long long)(lmax + 1) * (lmax + 2) / 2};
}

std::vector<double> cm_table(int lmax)
{
    std::vector<double> cm(lmax + 1);
    long double acc = 0.0L;
    for (int m = 0; m <= lmax; ++m) {
        if (m) acc += log2l((2.0L * m - 1.0L) / (2.0L * m));
        cm[m] = (double)(0.5L * (log2l((2.0L * m + 1.0L) / (4.0L * 3.14159265358979323846264338327950288L)) + acc));
    }
    return cm;
}

void ring_stage_analysis(const Dims& D, const double* map, std::vector<double2>& Phi)
{
    Phi.assign((size_t)(D.lmax + 1) * D.nrp * 2, make_double2(0.0, 0.0));
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < D.nrp; ++p) {
        const psb::ShtRing g = psb::sht_ring(D.nside, p);
        const int h = g.n / 2;
        std::vector<double2> A(h), B(h), T(h);
        int rad[16];
        const int nrad = psb::sht_factor(h, rad);
        for (int hemi = 0; hemi < 2; ++hemi) {
            if (hemi == 1 && p == D.nrp - 1) continue;
            HostCtx cx;
            psb::sht_ring_analyse(cx, map + (hemi ? g.startS : g.startN), g.n, g.shifted,
                                  12.566370614359172953850573533118 / (double)D.npix, D.lmax, A.data(), B.data(), T.data(), rad,
                                  nrad, Phi.data() + ((size_t)p * 2 + hemi), (long long)D.nrp * 2);
        }
    }
}

void ring_stage_synthesis(const Dims& D, const std::vector<double2>& Phi, const double* ref, double* map)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < D.nrp; ++p) {
        const psb::ShtRing g = psb::sht_ring(D.nside, p);
        const int h = g.n / 2;
        std::vector<double2> A(h), B(h + 1), T(h);
        int rad[16];
        const int nrad = psb::sht_factor(h, rad);
        for (int hemi = 0; hemi < 2; ++hemi) {
            if (hemi == 1 && p == D.nrp - 1) continue;
            HostCtx cx;
            const long long st = hemi ? g.startS : g.startN;
            psb::sht_ring_synthesise(cx, Phi.data() + ((size_t)p * 2 + hemi), (long long)D.nrp * 2, g.n, g.shifted, D.lmax,
                                     A.data(), B.data(), T.data(), T.data(), rad, nrad, ref ? ref + st : nullptr, map + st);
        }
    }
}

// the lane code of sht_leg_analysis_kernel for one (m, ring pair), rings summed in index order
void leg_analysis(const Dims& D, const std::vector<double2>& Phi, const std::vector<double>& cm, int accumulate, double* alm)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m <= D.lmax; ++m) {
        const long long base = psb::sht_alm_base(D.lmax, m);
        std::vector<double> sre(D.lmax + 1, 0.0), sim(D.lmax + 1, 0.0);
        for (int p = 0; p < D.nrp; ++p) {
            const psb::ShtRing g = psb::sht_ring(D.nside, p);
            psb::ShtLam q = psb::sht_lam_start(D.lmax, m, cm[m], g.z, g.s, true);
            if (q.e == psb::SHT_NEVER) continue;
            double ger = 0, gei = 0, gor = 0, goi = 0;
            auto load = [&]() {
                const double2 n = Phi[((size_t)m * D.nrp + p) * 2], s = Phi[((size_t)m * D.nrp + p) * 2 + 1];
                ger = n.x + s.x; gei = n.y + s.y; gor = n.x - s.x; goi = n.y - s.y;
            };
            if (q.e == 0) load();
            for (int l0 = m; l0 <= D.lmax; l0 += psb::SHT_C) {
                double2 cf[psb::SHT_C];
                psb::sht_coef_pass(l0, m, cf);
                for (int j = 0; j < psb::SHT_C; ++j) {
                    const int l = l0 + j;
                    const double Qj = j ? cf[j - 1].y : 1.0;
                    if (l <= D.lmax) {
                        sre[l] += Qj * (q.lc * ((j & 1) ? gor : ger));
                        sim[l] += Qj * (q.lc * ((j & 1) ? goi : gei));
                    }
                    psb::sht_mu_advance(q, cf[j].x);
                }
                psb::sht_mu_close(q, cf[psb::SHT_C - 2].y, cf[psb::SHT_C - 1].y);
                if (psb::sht_lam_rescale(q)) load();
            }
        }
        for (int l = m; l <= D.lmax; ++l) {
            double* a = alm + 2 * (base + l);
            a[0] = accumulate ? a[0] + sre[l] : sre[l];
            a[1] = accumulate ? a[1] + sim[l] : sim[l];
        }
    }
}

void leg_synthesis(const Dims& D, const double* alm, const std::vector<double>& cm, std::vector<double2>& Phi)
{
    Phi.assign((size_t)(D.lmax + 1) * D.nrp * 2, make_double2(0.0, 0.0));
#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m <= D.lmax; ++m) {
        const long long base = psb::sht_alm_base(D.lmax, m);
        for (int p = 0; p < D.nrp; ++p) {
            const psb::ShtRing g = psb::sht_ring(D.nside, p);
            psb::ShtLam q = psb::sht_lam_start(D.lmax, m, cm[m], g.z, g.s, true);
            if (q.e == psb::SHT_NEVER) continue;
            double fer = 0, fei = 0, forr = 0, foi = 0;
            for (int l0 = m; l0 <= D.lmax; l0 += psb::SHT_C) {
                double2 cf[psb::SHT_C];
                psb::sht_coef_pass(l0, m, cf);
                for (int j = 0; j < psb::SHT_C; ++j) {
                    const int l = l0 + j;
                    const double Qj = j ? cf[j - 1].y : 1.0;
                    const double ar = l <= D.lmax ? Qj * alm[2 * (base + l)] : 0.0, ai = l <= D.lmax ? Qj * alm[2 * (base + l) + 1] : 0.0;
                    if (j & 1) { forr = fma(q.lc, ar, forr); foi = fma(q.lc, ai, foi); }
                    else { fer = fma(q.lc, ar, fer); fei = fma(q.lc, ai, fei); }
                    psb::sht_mu_advance(q, cf[j].x);
                }
                psb::sht_mu_close(q, cf[psb::SHT_C - 2].y, cf[psb::SHT_C - 1].y);
                if (psb::sht_lam_rescale(q)) fer = fei = forr = foi = 0.0;
            }
            if (q.e == 0) {
                Phi[((size_t)m * D.nrp + p) * 2] = make_double2(fer + forr, fei + foi);
                Phi[((size_t)m * D.nrp + p) * 2 + 1] = make_double2(fer - forr, fei - foi);
            }
        }
    }
}

}  // namespace

extern "C" int sht_host_map2alm(int nside, int lmax, int niter, const double* map, double* alm)
{
    const Dims D = dims(nside, lmax);
    const std::vector<double> cm = cm_table(lmax);
    std::vector<double2> Phi;
    std::vector<double> resid(D.npix);
    ring_stage_analysis(D, map, Phi);
    leg_analysis(D, Phi, cm, 0, alm);
    for (int it = 0; it < niter; ++it) {
        leg_synthesis(D, alm, cm, Phi);
        ring_stage_synthesis(D, Phi, map, resid.data());
        ring_stage_analysis(D, resid.data(), Phi);
        leg_analysis(D, Phi, cm, 1, alm);
    }
    return 0;
}

extern "C" int sht_host_alm2map(int nside, int lmax, const double* alm, double* map)
{
    const Dims D = dims(nside, lmax);
    const std::vector<double> cm = cm_table(lmax);
    std::vector<double2> Phi;
    leg_synthesis(D, alm, cm, Phi);
    ring_stage_synthesis(D, Phi, nullptr, map);
    return 0;
}

// lambda_lm(theta of ring pair p), l = m..lmax, through the scaled recurrence; entries before the ring is representable are 0.
// returns the l at which the ring came alive (lmax+1: never; -1: skipped by mlim)
extern "C" int sht_host_lambda(int nside, int lmax, int m, int p, double* lam)
{
    const std::vector<double> cm = cm_table(lmax);
    const psb::ShtRing g = psb::sht_ring(nside, p);
    psb::ShtLam q = psb::sht_lam_start(lmax, m, cm[m], g.z, g.s, true);
    for (int l = m; l <= lmax; ++l) lam[l - m] = 0.0;
    if (q.e == psb::SHT_NEVER) return -1;
    int alive = q.e == 0 ? m : lmax + 1;
    for (int l0 = m; l0 <= lmax; l0 += psb::SHT_C) {
        double2 cf[psb::SHT_C];
        psb::sht_coef_pass(l0, m, cf);
        for (int j = 0; j < psb::SHT_C; ++j) {
            const int l = l0 + j;
            if (l <= lmax && q.e == 0) lam[l - m] = q.lc * (j ? cf[j - 1].y : 1.0);
            psb::sht_mu_advance(q, cf[j].x);
        }
        psb::sht_mu_close(q, cf[psb::SHT_C - 2].y, cf[psb::SHT_C - 1].y);
        if (psb::sht_lam_rescale(q)) alive = l0 + psb::SHT_C;
    }
    return alive;
}

// ring FFT round trip pieces for direct tests: out[m] (mmax+1 complex) of one ring of n reals
extern "C" int sht_host_ring_analyse(const double* f, int n, int shifted, double scale, int mmax, double* out)
{
    const int h = n / 2;
    std::vector<double2> A(h), B(h), T(h);
    int rad[16];
    const int nrad = psb::sht_factor(h, rad);
    HostCtx cx;
    psb::sht_ring_analyse(cx, f, n, shifted, scale, mmax, A.data(), B.data(), T.data(), rad, nrad, (double2*)out, 1);
    return nrad;
}

extern "C" int sht_host_ring_synthesise(const double* in, int n, int shifted, int mmax, double* f)
{
    const int h = n / 2;
    std::vector<double2> A(h), B(h + 1), T(h);
    int rad[16];
    const int nrad = psb::sht_factor(h, rad);
    HostCtx cx;
    psb::sht_ring_synthesise(cx, (const double2*)in, 1, n, shifted, mmax, A.data(), B.data(), T.data(), T.data(), rad, nrad,
                             nullptr, f);
    return nrad;
}

extern "C" void sht_host_ring(int nside, int p, double* out)
{
    const psb::ShtRing g = psb::sht_ring(nside, p);
    out[0] = g.n; out[1] = (double)g.startN; out[2] = (double)g.startS; out[3] = g.z; out[4] = g.s; out[5] = g.shifted;
}
