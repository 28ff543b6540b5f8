/* shtcpu.c -- CPU restatement of the ring-based spin-0 HEALPix transforms, in the shape of the reference's CPU path.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline leg): never linked, loaded or executed by the product.
 *
 * The reference reaches these transforms through Healpix.jl (`map2alm(map; lmax, niter)` at
 * /root/reference/src/workspace.jl:155,163 and src/modecoupling.jl:251,255,328-329; `alm2cl` at src/workspace.jl:202), which
 * hands them to libsharp: per-ring real FFTs, then for every m an upward lambda_lm recurrence per ring with an integer
 * scale against underflow, north and south rings folded by parity, `niter` Jacobi iterations around it.  Neither package
 * is in this image; this file restates that published structure in plain C with OpenMP (threads over rings for the FFTs,
 * over m for the Legendre sums) so that (a) the direct-sum oracle (oracle/shtoracle.py) has a fast twin for map sizes it
 * cannot reach and (b) the benchmark has a CPU arm in the reference's shape.  It shares no code with csrc/psb200_sht.cuh:
 * recursive smallest-prime-first FFT on the full complex ring, per-step rescaling by 2^-512, no skipping of rings.
 *
 * Geometry (Gorski et al. 2005; HEALPix `pix2ang_ring`): ring r = 1..4 nside - 1, n_phi = 4 min(r, nside, 4 nside - r),
 * z = 1 - r^2/(3 nside^2) on the caps, (2 nside - r) 2/(3 nside) on the belt, first pixel at phi = pi/n_phi on the caps and
 * on belt rings with r - nside even, else 0.  Alm order: index(l, m) = m (2 lmax + 1 - m)/2 + l, interleaved (re, im).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

static const double PI = 3.14159265358979323846264338327950288;

typedef struct { int n; long start; double z, s; int shifted; } ring_t;

static ring_t ring_info(int nside, int r)
{
    ring_t g;
    const long N = nside, npix = 12 * N * N, ncap = 2 * N * (N - 1);
    const long rr = r < 4 * N - r ? r : 4 * N - r;
    if (rr < N) {
        const double t = (double)(rr * rr) / (3.0 * (double)N * (double)N);
        g.n = (int)(4 * rr);
        g.start = r < N ? 2 * rr * (rr - 1) : npix - 2 * rr * (rr + 1);
        g.z = r < N ? 1.0 - t : -(1.0 - t);
        g.s = sqrt(t * (2.0 - t));
        g.shifted = 1;
    } else {
        g.n = (int)(4 * N);
        g.start = ncap + (r - N) * 4 * N;
        g.z = (double)(2 * N - r) * 2.0 / (3.0 * (double)N);
        g.s = sqrt((1.0 - g.z) * (1.0 + g.z));
        g.shifted = ((r - N) & 1) == 0;
    }
    return g;
}

/* out[k] = sum_j in[j istride] exp(-2 pi i j k / n); root[t rstride] = exp(-2 pi i t / n) */
static void fft_rec(int n, const cplx* in, long istride, cplx* out, const cplx* root, long rstride)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = 2;
    while (n % p) ++p;
    const int m = n / p;
    for (int r = 0; r < p; ++r) fft_rec(m, in + r * istride, istride * p, out + (long)r * m, root, rstride * p);
    cplx* y = (cplx*)malloc((size_t)p * sizeof(cplx));
    cplx* x = (cplx*)malloc((size_t)p * sizeof(cplx));
    for (int k = 0; k < m; ++k) {
        for (int r = 0; r < p; ++r) y[r] = out[(long)r * m + k] * root[(long)r * k * rstride];
        for (int q = 0; q < p; ++q) {
            cplx a = 0.0;
            for (int r = 0; r < p; ++r) a += y[r] * root[(long)((r * q) % p) * m * rstride];
            x[q] = a;
        }
        for (int q = 0; q < p; ++q) out[k + (long)q * m] = x[q];
    }
    free(y); free(x);
}

static long alm_base(int lmax, int m) { return (long)m * (2 * lmax + 1 - m) / 2; }

/* phase[(m nr + (r-1))] = (4 pi / npix) sum_k f_k exp(-i m phi_k) for every ring r, m = 0..lmax */
static void rings_to_phase(int nside, int lmax, const double* map, cplx* phase)
{
    const int nr = 4 * nside - 1;
    const double w = 4.0 * PI / (12.0 * (double)nside * (double)nside);
#pragma omp parallel for schedule(dynamic, 4)
    for (int r = 1; r <= nr; ++r) {
        const ring_t g = ring_info(nside, r);
        const int n = g.n;
        cplx* in = (cplx*)malloc((size_t)n * sizeof(cplx));
        cplx* out = (cplx*)malloc((size_t)n * sizeof(cplx));
        cplx* root = (cplx*)malloc((size_t)n * sizeof(cplx));
        for (int t = 0; t < n; ++t) {
            root[t] = cos(2.0 * PI * t / n) - I * sin(2.0 * PI * t / n);
            in[t] = map[g.start + t];
        }
        fft_rec(n, in, 1, out, root, 1);
        for (int m = 0; m <= lmax; ++m) {
            cplx v = out[m % n];
            if (g.shifted) {
                const double a = PI * (double)(m % (2 * n)) / (double)n;
                v *= cos(a) - I * sin(a);
            }
            phase[(long)m * nr + (r - 1)] = w * v;
        }
        free(in); free(out); free(root);
    }
}

/* f_k = Re F_0 + 2 Re sum_{m>0} F_m exp(i m phi_k); the ring written is f, or ref - f when ref is given */
static void phase_to_rings(int nside, int lmax, const cplx* phase, const double* ref, double* map)
{
    const int nr = 4 * nside - 1;
#pragma omp parallel for schedule(dynamic, 4)
    for (int r = 1; r <= nr; ++r) {
        const ring_t g = ring_info(nside, r);
        const int n = g.n;
        cplx* in = (cplx*)calloc((size_t)n, sizeof(cplx));
        cplx* out = (cplx*)malloc((size_t)n * sizeof(cplx));
        cplx* root = (cplx*)malloc((size_t)n * sizeof(cplx));
        for (int t = 0; t < n; ++t) root[t] = cos(2.0 * PI * t / n) - I * sin(2.0 * PI * t / n);
        for (int m = 0; m <= lmax; ++m) {
            cplx v = phase[(long)m * nr + (r - 1)];
            if (g.shifted) {
                const double a = PI * (double)(m % (2 * n)) / (double)n;
                v *= cos(a) + I * sin(a);
            }
            if (m == 0) { in[0] += creal(v); continue; }
            /* f_k gets v w^{mk} + conj(v) w^{-mk}, w = exp(+2 pi i/n); with the forward transform below (exp(-...)) the
               coefficient of exp(-2 pi i j k/n) is wanted: j = -m mod n for v, j = m mod n for conj(v) */
            in[(n - m % n) % n] += v;
            in[m % n] += conj(v);
        }
        fft_rec(n, in, 1, out, root, 1);
        for (int t = 0; t < n; ++t) {
            const double f = creal(out[t]);
            map[g.start + t] = ref ? ref[g.start + t] - f : f;
        }
        free(in); free(out); free(root);
    }
}

static const double SC_BIG = 1.15792089237316195423570985008687907853e77;   /* 2^256 */
static const double SC_DOWN = 7.4583407312002067432909653154629338374e-155; /* 2^-512 */

static void log2_lmm(int lmax, double* cm)
{
    long double acc = 0.0L;
    for (int m = 0; m <= lmax; ++m) {
        if (m) acc += log2l((2.0L * m - 1.0L) / (2.0L * m));
        cm[m] = (double)(0.5L * (log2l((2.0L * m + 1.0L) / (4.0L * 3.14159265358979323846264338327950288L)) + acc));
    }
}

/* analysis = 1: alm (+)= sum over rings of lambda_lm phase;  analysis = 0: phase = sum over l of lambda_lm alm */
static void legendre(int nside, int lmax, int analysis, int accumulate, cplx* phase, double* alm)
{
    const int nr = 4 * nside - 1, nrp = 2 * nside;
    double* cm = (double*)malloc((size_t)(lmax + 1) * sizeof(double));
    log2_lmm(lmax, cm);
#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m <= lmax; ++m) {
        const long base = alm_base(lmax, m);
        const int nl = lmax - m + 1;
        double* c1 = (double*)malloc((size_t)(nl + 1) * sizeof(double));
        double* c2 = (double*)malloc((size_t)(nl + 1) * sizeof(double));
        double* sre = (double*)calloc((size_t)nl, sizeof(double));
        double* sim = (double*)calloc((size_t)nl, sizeof(double));
        for (int l = m; l <= lmax; ++l) {       /* lambda_{l+1} = c1 x lambda_l - c2 lambda_{l-1} */
            const double l1 = l + 1.0, ia2 = (4.0 * l1 * l1 - 1.0) / ((l1 - m) * (l1 + m));
            const double a2 = ((double)(l - m) * (double)(l + m)) / (4.0 * (double)l * (double)l - 1.0);
            c1[l - m] = sqrt(ia2);
            c2[l - m] = sqrt(a2 * ia2);
        }
        for (int p = 0; p < nrp; ++p) {
            const int rN = p + 1, rS = 4 * nside - rN;
            const ring_t g = ring_info(nside, rN);
            const int has_s = rS != rN;
            const double L = cm[m] + (double)m * log2(g.s);
            int e = L < -200.0 ? (int)ceil((-L - 200.0) / 512.0) : 0;
            double lc = exp2(L + 512.0 * e), lp = 0.0;
            if (m & 1) lc = -lc;
            cplx gN = 0.0, gS = 0.0, fe = 0.0, fo = 0.0;
            if (analysis) { gN = phase[(long)m * nr + (rN - 1)]; if (has_s) gS = phase[(long)m * nr + (rS - 1)]; }
            const cplx ge = gN + gS, go = gN - gS;
            for (int l = m; l <= lmax; ++l) {
                if (e == 0) {
                    if (analysis) {
                        const cplx v = ((l - m) & 1) ? go : ge;
                        sre[l - m] += lc * creal(v);
                        sim[l - m] += lc * cimag(v);
                    } else {
                        const cplx a = alm[2 * (base + l)] + I * alm[2 * (base + l) + 1];
                        if ((l - m) & 1) fo += lc * a; else fe += lc * a;
                    }
                }
                const double ln = c1[l - m] * (g.z * lc) - c2[l - m] * lp;
                lp = lc; lc = ln;
                if (e > 0 && fabs(lc) > SC_BIG) { lc *= SC_DOWN; lp *= SC_DOWN; --e; }
            }
            if (!analysis) {
                phase[(long)m * nr + (rN - 1)] = fe + fo;
                if (has_s) phase[(long)m * nr + (rS - 1)] = fe - fo;
            }
        }
        if (analysis)
            for (int l = m; l <= lmax; ++l) {
                double* a = alm + 2 * (base + l);
                a[0] = accumulate ? a[0] + sre[l - m] : sre[l - m];
                a[1] = accumulate ? a[1] + sim[l - m] : sim[l - m];
            }
        free(c1); free(c2); free(sre); free(sim);
    }
    free(cm);
}

/* Healpix.jl map2alm(map; lmax, niter): pixel-weighted analysis + niter Jacobi iterations.  alm: (lmax+1)(lmax+2) doubles */
int shtcpu_map2alm(int nside, int lmax, int niter, const double* map, double* alm)
{
    const long npix = 12L * nside * nside, nr = 4L * nside - 1;
    cplx* phase = (cplx*)malloc((size_t)(lmax + 1) * nr * sizeof(cplx));
    double* resid = (double*)malloc((size_t)npix * sizeof(double));
    if (!phase || !resid) { free(phase); free(resid); return 1; }
    rings_to_phase(nside, lmax, map, phase);
    legendre(nside, lmax, 1, 0, phase, alm);
    for (int it = 0; it < niter; ++it) {
        legendre(nside, lmax, 0, 0, phase, alm);
        phase_to_rings(nside, lmax, phase, map, resid);
        rings_to_phase(nside, lmax, resid, phase);
        legendre(nside, lmax, 1, 1, phase, alm);
    }
    free(phase); free(resid);
    return 0;
}

int shtcpu_alm2map(int nside, int lmax, const double* alm, double* map)
{
    const long nr = 4L * nside - 1;
    cplx* phase = (cplx*)malloc((size_t)(lmax + 1) * nr * sizeof(cplx));
    if (!phase) return 1;
    legendre(nside, lmax, 0, 0, phase, (double*)alm);
    phase_to_rings(nside, lmax, phase, NULL, map);
    free(phase);
    return 0;
}

int shtcpu_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}
