// psb200_sht.cuh -- W-spectrum production: spin-0 HEALPix map2alm / alm2map / alm2cl on the device.
//
// What it replaces on the host side of the reference (which stays available):
//   effective_weight_alm!   /root/reference/src/workspace.jl:141-171   map2alm(mask_i .* mask_j [.* sigma^2 .* Omega_pix]; lmax)
//   window_function_W!      /root/reference/src/workspace.jl:174-213   mean over (wX, wY) of alm2cl(w_X, w_Y)[0:lmax]
//   map2alm(mask) feeding mcm   src/modecoupling.jl:250-256, :328-329
// `map2alm` / `alm2cl` themselves live in Healpix.jl (un-vendored; compat "3, 4"): pixel-weighted analysis on the RING
// pixelisation followed by `niter` Jacobi iterations a <- a + A(f - S a) (default 3).  This file is a from-scratch
// B200 design of that transform, not a port of libsharp:
//
//   map --(ring kernel: one block per ring, mixed-radix Stockham FFT in shared memory, any ring length 4r)--> Phi[m][ring pair]
//       --(Legendre kernel: one WARP per (m, chunk of 32 R ring pairs); every lane runs the lambda_lm recurrence of R ring
//          pairs in registers along l, north/south folded by parity; per 16 l the lanes' partial sums are combined by a
//          31-shuffle butterfly and written once)--> per-chunk partial alm --(finish: fixed-order sum over chunks)--> alm
// and the mirror image for synthesis (lanes accumulate F_m(ring) in registers; no reduction at all).  Nothing is
// atomically accumulated, so results do not depend on scheduling.  The Legendre stage is FP64-pipe bound
// (4 FP64 instructions per (l, m, ring pair): 2 of the un-normalised recurrence + 2 accumulate), the ring stage is
// shared-memory bound and small.
//
// Dynamic range.  lambda_mm ~ sin^m(theta) underflows Float64 long before l reaches the classical region
// (theta = 1e-3, m = 3000: 1e-9000).  Every ring carries an integer e with lambda = lt * 2^(-800 e), |lt| kept inside
// [2^-400, 2^487]; the start value comes from log2 lambda_mm = cm[m] + m log2 sin(theta) (cm tabulated on the host in long
// double); the check runs once per 16 steps (growth per 16 steps < 2^87); a ring contributes only while e = 0 (what it
// would add before is below 2^-400 of the scale of lambda).  Rings with m > lmax sin(theta) + 16 (lmax/2)^(1/3) are never
// started (sht_mlim: everything skipped is below 1e-30).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef PSB_HD
#if defined(__CUDACC__)
#define PSB_HD __host__ __device__ __forceinline__
#else
#define PSB_HD inline
#endif
#endif

namespace psb {

constexpr int SHT_C = 16;                 // l steps between reductions / rescale checks (8: same time per step, measured)
constexpr int SHT_WARPS = 4;              // warps per block of the Legendre kernels (independent of each other)
constexpr double SHT_BIG = 2.58224987808690858965591917200e120;      // 2^400
constexpr double SHT_DOWN = 1.49969681389882132575131978222e-241;    // 2^-800
constexpr int SHT_NEVER = 0x7fffffff;

struct ShtDims {
    int nside, lmax, nrp, nchunks;        // nrp = 2 nside ring pairs (north ring p+1, south ring 4 nside - p - 1; p = nrp-1: the equator alone)
    long long npix, nalm;                 // 12 nside^2, (lmax+1)(lmax+2)/2
};

PSB_HD long long sht_alm_base(int lmax, int m) { return (long long)m * (2 * lmax + 1 - m) / 2; }   // index(l, m) = base + l

// ring pair p: pixels per ring, first pixel of the north / south ring, cos and sin of the north colatitude, and
// whether the first pixel sits at phi = pi/n (caps, belt rings with r - nside even) or at 0
struct ShtRing { int n; long long startN, startS; double z, s; int shifted; };

PSB_HD ShtRing sht_ring(int nside, int p)
{
    ShtRing g;
    const long long N = nside, r = p + 1, npix = 12 * N * N, ncap = 2 * N * (N - 1);
    if (r < N) {
        const double tmp = (double)(r * r) / (3.0 * (double)N * (double)N);
        g.n = (int)(4 * r);
        g.startN = 2 * r * (r - 1);
        g.startS = npix - 2 * r * (r + 1);
        g.z = 1.0 - tmp;
        g.s = sqrt(tmp * (2.0 - tmp));
        g.shifted = 1;
    } else {
        g.n = (int)(4 * N);
        g.startN = ncap + (r - N) * 4 * N;
        g.startS = ncap + (3 * N - r) * 4 * N;                 // ring 4N - r
        g.z = (double)(2 * N - r) * 2.0 / (3.0 * (double)N);
        g.s = sqrt((1.0 + g.z) * (1.0 - g.z));
        g.shifted = ((r - N) & 1) == 0;
    }
    return g;
}

// rings with m above this are never started.  The margin is 16 turning-point widths (lmax/2)^(1/3): every lambda_lm
// it skips is below 1e-30 (measured, lmax 191..8191; tests/test_sht.py).  libsharp's `sharp_get_mlim`, which the
// reference's transform uses, takes lmax sin(theta) + max(100, lmax/100) and drops values up to 1e-9 at lmax 6143.
PSB_HD double sht_mlim(int lmax, double sth)
{
    const double w = 16.0 * cbrt(0.5 * (double)lmax);
    return lmax * sth + (w < 50.0 ? 50.0 : w);
}

// Recurrence inside a pass of SHT_C steps starting at l0 (l0 - m a multiple of SHT_C).  With A_l^2 = (l^2 - m^2)/(4 l^2 - 1)
// the normalised functions obey lambda_{l+1} = (x lambda_l - A_l lambda_{l-1})/A_{l+1}; the kernels step the UN-normalised
// mu_j = lambda_{l0+j}/Q_j, Q_0 = 1, Q_{j+1} = Q_j/A_{l0+j+1}, which obeys
//     mu_{j+1} = x mu_j - d_j mu_{j-1},   d_0 = A_{l0} (mu_{-1} = lambda_{l0-1}),  d_j = A_{l0+j}^2 (j >= 1)
// -- one multiply and one FMA per step instead of two multiplies and one FMA, no square root in d_j.  Q_j only depends on
// (l, m): the analysis multiplies the REDUCED sum of step j by it, the synthesis folds it into a_lm when it stages them,
// and the pass ends with lambda_{l0+16} = Q_16 mu_16, lambda_{l0+15} = Q_15 mu_15.  Q_16 < 2^90 (first pass, m ~ lmax).
// Table entry of step j: (d_j, Q_{j+1}); column m holds the entries of l = m .. lmax + SHT_C - 1 (whole passes).
PSB_HD long long sht_coef_base(int lmax, int m) { return sht_alm_base(lmax, m) + (long long)(SHT_C - 1) * m; }   // index = base + l
PSB_HD long long sht_coef_size(int lmax) { return sht_alm_base(lmax, lmax) + lmax + (long long)(SHT_C - 1) * lmax + SHT_C; }

PSB_HD void sht_coef_pass(int l0, int m, double2* out)      // out[0..SHT_C-1]
{
    double Q = 1.0;
    for (int j = 0; j < SHT_C; ++j) {
        const double l = (double)(l0 + j), l1 = l + 1.0, md = (double)m;
        const double a2 = ((l - md) * (l + md)) / (4.0 * l * l - 1.0);                 // A_l^2 (0 at l = m; l = m = 0: -0/-1)
        const double ia2 = (4.0 * l1 * l1 - 1.0) / ((l1 - md) * (l1 + md));            // 1/A_{l+1}^2
        Q *= sqrt(ia2);
        out[j] = make_double2(j ? a2 : sqrt(fabs(a2)), Q);
    }
}

// state of one ring's recurrence: lambda_l = lc 2^(-800 e), lambda_{l-1} = lp 2^(-800 e)
struct ShtLam { double x, lp, lc; int e; };

PSB_HD ShtLam sht_lam_start(int lmax, int m, double cm_m, double z, double sth, bool exists)
{
    ShtLam q;
    q.x = z; q.lp = 0.0; q.lc = 0.0; q.e = SHT_NEVER;
    if (!exists || (double)m > sht_mlim(lmax, sth)) { q.x = 0.0; return q; }
    const double L = cm_m + (double)m * log2(sth);
    int e = 0;
    if (L < -400.0) e = (int)floor((400.0 - L) / 800.0);
    q.e = e;
    const double v = exp2(L + 800.0 * (double)e);
    q.lc = (m & 1) ? -v : v;
    return q;
}

// one step of the un-normalised recurrence on (lp, lc) = (mu_{j-1}, mu_j)
PSB_HD void sht_mu_advance(ShtLam& q, double d)
{
    const double t = d * q.lp;
    const double n = fma(q.x, q.lc, -t);
    q.lp = q.lc;
    q.lc = n;
}

// end of a pass: back to lambda.  q15 = Q_15 (1 when SHT_C == 1), q16 = Q_16
PSB_HD void sht_mu_close(ShtLam& q, double q15, double q16)
{
    q.lp *= q15;
    q.lc *= q16;
}

// true when the ring has just become representable (e reached 0)
PSB_HD bool sht_lam_rescale(ShtLam& q)
{
    if (q.e > 0 && q.e != SHT_NEVER && (fabs(q.lc) > SHT_BIG || fabs(q.lp) > SHT_BIG)) {
        q.lc *= SHT_DOWN;
        q.lp *= SHT_DOWN;
        return --q.e == 0;
    }
    return false;
}

PSB_HD void sht_sincospi(double a, double* s, double* c)
{
#if defined(__CUDA_ARCH__)
    sincospi(a, s, c);
#else
    // host build (tests/hostcheck): reduce exactly, then libm
    double r = fmod(a, 2.0);
    *s = sin(3.14159265358979323846264338327950288 * r);
    *c = cos(3.14159265358979323846264338327950288 * r);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// Ring stage.  All functions are written against a context {tid, nthr, sync()} so that tests/hostcheck can run the
// very same code with one "thread"; on the device the context is the thread block.
// ------------------------------------------------------------------------------------------------------------------

// radices of h: the largest odd prime first (the first stage has no stage twiddles, which lets sht_fft evaluate its
// outputs q and p - q from one pass over the inputs), then 4s, a 2, and the remaining primes in increasing order
PSB_HD int sht_factor(int h, int* rad)
{
    int n = 0;
    while (h % 4 == 0) { rad[n++] = 4; h /= 4; }
    for (int p = 2; p * p <= h; ++p)
        while (h % p == 0) { rad[n++] = p; h /= p; }
    if (h > 1) rad[n++] = h;
    if (n > 1 && (rad[n - 1] & 1)) {
        const int big = rad[n - 1];
        for (int i = n - 1; i > 0; --i) rad[i] = rad[i - 1];
        rad[0] = big;
    }
    return n;
}

// T[t] = exp(-2 pi i t / h)
template <class Ctx> PSB_HD void sht_twiddles(Ctx& cx, double2* T, int h)
{
    for (int t = cx.tid; t < h; t += cx.nthr) {
        double s, c;
        sht_sincospi(2.0 * (double)t / (double)h, &s, &c);
        T[t] = make_double2(c, -s);
    }
}

// Stockham autosort FFT of the h complex numbers in A (result returned in the buffer the function returns).
// Radix 4 and 2: one thread per butterfly, inputs read once (4 + 3 shared loads for 4 outputs); the largest odd prime p
// goes first and costs h (p+1)/2 half-price complex MACs (a prime ring length is O(h p): the HEALPix caps have them all);
// any other odd radix: one thread per output, p complex MACs against the table.
template <class Ctx> PSB_HD double2* sht_fft(Ctx& cx, double2* A, double2* B, const double2* T, int h, const int* rad, int nrad)
{
    int Ns = 1;
    for (int q = 0; q < nrad; ++q) {
        const int Rr = rad[q], M = Ns * Rr, hR = h / Rr, hM = h / M;
        if (q == 0 && (Rr & 1)) {
            // first stage, odd radix p (Ns = 1): out[j p + q] = sum_r A[j + r h/p] w^(q r), w = exp(-2 pi i / p); outputs q and
            // p - q share every load and half the products (w^((p-q) r) = conj w^(q r)); consecutive threads take consecutive
            // j, so the table load is a warp-wide broadcast
            const int hq = (Rr + 1) / 2;
            for (int it = cx.tid; it < hR * hq; it += cx.nthr) {
                const int j = it % hR, qi = it / hR;
                const int step = qi * hR;
                int idx = 0;
                double P = 0.0, Q = 0.0, U = 0.0, V = 0.0;
                for (int r = 0; r < Rr; ++r) {
                    const double2 a = A[j + r * hR];
                    const double2 w = T[idx];
                    P = fma(a.x, w.x, P); Q = fma(a.y, w.y, Q);
                    U = fma(a.x, w.y, U); V = fma(a.y, w.x, V);
                    idx += step;
                    if (idx >= h) idx -= h;
                }
                B[j * Rr + qi] = make_double2(P - Q, U + V);
                if (qi) B[j * Rr + Rr - qi] = make_double2(P + Q, V - U);
            }
        } else if (Rr == 4) {
            for (int j = cx.tid; j < hR; j += cx.nthr) {
                const int k = j % Ns, j0 = (j / Ns) * M + k;
                const double2 a0 = A[j], a1 = A[j + hR], a2 = A[j + 2 * hR], a3 = A[j + 3 * hR];
                const double2 w1 = T[k * hM], w2 = T[2 * k * hM], w3 = T[3 * k * hM];          // 3 k hM < 3 h / 4
                const double b1r = a1.x * w1.x - a1.y * w1.y, b1i = a1.x * w1.y + a1.y * w1.x;
                const double b2r = a2.x * w2.x - a2.y * w2.y, b2i = a2.x * w2.y + a2.y * w2.x;
                const double b3r = a3.x * w3.x - a3.y * w3.y, b3i = a3.x * w3.y + a3.y * w3.x;
                const double s0r = a0.x + b2r, s0i = a0.y + b2i, d0r = a0.x - b2r, d0i = a0.y - b2i;
                const double s1r = b1r + b3r, s1i = b1i + b3i, d1r = b1r - b3r, d1i = b1i - b3i;
                B[j0] = make_double2(s0r + s1r, s0i + s1i);
                B[j0 + Ns] = make_double2(d0r + d1i, d0i - d1r);              // a0 - i b1 - b2 + i b3
                B[j0 + 2 * Ns] = make_double2(s0r - s1r, s0i - s1i);
                B[j0 + 3 * Ns] = make_double2(d0r - d1i, d0i + d1r);          // a0 + i b1 - b2 - i b3
            }
        } else if (Rr == 2) {
            for (int j = cx.tid; j < hR; j += cx.nthr) {
                const int k = j % Ns, j0 = (j / Ns) * M + k;
                const double2 a0 = A[j], a1 = A[j + hR];
                const double2 w1 = T[k * hM];
                const double b1r = a1.x * w1.x - a1.y * w1.y, b1i = a1.x * w1.y + a1.y * w1.x;
                B[j0] = make_double2(a0.x + b1r, a0.y + b1i);
                B[j0 + Ns] = make_double2(a0.x - b1r, a0.y - b1i);
            }
        } else {
            for (int o = cx.tid; o < h; o += cx.nthr) {
                const int k = o % Ns, qq = (o / Ns) % Rr, jhi = o / M;
                const int j = jhi * Ns + k;
                int step = k * hM + qq * hR;
                if (step >= h) step -= h;
                int idx = 0;
                double ar = 0.0, ai = 0.0;
                for (int r = 0; r < Rr; ++r) {
                    const double2 a = A[j + r * hR];
                    const double2 w = T[idx];
                    ar = fma(a.x, w.x, ar); ar = fma(-a.y, w.y, ar);
                    ai = fma(a.x, w.y, ai); ai = fma(a.y, w.x, ai);
                    idx += step;
                    if (idx >= h) idx -= h;
                }
                B[o] = make_double2(ar, ai);
            }
        }
        cx.sync();
        double2* t = A; A = B; B = t;
        Ns = M;
    }
    return A;
}

// Analysis of one ring: f (n reals) -> out[m] = scale * exp(-i m phi0) * sum_k f_k exp(-2 pi i m k / n), m = 0..mmax,
// written with stride `ostride` (in double2 units).  A, B, T: h = n/2 complex numbers each (shared memory on the device).
template <class Ctx>
PSB_HD void sht_ring_analyse(Ctx& cx, const double* f, int n, int shifted, double scale, int mmax,
                             double2* A, double2* B, double2* T, const int* rad, int nrad, double2* out, long long ostride)
{
    const int h = n / 2;
    sht_twiddles(cx, T, h);
    for (int k = cx.tid; k < h; k += cx.nthr) A[k] = make_double2(f[2 * k], f[2 * k + 1]);
    cx.sync();
    double2* Z = sht_fft(cx, A, B, T, h, rad, nrad);
    double2* X = (Z == A) ? B : A;                       // X[0..h-1]; X[h] is real and kept apart (no room for h+1 entries)
    // X[j] = 1/2 [(Z_j + conj Z_{h-j}) - i w_j (Z_j - conj Z_{h-j})],  w_j = exp(-2 pi i j / n)
    for (int j = cx.tid; j < h; j += cx.nthr) {
        const double2 a = Z[j], b = Z[j ? h - j : 0];
        double s, c;
        sht_sincospi((double)j / (double)h, &s, &c);     // w = c - i s
        const double er = 0.5 * (a.x + b.x), ei = 0.5 * (a.y - b.y);       // E = (Z_j + conj Z_{h-j})/2
        const double dr = 0.5 * (a.x - b.x), di = 0.5 * (a.y + b.y);       // D = (Z_j - conj Z_{h-j})/2
        // -i w D = -i (c - i s)(dr + i di) = (c di - s dr) - i (c dr + s di)
        X[j] = make_double2(er + (c * di - s * dr), ei - (c * dr + s * di));
    }
    cx.sync();
    const double xh = Z[0].x - Z[0].y;                   // X[h] = Re Z_0 - Im Z_0
    for (int m = cx.tid; m <= mmax; m += cx.nthr) {
        int j = m % n;
        double2 v;
        if (j < h) v = X[j];
        else if (j == h) v = make_double2(xh, 0.0);
        else { v = X[n - j]; v.y = -v.y; }
        double pr = scale, pi = 0.0;
        if (shifted) {                                   // exp(-i pi m / n)
            double s, c;
            sht_sincospi((double)(m % (2 * n)) / (double)n, &s, &c);
            pr = scale * c; pi = -scale * s;
        }
        out[(long long)m * ostride] = make_double2(v.x * pr - v.y * pi, v.x * pi + v.y * pr);
    }
    cx.sync();
}

// Synthesis of one ring: in[m] = F_m (stride istride) -> f_k = Re F_0 + 2 Re sum_{m>0} F_m exp(i m phi_k), n reals.
// If `ref` is given the ring written is ref - f (the residual of a Jacobi iteration).
template <class Ctx>
PSB_HD void sht_ring_synthesise(Ctx& cx, const double2* in, long long istride, int n, int shifted, int mmax,
                                double2* A, double2* B, double2* T, double2* W, const int* rad, int nrad, const double* ref,
                                double* f)
{
    const int h = n / 2;
    sht_twiddles(cx, T, h);
    // S[j], j = 0..h (B holds h+1 entries):  sum_{m = j, j+n, ...} P_m  +  sum_{m = n-j, 2n-j, ... > 0} conj P_m,
    // P_m = F_m exp(i m phi0), P_0 -> Re.  Short rings fold mmax/n terms per entry: the terms of an entry are dealt to
    // G = nthr/(h+1) threads whose partial sums (W, nthr entries) are added in a fixed order.
    const int G = cx.nthr / (h + 1) > 1 ? cx.nthr / (h + 1) : 1;
    for (int it = cx.tid; it < (h + 1) * G; it += cx.nthr) {
        const int j = it % (h + 1), g = it / (h + 1);
        double sr = 0.0, si = 0.0;
        for (int m = j + g * n; m <= mmax; m += G * n) {
            double2 v = in[(long long)m * istride];
            double pr = 1.0, pi = 0.0;
            if (shifted) { double s, c; sht_sincospi((double)(m % (2 * n)) / (double)n, &s, &c); pr = c; pi = s; }
            sr += v.x * pr - v.y * pi;
            si += v.x * pi + v.y * pr;
        }
        for (int m = n - j + g * n; m <= mmax; m += G * n) {
            double2 v = in[(long long)m * istride];
            double pr = 1.0, pi = 0.0;
            if (shifted) { double s, c; sht_sincospi((double)(m % (2 * n)) / (double)n, &s, &c); pr = c; pi = s; }
            sr += v.x * pr - v.y * pi;
            si -= v.x * pi + v.y * pr;
        }
        if (G > 1) W[g * (h + 1) + j] = make_double2(sr, si);
        else B[j] = make_double2(sr, (j == 0 || j == h) ? 0.0 : si);          // S_0 and S_h are real
    }
    cx.sync();
    if (G > 1) {
        for (int j = cx.tid; j <= h; j += cx.nthr) {
            double sr = 0.0, si = 0.0;
            for (int g = 0; g < G; ++g) { sr += W[g * (h + 1) + j].x; si += W[g * (h + 1) + j].y; }
            B[j] = make_double2(sr, (j == 0 || j == h) ? 0.0 : si);
        }
        cx.sync();
    }
    // conj Z'_j,  Z'_j = (S_j + conj S_{h-j}) + i exp(2 pi i j/n) (S_j - conj S_{h-j})
    for (int j = cx.tid; j < h; j += cx.nthr) {
        const double2 a = B[j];
        const double2 b = B[h - j];
        double s, c;
        sht_sincospi((double)j / (double)h, &s, &c);     // exp(+2 pi i j / n) = c + i s
        const double er = a.x + b.x, ei = a.y - b.y;
        const double dr = a.x - b.x, di = a.y + b.y;
        // i (c + i s)(dr + i di) = i (c dr - s di) - (c di + s dr)
        const double zr = er - (c * di + s * dr), zi = ei + (c * dr - s * di);
        A[j] = make_double2(zr, -zi);
    }
    cx.sync();
    double2* Z = sht_fft(cx, A, B, T, h, rad, nrad);     // conj of the inverse transform
    for (int k = cx.tid; k < h; k += cx.nthr) {
        const double f0 = Z[k].x, f1 = -Z[k].y;
        if (ref) { f[2 * k] = ref[2 * k] - f0; f[2 * k + 1] = ref[2 * k + 1] - f1; }
        else { f[2 * k] = f0; f[2 * k + 1] = f1; }
    }
    cx.sync();
}

#if defined(__CUDACC__)

struct ShtBlockCtx {
    int tid, nthr;
    __device__ void sync() { __syncthreads(); }
};

// Phi[(m nrp + p) 2 + hemi] (double2): hemi 0 = north ring p+1, 1 = its southern partner (zero for the equator)
// grid.x = ring pairs [p_lo, p_lo + gridDim.x), grid.y = 2 hemispheres; dynamic shared memory (3 h_max + 1 + blockDim.x) double2
__global__ void __launch_bounds__(512) sht_ring_analysis_kernel(ShtDims D, int p_lo, const double* __restrict__ map,
                                                                double2* __restrict__ Phi)
{
    extern __shared__ double2 sht_smem[];
    __shared__ int rad[16];
    __shared__ int nrad;
    const int p = p_lo + blockIdx.x, hemi = blockIdx.y;
    const ShtRing g = sht_ring(D.nside, p);
    double2* out = Phi + ((long long)p * 2 + hemi);
    const long long ostride = (long long)D.nrp * 2;
    if (hemi == 1 && p == D.nrp - 1) {                 // the equator has no partner
        for (int m = threadIdx.x; m <= D.lmax; m += blockDim.x) out[(long long)m * ostride] = make_double2(0.0, 0.0);
        return;
    }
    const int h = g.n / 2;
    if (threadIdx.x == 0) nrad = sht_factor(h, rad);
    __syncthreads();
    ShtBlockCtx cx{(int)threadIdx.x, (int)blockDim.x};
    sht_ring_analyse(cx, map + (hemi ? g.startS : g.startN), g.n, g.shifted, 12.566370614359172953850573533118 / (double)D.npix,
                     D.lmax, sht_smem, sht_smem + h, sht_smem + 2 * h, rad, nrad, out, ostride);
}

__global__ void __launch_bounds__(512) sht_ring_synthesis_kernel(ShtDims D, int p_lo, const double2* __restrict__ Phi,
                                                                 const double* __restrict__ ref, double* __restrict__ map)
{
    extern __shared__ double2 sht_smem[];
    __shared__ int rad[16];
    __shared__ int nrad;
    const int p = p_lo + blockIdx.x, hemi = blockIdx.y;
    if (hemi == 1 && p == D.nrp - 1) return;
    const ShtRing g = sht_ring(D.nside, p);
    const int h = g.n / 2;
    if (threadIdx.x == 0) nrad = sht_factor(h, rad);
    __syncthreads();
    ShtBlockCtx cx{(int)threadIdx.x, (int)blockDim.x};
    const long long st = hemi ? g.startS : g.startN;
    sht_ring_synthesise(cx, Phi + ((long long)p * 2 + hemi), (long long)D.nrp * 2, g.n, g.shifted, D.lmax,
                        sht_smem, sht_smem + h, sht_smem + 2 * h + 1, sht_smem + 3 * h + 1, rad, nrad, ref ? ref + st : nullptr,
                        map + st);
}

// coef[sht_coef_base(m) + l] = (d_j, Q_{j+1}) of sht_coef_pass: one thread per (m, pass)
__global__ void sht_coef_kernel(int lmax, double2* __restrict__ coef)
{
    const int m = blockIdx.y;
    const int l0 = m + (blockIdx.x * blockDim.x + threadIdx.x) * SHT_C;
    if (l0 > lmax) return;
    double2 e[SHT_C];
    sht_coef_pass(l0, m, e);
    double2* out = coef + sht_coef_base(lmax, m) + l0;
#pragma unroll
    for (int j = 0; j < SHT_C; ++j) out[j] = e[j];
}

// cmin[m] = first chunk of ring pairs that holds a ring the transform starts for this m (rings are skipped from the
// poles inwards: sin(theta) grows with p); chunk c covers ring pairs [c 32 R, (c+1) 32 R)
__global__ void sht_cmin_kernel(ShtDims D, int R, int* __restrict__ cmin)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > D.lmax) return;
    int c = 0;
    for (; c < D.nchunks; ++c) {
        int plast = (c + 1) * 32 * R - 1;
        if (plast > D.nrp - 1) plast = D.nrp - 1;
        if ((double)m <= sht_mlim(D.lmax, sht_ring(D.nside, plast).s)) break;
    }
    cmin[m] = c;
}

// butterfly of the analysis: every lane holds the partial sums of 16 l steps, v[2j] and v[2j+1]; EVEN lanes hold
// (re, im) there and ODD lanes (im, re) -- they swap their G registers when they load them -- so the first exchange,
// between the lanes of a pair, needs no selects: afterwards even lanes own the 16 real parts and odd lanes the 16
// imaginary parts.  Four select-and-exchange stages (xor 16, 8, 4, 2) then halve 16 -> 1: on return v[0] of lane L is
// the warp total of step L >> 1, component L & 1.  31 shuffles of 64 bits; the order of the additions is fixed.
__device__ __forceinline__ void sht_butterfly(double (&v)[32], int lane)
{
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = v[2 * j] + __shfl_xor_sync(0xffffffffu, v[2 * j + 1], 1);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int d = 16 >> t, half = 8 >> t;          // compile-time once unrolled: v stays in registers
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
}

// the same for 8 l steps (v[0..15]): three select-and-exchange stages (xor 16, 8, 4), then the lanes L and L ^ 2 hold two
// halves of the same sum: v[0] of lane L is the warp total of step L >> 2, component L & 1 (in both lanes of such a pair)
__device__ __forceinline__ void sht_butterfly8(double (&v)[16], int lane)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[2 * j] + __shfl_xor_sync(0xffffffffu, v[2 * j + 1], 1);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int d = 16 >> t, half = 4 >> t;
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
}

// Legendre stage of the analysis: one warp per (m, chunk).  partial[chunk][2 (base(m) + l) + {0 re, 1 im}]
// R ring pairs per lane, C = 16 l steps between reductions; the coefficients of the next C steps are in flight while the
// current ones are used; while no ring of the warp is representable yet the warp only advances the recurrences.
// V = l steps per reduction (16: one butterfly per pass; 8: two, with 32 registers less)
template <int R, int V>
__global__ void __launch_bounds__(32 * SHT_WARPS, (R <= 4 && V == 8) ? 5 : 1) sht_leg_analysis_kernel(ShtDims D, const double4* __restrict__ Phi,
                                                                         const double2* __restrict__ coef,
                                                                         const double* __restrict__ cm,
                                                                         const int* __restrict__ cmin,
                                                                         double* __restrict__ partial)
{
    constexpr int C = SHT_C;
    __shared__ double2 sco[SHT_WARPS][C];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool odd = (lane & 1) != 0;                  // odd lanes keep (im, re) where even lanes keep (re, im): sht_butterfly
    const int bpm = (D.nchunks + SHT_WARPS - 1) / SHT_WARPS;
    const int m = blockIdx.x / bpm;
    const int chunk = (blockIdx.x % bpm) * SHT_WARPS + wid;
    if (chunk >= D.nchunks || chunk < cmin[m]) return;
    const long long base = sht_alm_base(D.lmax, m);
    const double2* cf = coef + sht_coef_base(D.lmax, m);
    const double cm_m = cm[m];
    ShtLam q[R];
    double ger[R], gei[R], gor[R], goi[R];
    const int p0 = (chunk * 32 + lane) * R;
    bool alive = false;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int p = p0 + s;
        const bool ex = p < D.nrp;
        const ShtRing g = sht_ring(D.nside, ex ? p : 0);
        q[s] = sht_lam_start(D.lmax, m, cm_m, g.z, g.s, ex);
        ger[s] = gei[s] = gor[s] = goi[s] = 0.0;
        if (q[s].e == 0) {
            const double4 t = Phi[(long long)m * D.nrp + p];
            ger[s] = odd ? t.y + t.w : t.x + t.z; gei[s] = odd ? t.x + t.z : t.y + t.w;
            gor[s] = odd ? t.y - t.w : t.x - t.z; goi[s] = odd ? t.x - t.z : t.y - t.w;
            alive = true;
        }
    }
    double* out = partial + (long long)chunk * 2 * D.nalm + 2 * base;
    double2 cnext = make_double2(0.0, 0.0);
    if (lane < C) cnext = cf[m + lane];
    for (int l0 = m; l0 <= D.lmax; l0 += C) {
        if (lane < C) sco[wid][lane] = cnext;
        __syncwarp();
        if (lane < C && l0 + C <= D.lmax) cnext = cf[l0 + C + lane];
        if (__any_sync(0xffffffffu, alive)) {
#pragma unroll
            for (int hh = 0; hh < C / V; ++hh) {
                double v[2 * V];
#pragma unroll
                for (int i = 0; i < 2 * V; ++i) v[i] = 0.0;
#pragma unroll
                for (int jj = 0; jj < V; ++jj) {
                    const int j = hh * V + jj;
                    const double2 c = sco[wid][j];
#pragma unroll
                    for (int s = 0; s < R; ++s) {
                        if ((j & 1) == 0) { v[2 * jj] = fma(q[s].lc, ger[s], v[2 * jj]); v[2 * jj + 1] = fma(q[s].lc, gei[s], v[2 * jj + 1]); }
                        else              { v[2 * jj] = fma(q[s].lc, gor[s], v[2 * jj]); v[2 * jj + 1] = fma(q[s].lc, goi[s], v[2 * jj + 1]); }
                        sht_mu_advance(q[s], c.x);
                    }
                }
                // the sums were of mu_j: times Q_j
                if constexpr (V == 16) {
                    sht_butterfly(v, lane);
                    const int jl = lane >> 1, l = l0 + jl;
                    const double Qj = jl ? sco[wid][jl - 1].y : 1.0;
                    if (l <= D.lmax) out[2 * (long long)l + (lane & 1)] = v[0] * Qj;
                } else {
                    sht_butterfly8(v, lane);
                    const int jl = hh * V + (lane >> 2), l = l0 + jl;
                    const double Qj = jl ? sco[wid][jl - 1].y : 1.0;
                    if (l <= D.lmax && (lane & 2) == 0) out[2 * (long long)l + (lane & 1)] = v[0] * Qj;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const double d = sco[wid][j].x;
#pragma unroll
                for (int s = 0; s < R; ++s) sht_mu_advance(q[s], d);
            }
            if (lane < 2 * C && l0 + (lane >> 1) <= D.lmax) out[2 * (long long)l0 + lane] = 0.0;
        }
        {
            const double q15 = sco[wid][C - 2].y, q16 = sco[wid][C - 1].y;
#pragma unroll
            for (int s = 0; s < R; ++s) sht_mu_close(q[s], q15, q16);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (sht_lam_rescale(q[s])) {
                const double4 t = Phi[(long long)m * D.nrp + p0 + s];
                ger[s] = odd ? t.y + t.w : t.x + t.z; gei[s] = odd ? t.x + t.z : t.y + t.w;
                gor[s] = odd ? t.y - t.w : t.x - t.z; goi[s] = odd ? t.x - t.z : t.y - t.w;
                alive = true;
            }
        }
    }
}

// sum of the per-chunk partial alm in chunk order; accumulate != 0: alm += sum (Jacobi step)
__global__ void sht_analysis_finish_kernel(ShtDims D, const double* __restrict__ partial, const int* __restrict__ cmin,
                                           int accumulate, double* __restrict__ alm)
{
    const int m = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // double index inside the m column: 2 (l - m) + comp
    if (i >= 2 * (D.lmax - m + 1)) return;
    const long long off = 2 * (sht_alm_base(D.lmax, m) + m) + i;
    double s = 0.0;
    for (int c = cmin[m]; c < D.nchunks; ++c) s += partial[(long long)c * 2 * D.nalm + off];
    alm[off] = accumulate ? alm[off] + s : s;
}

// Legendre stage of the synthesis: one warp per (m, chunk); Phi[m nrp + p] = (F_N re, im, F_S re, im)
template <int R>
__global__ void __launch_bounds__(32 * SHT_WARPS, R <= 4 ? 5 : 1) sht_leg_synthesis_kernel(ShtDims D, const double2* __restrict__ alm,
                                                                          const double2* __restrict__ coef,
                                                                          const double* __restrict__ cm,
                                                                          const int* __restrict__ cmin,
                                                                          double4* __restrict__ Phi)
{
    constexpr int C = SHT_C;
    __shared__ double4 sca[SHT_WARPS][C];          // (d_j, Q_{j+1}, Q_j a re, Q_j a im) of one l
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bpm = (D.nchunks + SHT_WARPS - 1) / SHT_WARPS;
    const int m = blockIdx.x / bpm;
    const int chunk = (blockIdx.x % bpm) * SHT_WARPS + wid;
    if (chunk >= D.nchunks) return;
    const int p0 = (chunk * 32 + lane) * R;
    if (chunk < cmin[m]) {                                         // nothing starts here: F = 0
#pragma unroll
        for (int s = 0; s < R; ++s)
            if (p0 + s < D.nrp) Phi[(long long)m * D.nrp + p0 + s] = make_double4(0.0, 0.0, 0.0, 0.0);
        return;
    }
    const long long base = sht_alm_base(D.lmax, m);
    const double2* cf = coef + sht_coef_base(D.lmax, m);
    const double cm_m = cm[m];
    ShtLam q[R];
    double fer[R], fei[R], forr[R], foi[R];
    bool alive = false;
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int p = p0 + s;
        const bool ex = p < D.nrp;
        const ShtRing g = sht_ring(D.nside, ex ? p : 0);
        q[s] = sht_lam_start(D.lmax, m, cm_m, g.z, g.s, ex);
        fer[s] = fei[s] = forr[s] = foi[s] = 0.0;
        alive |= q[s].e == 0;
    }
    // raw loads of the next pass are kept in registers and only combined when they are stored to shared memory, one
    // pass later: nothing depends on them while they are in flight
    double2 nc = make_double2(0.0, 0.0), na = make_double2(0.0, 0.0);
    double nq = 1.0;
    if (lane < C) {
        const int l = m + lane;
        nc = cf[l];
        if (lane) nq = cf[l - 1].y;
        if (l <= D.lmax) na = alm[base + l];
    }
    for (int l0 = m; l0 <= D.lmax; l0 += C) {
        if (lane < C) sca[wid][lane] = make_double4(nc.x, nc.y, nq * na.x, nq * na.y);
        __syncwarp();
        if (lane < C && l0 + C <= D.lmax) {
            const int l = l0 + C + lane;
            nc = cf[l];
            if (lane) nq = cf[l - 1].y;
            na = l <= D.lmax ? alm[base + l] : make_double2(0.0, 0.0);
        }
        if (__any_sync(0xffffffffu, alive)) {
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const double4 c = sca[wid][j];
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    if ((j & 1) == 0) { fer[s] = fma(q[s].lc, c.z, fer[s]); fei[s] = fma(q[s].lc, c.w, fei[s]); }
                    else              { forr[s] = fma(q[s].lc, c.z, forr[s]); foi[s] = fma(q[s].lc, c.w, foi[s]); }
                    sht_mu_advance(q[s], c.x);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const double d = sca[wid][j].x;
#pragma unroll
                for (int s = 0; s < R; ++s) sht_mu_advance(q[s], d);
            }
        }
        {
            const double q15 = sca[wid][C - 2].y, q16 = sca[wid][C - 1].y;
#pragma unroll
            for (int s = 0; s < R; ++s) sht_mu_close(q[s], q15, q16);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < R; ++s)
            if (sht_lam_rescale(q[s])) { fer[s] = fei[s] = forr[s] = foi[s] = 0.0; alive = true; }    // what was summed so far was scaled garbage
    }
#pragma unroll
    for (int s = 0; s < R; ++s) {
        if (p0 + s >= D.nrp) continue;
        double4 t = make_double4(0.0, 0.0, 0.0, 0.0);
        if (q[s].e == 0) t = make_double4(fer[s] + forr[s], fei[s] + foi[s], fer[s] - forr[s], fei[s] - foi[s]);
        Phi[(long long)m * D.nrp + p0 + s] = t;
    }
}

// cl[l] = (Re(a_l0 conj b_l0) + 2 sum_{m=1..l} Re(a_lm conj b_lm))/(2l+1).  Two stages so that the m sum is not one
// serial chain per l: blockIdx.y = segment of SHT_CL_SEG consecutive m, partial[seg][l]; then the segments in order.
constexpr int SHT_CL_SEG = 64;
__global__ void sht_alm2cl_partial_kernel(int lmax, const double2* __restrict__ a, const double2* __restrict__ b,
                                          double* __restrict__ partial)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > lmax) return;
    const int m0 = blockIdx.y * SHT_CL_SEG;
    double s = 0.0;
    for (int m = m0; m < m0 + SHT_CL_SEG && m <= l; ++m) {
        const long long i = sht_alm_base(lmax, m) + l;
        const double2 x = a[i], y = b[i];
        const double t = x.x * y.x + x.y * y.y;
        s += m ? 2.0 * t : t;
    }
    partial[(long long)blockIdx.y * (lmax + 1) + l] = s;
}

__global__ void sht_alm2cl_finish_kernel(int lmax, int nseg, const double* __restrict__ partial, double* __restrict__ cl)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > lmax) return;
    double s = 0.0;
    for (int g = 0; g < nseg && g * SHT_CL_SEG <= l; ++g) s += partial[(long long)g * (lmax + 1) + l];
    cl[l] = s / (2.0 * (double)l + 1.0);
}

// out = scale * a [* b [* c]]   (effective_weight_alm!: mask_i .* mask_j .* sigma^2 .* Omega_pix)
__global__ void sht_product_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b,
                                   const double* __restrict__ c, double scale, double* __restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = a[i];
        if (b) v *= b[i];
        if (c) v *= c[i] * scale; else v *= scale;
        out[i] = v;
    }
}

#endif  // __CUDACC__

}  // namespace psb
