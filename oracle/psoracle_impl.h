/*
 * psoracle_impl.h -- body of the CPU oracle, compiled once per real type.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / the timed CPU arm.
 *
 * Included by psoracle.c with
 *     REAL   = double        FN(x) = pso_##x        (reference-shaped Float64 path)
 *     REAL   = long double   FN(x) = pso_##x##_ld   (80-bit path used to attribute error)
 *
 * What it restates (all citations relative to /root/reference):
 *   - Wigner-3j family evaluation done by the un-vendored dependency WignerFamilies.jl
 *     (Project.toml:25,43; compat "1", no Manifest => no exact pin).  Call sites:
 *     src/modecoupling.jl:70-73, src/covariance.jl:105-108,166-169,220-223,276-283,
 *     350-357,387-390,433-436.  Published algorithm: Schulten & Gordon (1975) three-term
 *     recurrence in j, evaluated the Luscombe & Luban (1998) way: ratio recursions
 *     inward from both ends through the non-classical regions, the three-term
 *     recurrence across the classical region from both sides, matched in the middle,
 *     normalised with sum_j (2j+1) f(j)^2 = 1 and signed with
 *     sgn f(j_max) = (-1)^(j2-j3-m1).  The m2=m3=0 family (B==0, ratios undefined)
 *     takes the dedicated two-step path, every other-parity entry exactly 0.
 *   - Xi_TT / Xi_EE / Xi_EB / Xi_TE          src/modecoupling.jl:3-66
 *   - inner_mcm00!/02!/++!/--!               src/modecoupling.jl:78-159
 *   - loop_covTTTT!/EEEE!/TTTE!/TETE!/TEEE!/TEEE_planck!/TTEE!
 *                                            src/covariance.jl:92-446
 *   Loop shape follows the reference: one task per l1 row scheduled dynamically
 *   (@qthreads, src/modecoupling.jl:84), a per-thread buffer of 2*lmax+1 reals
 *   (src/modecoupling.jl:82), the family materialised, squared / multiplied in
 *   place, then one Xi dot product per window spectrum.
 */

/* ---- recurrence coefficients ------------------------------------------------------
 * j A(j+1) f(j+1) + B(j) f(j) + (j+1) A(j) f(j-1) = 0,
 *   A(j) = sqrt[(j^2-(j2-j3)^2)((j2+j3+1)^2-j^2)(j^2-m1^2)],  m1 = -m2-m3
 *   B(j) = -(2j+1)[j2(j2+1)m1 - j3(j3+1)m1 - j(j+1)(m3-m2)]
 * For m1 == 0 (the only case on the PowerSpectra hot path) A(j) = j a(j) and
 * B(j) = (2j+1) j (j+1) (m3-m2); dividing the relation by j(j+1) gives the reduced
 *   a(j+1) f(j+1) + (2j+1)(m3-m2) f(j) + a(j) f(j-1) = 0,
 * which is also well defined at j = 0 (l1 == l2), where the unreduced one reads 0 = 0.
 */
typedef struct {
    int j2, j3, m2, m3, m1;
    int nmin, nmax;
    const REAL* root;   /* root[j - nmin] = sqrt[(j^2-(j2-j3)^2)((j2+j3+1)^2-j^2)(j^2-m1^2 if m1 != 0)], j = nmin..nmax+1 */
} FN(fam_t);

/* The square root every coefficient needs is evaluated ONCE per j (X(j) and Z(j+1) share it), into a
 * per-thread scratch, before the recurrences run. */
static void FN(fill_roots)(const FN(fam_t)* w, REAL* root)
{
    const REAL d = (REAL)(w->j2 - w->j3), s = (REAL)(w->j2 + w->j3 + 1), m = (REAL)w->m1;
    for (int j = w->nmin; j <= w->nmax + 1; ++j) {
        const REAL jj = (REAL)j;
        REAL a2 = (jj * jj - d * d) * (s * s - jj * jj);
        if (w->m1 != 0) a2 *= (jj * jj - m * m);
        if (a2 < 0) a2 = 0;
        root[j - w->nmin] = SQRT(a2);
    }
}
static inline REAL FN(Xc)(const FN(fam_t)* w, int j)   /* coefficient of f(j+1) */
{
    const REAL r = w->root[j + 1 - w->nmin];
    return w->m1 == 0 ? r : (REAL)j * r;
}
static inline REAL FN(Zc)(const FN(fam_t)* w, int j)   /* coefficient of f(j-1) */
{
    const REAL r = w->root[j - w->nmin];
    return w->m1 == 0 ? r : (REAL)(j + 1) * r;
}
static inline REAL FN(Yc)(const FN(fam_t)* w, int j)   /* coefficient of f(j) */
{
    REAL jj = (REAL)j;
    if (w->m1 == 0) return (2 * jj + 1) * (REAL)(w->m3 - w->m2);
    return -(2 * jj + 1) * ((REAL)w->j2 * (w->j2 + 1) * w->m1 - (REAL)w->j3 * (w->j3 + 1) * w->m1
                            - jj * (jj + 1) * (REAL)(w->m3 - w->m2));
}

/* Whole family f(j) = (j j2 j3; -m2-m3 m2 m3), j = nmin..nmax, into out[0..n-1].
 * Returns n (0 if the family is empty). */
static int FN(family)(int j2, int j3, int m2, int m3, REAL* out, int* pnmin, int* pnmax, REAL* scratch)
{
    FN(fam_t) w;
    w.j2 = j2; w.j3 = j3; w.m2 = m2; w.m3 = m3; w.m1 = -m2 - m3;
    int dm = abs(j2 - j3), am = abs(m2 + m3);
    w.nmin = dm > am ? dm : am;
    w.nmax = j2 + j3;
    if (pnmin) *pnmin = w.nmin;
    if (pnmax) *pnmax = w.nmax;
    int n = w.nmax - w.nmin + 1;
    if (n <= 0) return 0;
    const int nmin = w.nmin, nmax = w.nmax;
    FN(fill_roots)(&w, scratch);
    w.root = scratch;
#define PSI(j) out[(j) - nmin]

    if (n == 1) {
        PSI(nmin) = 1;
    } else if (m2 == 0 && m3 == 0) {
        /* B == 0: f(j+1) = -[Z(j)/X(j)] f(j-1); entries with j2+j3+j odd are exactly 0. */
        PSI(nmin) = 1;
        PSI(nmin + 1) = 0;
        for (int j = nmin + 1; j < nmax; ++j)
            PSI(j + 1) = -(FN(Zc)(&w, j) / FN(Xc)(&w, j)) * PSI(j - 1);
    } else {
        /* --- non-classical region at the top: r(j) = f(j)/f(j-1), downward from nmax.
         * Ratios are parked in out[] on [nplus, nmax]; stop at the first |r| >= 1. --- */
        int nplus = nmax;
        {
            int j = nmax;
            REAL r = -FN(Zc)(&w, j) / FN(Yc)(&w, j);
            PSI(j) = r;
            while (FABS(r) < 1 && j - 1 > nmin) {
                --j;
                r = -FN(Zc)(&w, j) / (FN(Yc)(&w, j) + FN(Xc)(&w, j) * r);
                PSI(j) = r;
            }
            nplus = j;
        }
        /* --- non-classical region at the bottom: s(j) = f(j)/f(j+1), upward from nmin.
         * Ratios parked on [nmin, nminus], nminus < nplus. --- */
        int nminus = nmin;
        {
            int j = nmin;
            REAL s = -FN(Xc)(&w, j) / FN(Yc)(&w, j);
            PSI(j) = s;
            while (FABS(s) < 1 && j + 1 < nplus) {
                ++j;
                s = -FN(Xc)(&w, j) / (FN(Yc)(&w, j) + FN(Zc)(&w, j) * s);
                PSI(j) = s;
            }
            nminus = j;
        }
        /* --- classical region: lower branch L (L(nminus) = 1) marched up to nc+1, upper
         * branch U (U(nplus) = 1) marched down to nc; they overlap on {nc, nc+1}. --- */
        const int nc = (nminus + nplus) / 2;
        const REAL s_at = PSI(nminus), r_at = PSI(nplus);
        PSI(nminus) = 1;
        for (int j = nminus - 1; j >= nmin; --j) PSI(j) = PSI(j) * PSI(j + 1);
        REAL Lc, Lc1;
        {
            REAL fm = 1, f0 = 1 / s_at;              /* L(nminus), L(nminus+1) */
            int j = nminus + 1;                       /* f0 = L(j) */
            while (j <= nc) {
                PSI(j) = f0;
                REAL fp = -(FN(Yc)(&w, j) * f0 + FN(Zc)(&w, j) * fm) / FN(Xc)(&w, j);
                fm = f0; f0 = fp; ++j;
            }
            /* now j == nc+1, f0 = L(nc+1), fm = L(nc) */
            Lc = fm; Lc1 = f0;
        }
        PSI(nplus) = 1;
        for (int j = nplus + 1; j <= nmax; ++j) PSI(j) = PSI(j) * PSI(j - 1);
        REAL Uc, Uc1;
        {
            REAL gp = 1, g0 = 1 / r_at;              /* U(nplus), U(nplus-1) */
            int j = nplus - 1;                        /* g0 = U(j) */
            while (j >= nc + 1) {
                PSI(j) = g0;
                REAL gm = -(FN(Xc)(&w, j) * gp + FN(Yc)(&w, j) * g0) / FN(Zc)(&w, j);
                gp = g0; g0 = gm; --j;
            }
            /* now j == nc, g0 = U(nc), gp = U(nc+1) */
            Uc = g0; Uc1 = gp;
        }
        /* least-squares scale of the lower branch onto the upper one over the overlap */
        const REAL lam = (Uc * Lc + Uc1 * Lc1) / (Lc * Lc + Lc1 * Lc1);
        for (int j = nmin; j <= nc; ++j) PSI(j) *= lam;
    }
    /* normalise: sum (2j+1) f^2 = 1; sign: f(nmax) (-1)^(j2-j3-m1) > 0 */
    REAL norm = 0;
    for (int j = nmin; j <= nmax; ++j) norm += (REAL)(2 * j + 1) * PSI(j) * PSI(j);
    REAL sc = 1 / SQRT(norm);
    int neg = ((j2 - j3 - w.m1) % 2) != 0;
    if ((PSI(nmax) < 0) != neg) sc = -sc;
    for (int j = nmin; j <= nmax; ++j) PSI(j) *= sc;
#undef PSI
    return n;
}

/* ---- Xi projectors, src/modecoupling.jl:3-66 ---------------------------------------
 * w = buffer of the (already squared / multiplied) family on [nmin, nmax];
 * W[0..lenW-1] is 0-based in l3.  par: 0 = every l3 (Xi_TT), 1 = l1+l2+l3 even
 * (Xi_EE, Xi_TE), 2 = l1+l2+l3 odd (Xi_EB). */
static inline REAL FN(xi)(const double* W, int lenW, const REAL* w, int nmin, int nmax,
                          int l1, int l2, int par)
{
    int s = nmin > 0 ? nmin : 0;
    int e = nmax < lenW - 1 ? nmax : lenW - 1;
    int step = 1;
    if (par == 1) { if ((l1 + l2 + s) & 1) ++s; step = 2; }
    if (par == 2) { if (!((l1 + l2 + s) & 1)) ++s; step = 2; }
    REAL acc = 0;
    if (pso_abs_mode) {       /* condition sum: sum |term|, used only to scale test tolerances */
        for (int l3 = s; l3 <= e; l3 += step)
            acc += FABS((REAL)(2 * l3 + 1) * w[l3 - nmin] * (REAL)W[l3]);
    } else {
        for (int l3 = s; l3 <= e; l3 += step)
            acc += (REAL)(2 * l3 + 1) * w[l3 - nmin] * (REAL)W[l3];
    }
    return acc / (4 * PI_R);
}

/* kinds: 0 = M00 (TT), 1 = M02 (TE/ET/TB/BT), 2 = M++, 3 = M--  (src/modecoupling.jl:78-159)
 * Only rows l1 = lmin + row0 + k*rstep are computed (rstep = 1: all rows), so a
 * deterministic sample of rows can be timed.  M is column-major, M[(l1-lmin)+(l2-lmin)*ld].
 * Returns the number of 3j terms evaluated (full families, as the reference does). */
long long FN(mcm_rows)(int kind, int lmin, int lmax, const double* V, int nV, double* M, long ld,
                       const int* rows, int nrows);
long long FN(mcm)(int kind, int lmin, int lmax, const double* V, int nV, double* M, long ld,
                  int row0, int rstep)
{
    if (kind < 0 || kind > 3 || lmin < 0 || lmax < lmin || nV < 1 || rstep < 1 || row0 < 0) return -1;
    int n = 0;
    for (int l1 = lmin + row0; l1 <= lmax; l1 += rstep) ++n;
    int* rows = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    n = 0;
    for (int l1 = lmin + row0; l1 <= lmax; l1 += rstep) rows[n++] = l1;
    long long t = FN(mcm_rows)(kind, lmin, lmax, V, nV, M, ld, rows, n);
    free(rows);
    return t;
}
/* Same for an explicit list of rows l1 (absolute multipoles, lmin <= l1 <= lmax, any order). */
long long FN(mcm_rows)(int kind, int lmin, int lmax, const double* V, int nV, double* M, long ld,
                       const int* rows, int nrows)
{
    if (kind < 0 || kind > 3 || lmin < 0 || lmax < lmin || nV < 1 || nrows < 0 || (nrows && !rows)) return -1;
    for (int i = 0; i < nrows; ++i) if (rows[i] < lmin || rows[i] > lmax) return -1;
    long long terms = 0;
    int nbuf = 2 * lmax + 1;
#pragma omp parallel reduction(+ : terms)
    {
        REAL* b0 = (REAL*)malloc(sizeof(REAL) * nbuf);
        REAL* b2 = (REAL*)malloc(sizeof(REAL) * nbuf);
        REAL* rootbuf = (REAL*)malloc(sizeof(REAL) * (nbuf + 2));
#pragma omp for schedule(dynamic, 1)
        for (int ir = 0; ir < nrows; ++ir) {
            const int l1 = rows[ir];
            for (int l2 = l1; l2 <= lmax; ++l2) {
                int nmin, nmax, n;
                REAL xi;
                if (kind == 0) {
                    n = FN(family)(l1, l2, 0, 0, b0, &nmin, &nmax, rootbuf);
                    for (int i = 0; i < n; ++i) b0[i] = b0[i] * b0[i];
                    xi = FN(xi)(V, nV, b0, nmin, nmax, l1, l2, 0);
                    terms += n;
                } else if (kind == 1) {
                    n = FN(family)(l1, l2, 0, 0, b0, &nmin, &nmax, rootbuf);
                    FN(family)(l1, l2, -2, 2, b2, &nmin, &nmax, rootbuf);
                    for (int i = 0; i < n; ++i) b0[i] *= b2[i];
                    xi = FN(xi)(V, nV, b0, nmin, nmax, l1, l2, 1);
                    terms += 2 * n;
                } else {
                    n = FN(family)(l1, l2, -2, 2, b2, &nmin, &nmax, rootbuf);
                    for (int i = 0; i < n; ++i) b2[i] = b2[i] * b2[i];
                    xi = FN(xi)(V, nV, b2, nmin, nmax, l1, l2, kind == 2 ? 1 : 2);
                    terms += n;
                }
                M[(long)(l1 - lmin) + (long)(l2 - lmin) * ld] = (double)((REAL)(2 * l2 + 1) * xi);
                M[(long)(l2 - lmin) + (long)(l1 - lmin) * ld] = (double)((REAL)(2 * l1 + 1) * xi);
            }
        }
        free(b0); free(b2); free(rootbuf);
    }
    return terms;
}

/* blocks: 0 TTTT, 1 EEEE, 2 TTTE, 3 TETE, 4 TEEE_planck, 5 TEEE, 6 TTEE
 * (src/covariance.jl:92-122,153-183,208-235,261-302,376-402,337-372,422-446).
 * sp / rt / W follow the positional order of the reference signatures; every vector
 * is 0-based in l.  Returns the number of 3j terms evaluated, -1 on bad arguments. */
long long FN(cov_rows)(int block, int lmin, int lmax, const double* const* sp, int nsp,
                       const double* const* rt, int nrt, const double* const* W, int nW, int lenW,
                       double* C, long ld, const int* rows, int nrows);
long long FN(cov)(int block, int lmin, int lmax, const double* const* sp, int nsp,
                  const double* const* rt, int nrt, const double* const* W, int nW, int lenW,
                  double* C, long ld, int row0, int rstep)
{
    if (lmin < 0 || lmax < lmin || rstep < 1 || row0 < 0) return -1;
    int n = 0;
    for (int l1 = lmin + row0; l1 <= lmax; l1 += rstep) ++n;
    int* rows = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    n = 0;
    for (int l1 = lmin + row0; l1 <= lmax; l1 += rstep) rows[n++] = l1;
    long long t = FN(cov_rows)(block, lmin, lmax, sp, nsp, rt, nrt, W, nW, lenW, C, ld, rows, n);
    free(rows);
    return t;
}
long long FN(cov_rows)(int block, int lmin, int lmax, const double* const* sp, int nsp,
                       const double* const* rt, int nrt, const double* const* W, int nW, int lenW,
                       double* C, long ld, const int* rows, int nrows)
{
    static const int need_sp[7] = {4, 4, 4, 4, 4, 4, 4};
    static const int need_rt[7] = {4, 4, 2, 2, 2, 2, 0};
    static const int need_W[7] = {8, 8, 4, 5, 4, 4, 2};
    if (block < 0 || block > 6 || lmin < 0 || lmax < lmin || lenW < 1 || nrows < 0 || (nrows && !rows)) return -1;
    if (nsp != need_sp[block] || nrt != need_rt[block] || nW != need_W[block]) return -1;
    for (int i = 0; i < nrows; ++i) if (rows[i] < lmin || rows[i] > lmax) return -1;
    long long terms = 0;
    int nbuf = 2 * lmax + 1;
#pragma omp parallel reduction(+ : terms)
    {
        REAL* b0 = (REAL*)malloc(sizeof(REAL) * nbuf);
        REAL* b2 = (REAL*)malloc(sizeof(REAL) * nbuf);
        REAL* rootbuf = (REAL*)malloc(sizeof(REAL) * (nbuf + 2));
#pragma omp for schedule(dynamic, 1)
        for (int ir = 0; ir < nrows; ++ir) {
            const int l1 = rows[ir];
            for (int l2 = l1; l2 <= lmax; ++l2) {
                int nmin, nmax, n;
                REAL c = 0;
#define S(k, l) ((REAL)sp[k][l])
#define AT(x) (pso_abs_mode ? FABS(x) : (x))   /* each additive term of the block formula */
#define R(k, l) ((REAL)rt[k][l])
                if (block == 0 || block == 1) {
                    /* spectra: ip, jq, iq, jp ; ratios: ip, jq, iq, jp */
                    int par = block == 0 ? 0 : 1;
                    n = block == 0 ? FN(family)(l1, l2, 0, 0, b0, &nmin, &nmax, rootbuf)
                                   : FN(family)(l1, l2, -2, 2, b0, &nmin, &nmax, rootbuf);
                    for (int i = 0; i < n; ++i) b0[i] = b0[i] * b0[i];
                    terms += n;
                    REAL x[8];
                    for (int k = 0; k < 8; ++k) x[k] = FN(xi)(W[k], lenW, b0, nmin, nmax, l1, l2, par);
                    c = AT(SQRT(S(0, l1) * S(0, l2) * S(1, l1) * S(1, l2)) * x[0]) +
                        AT(SQRT(S(2, l1) * S(2, l2) * S(3, l1) * S(3, l2)) * x[1]) +
                        AT(SQRT(S(0, l1) * S(0, l2)) * x[2] * R(1, l1) * R(1, l2)) +
                        AT(SQRT(S(1, l1) * S(1, l2)) * x[3] * R(0, l1) * R(0, l2)) +
                        AT(SQRT(S(2, l1) * S(2, l2)) * x[4] * R(3, l1) * R(3, l2)) +
                        AT(SQRT(S(3, l1) * S(3, l2)) * x[5] * R(2, l1) * R(2, l2)) +
                        AT(x[6] * R(0, l1) * R(1, l1) * R(0, l2) * R(1, l2)) +
                        AT(x[7] * R(2, l1) * R(3, l1) * R(2, l2) * R(3, l2));
                } else if (block == 2) {
                    /* TTTE: spectra TTip, TTjp, TEiq, TEjq ; ratios ip, jp */
                    n = FN(family)(l1, l2, 0, 0, b0, &nmin, &nmax, rootbuf);
                    for (int i = 0; i < n; ++i) b0[i] = b0[i] * b0[i];
                    terms += n;
                    REAL x[4];
                    for (int k = 0; k < 4; ++k) x[k] = FN(xi)(W[k], lenW, b0, nmin, nmax, l1, l2, 0);
                    c = (AT(SQRT(S(0, l1) * S(0, l2)) * (S(3, l1) + S(3, l2)) * x[0]) +
                         AT(SQRT(S(1, l1) * S(1, l2)) * (S(2, l1) + S(2, l2)) * x[1]) +
                         AT((S(3, l1) + S(3, l2)) * x[2] * R(0, l1) * R(0, l2)) +
                         AT((S(2, l1) + S(2, l2)) * x[3] * R(1, l1) * R(1, l2))) / 2;
                } else if (block == 3) {
                    /* TETE: spectra TTip, EEjq, TEiq, TEjp ; ratios TT_ip, PP_jq */
                    n = FN(family)(l1, l2, 0, 0, b0, &nmin, &nmax, rootbuf);
                    FN(family)(l1, l2, -2, 2, b2, &nmin, &nmax, rootbuf);
                    for (int i = 0; i < n; ++i) { b2[i] *= b0[i]; b0[i] *= b0[i]; }
                    terms += 2 * n;
                    REAL x1 = FN(xi)(W[0], lenW, b2, nmin, nmax, l1, l2, 1);
                    REAL x2 = FN(xi)(W[1], lenW, b0, nmin, nmax, l1, l2, 0);
                    REAL x3 = FN(xi)(W[2], lenW, b2, nmin, nmax, l1, l2, 1);
                    REAL x4 = FN(xi)(W[3], lenW, b2, nmin, nmax, l1, l2, 1);
                    REAL x5 = FN(xi)(W[4], lenW, b2, nmin, nmax, l1, l2, 1);
                    c = AT(SQRT(S(0, l1) * S(0, l2) * S(1, l1) * S(1, l2)) * x1) +
                        AT((REAL)0.5 * (S(2, l1) * S(3, l2) + S(3, l1) * S(2, l2)) * x2) +
                        AT(SQRT(S(0, l1) * S(0, l2)) * x3 * R(1, l1) * R(1, l2)) +
                        AT(SQRT(S(1, l1) * S(1, l2)) * x4 * R(0, l1) * R(0, l2)) +
                        AT(x5 * R(0, l1) * R(0, l2) * R(1, l1) * R(1, l2));
                } else if (block == 4 || block == 5) {
                    /* TEEE: spectra EEjq, EEjp, TEip, TEiq ; ratios EE_jq, EE_jp */
                    if (block == 4) {
                        n = FN(family)(l1, l2, -2, 2, b2, &nmin, &nmax, rootbuf);
                        for (int i = 0; i < n; ++i) b2[i] = b2[i] * b2[i];
                        terms += n;
                    } else {
                        n = FN(family)(l1, l2, 0, 0, b0, &nmin, &nmax, rootbuf);
                        FN(family)(l1, l2, -2, 2, b2, &nmin, &nmax, rootbuf);
                        for (int i = 0; i < n; ++i) b2[i] *= b0[i];
                        terms += 2 * n;
                    }
                    REAL x[4];
                    for (int k = 0; k < 4; ++k) x[k] = FN(xi)(W[k], lenW, b2, nmin, nmax, l1, l2, 1);
                    c = (AT(SQRT(S(0, l1) * S(0, l2)) * (S(2, l1) + S(2, l2)) * x[0]) +
                         AT(SQRT(S(1, l1) * S(1, l2)) * (S(3, l1) + S(3, l2)) * x[1]) +
                         AT((S(2, l1) + S(2, l2)) * x[2] * R(0, l1) * R(0, l2)) +
                         AT((S(3, l1) + S(3, l2)) * x[3] * R(1, l1) * R(1, l2))) / 2;
                } else {
                    /* TTEE: spectra TEip, TEiq, TEjq, TEjp */
                    n = FN(family)(l1, l2, 0, 0, b0, &nmin, &nmax, rootbuf);
                    for (int i = 0; i < n; ++i) b0[i] = b0[i] * b0[i];
                    terms += n;
                    REAL x1 = FN(xi)(W[0], lenW, b0, nmin, nmax, l1, l2, 0);
                    REAL x2 = FN(xi)(W[1], lenW, b0, nmin, nmax, l1, l2, 0);
                    c = (AT((S(0, l1) * S(2, l2) + S(2, l1) * S(0, l2)) * x1) +
                         AT((S(1, l1) * S(3, l2) + S(3, l1) * S(1, l2)) * x2)) / 2;
                }
#undef S
#undef AT
#undef R
                C[(long)(l1 - lmin) + (long)(l2 - lmin) * ld] = (double)c;
                C[(long)(l2 - lmin) + (long)(l1 - lmin) * ld] = (double)c;
            }
        }
        free(b0); free(b2); free(rootbuf);
    }
    return terms;
}

/* one family into a double buffer (for the known-answer tests against exact 3j) */
int FN(w3j_family)(int j2, int j3, int m2, int m3, double* out, int nout, int* nmin, int* nmax)
{
    int lo = abs(j2 - j3), am = abs(m2 + m3);
    if (am > lo) lo = am;
    int n = j2 + j3 - lo + 1;
    if (n <= 0) { if (nmin) *nmin = lo; if (nmax) *nmax = j2 + j3; return 0; }
    if (nout < n) return -1;
    REAL* b = (REAL*)malloc(sizeof(REAL) * n);
    REAL* rootbuf = (REAL*)malloc(sizeof(REAL) * (n + 2));
    FN(family)(j2, j3, m2, m3, b, nmin, nmax, rootbuf);
    free(rootbuf);
    for (int i = 0; i < n; ++i) out[i] = (double)b[i];
    free(b);
    return n;
}

/* ---- QuickPol Xi matrix, src/beam.jl:72-101 (quickpolXi!) and :16-28 (Xisum) ----------
 * For l'' = 2..lmax (one task per row, src/beam.jl:81) and l over the stored band of that
 * row (specrowrange, src/beam.jl:59-63: l = max(2, l''-band_lo) .. min(lmax, l''+band_hi)):
 *   wF1 = WignerF(l, l'', -s1, -nu1), wF2 = WignerF(l, l'', -s2, -nu2)      (:86-87)
 *   Xi[l'', l] = sgn * sum_{l' = max(first1,first2)}^{min(last1,last2)} W[l'] w3j1[l'] w3j2[l']
 *   sgn = (-1)^(s1+s2+nu1+nu2)                                                (:98-99)
 * No (2l'+1) and no 1/4pi here: the caller's W (quickpolW, :43-56) carries the weights.
 * Xb is BandedMatrices' own storage of parent(Xi): Xb[(band_hi + l'' - l) + l*ldb]
 * (column l of the band, ldb >= band_lo+band_hi+1).  Entries the reference loop does not
 * visit (rows/columns < 2) are not written.
 * Two places where the reference leaves behaviour undefined and this restatement decides:
 *   - l' > lenW-1: the reference indexes W under @inbounds (out of bounds unless the scan
 *     weights have lmax >= 2*lmax of Xi); here those terms are dropped (W = 0 there);
 *   - |s| > l or |nu| > l'' (projection larger than the angular momentum): the true symbol
 *     is 0 and the entry is stored as 0.
 * abs_mode: sum of |terms| (condition sum for the test tolerance).
 * Returns the number of 3j terms evaluated (both families, full length). */
long long FN(quickpol_xi)(int nu1, int nu2, int s1, int s2, int lmax, const double* W, int lenW,
                          int band_lo, int band_hi, double* Xb, long ldb)
{
    if (lmax < 0 || lenW < 1 || band_lo < 0 || band_hi < 0 || ldb < (long)band_lo + band_hi + 1 || !W || !Xb)
        return -1;
    long long terms = 0;
    const int nbuf = 2 * lmax + 1;
    const REAL sgn = ((s1 + s2 + nu1 + nu2) % 2) ? -1 : 1;
#pragma omp parallel reduction(+ : terms)
    {
        REAL* b1 = (REAL*)malloc(sizeof(REAL) * (nbuf > 0 ? nbuf : 1));
        REAL* b2 = (REAL*)malloc(sizeof(REAL) * (nbuf > 0 ? nbuf : 1));
        REAL* rootbuf = (REAL*)malloc(sizeof(REAL) * (nbuf + 2));
#pragma omp for schedule(dynamic, 1)
        for (int lpp = 2; lpp <= lmax; ++lpp) {
            int lo = lpp - band_lo, hi = lpp + band_hi;
            if (lo < 2) lo = 2;
            if (hi > lmax) hi = lmax;
            for (int l = lo; l <= hi; ++l) {
                REAL acc = 0;
                if (abs(s1) <= l && abs(s2) <= l && abs(nu1) <= lpp && abs(nu2) <= lpp) {
                    int min1, max1, min2, max2;
                    int n1 = FN(family)(l, lpp, -s1, -nu1, b1, &min1, &max1, rootbuf);
                    int n2 = FN(family)(l, lpp, -s2, -nu2, b2, &min2, &max2, rootbuf);
                    terms += n1 + n2;
                    if (n1 > 0 && n2 > 0) {
                        int a = min1 > min2 ? min1 : min2;
                        int e = max1 < max2 ? max1 : max2;
                        if (e > lenW - 1) e = lenW - 1;
                        for (int lp = a; lp <= e; ++lp) {
                            REAL t = (REAL)W[lp] * b1[lp - min1] * b2[lp - min2];
                            acc += pso_abs_mode ? FABS(t) : t;
                        }
                    }
                }
                Xb[(long)(band_hi + lpp - l) + (long)l * ldb] = (double)(pso_abs_mode ? acc : sgn * acc);
            }
        }
        free(b1); free(b2); free(rootbuf);
    }
    return terms;
}
