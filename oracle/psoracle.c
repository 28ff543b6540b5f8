/*
 * psoracle.c -- CPU oracle for the PowerSpectra.jl Wigner-3j hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see psoracle_impl.h for the full header and the
 * reference file:line map).  The product (libpsb200.so) never links, loads or
 * calls this file; it exists so tests/, smoke() and bench.py's CPU arm have a
 * CPU restatement of src/modecoupling.jl:3-159 and src/covariance.jl:92-446 to
 * compare with and to time.
 *
 * PARITY STATUS: "parity unpinned" against runnable reference fixtures at the
 * mask level -- every mask/map FITS file and every .npy/.jld2 golden was stripped
 * from /root/reference (.MISSING_LARGE_BLOBS) and Julia is not installed.  What
 * pins it instead (tests/test_oracle.py):
 *   - exact Wigner 3j values from sympy for both families, all pairs l<=24 + spot
 *     checks at higher l (abs 1e-14);
 *   - mpmath 40-digit recurrences at l up to 6143;
 *   - the NaMaster diagonals the reference's own tests hold
 *     (test/data/mcm_TT_diag.txt, mcm_EE_diag.txt, mcm_TE_diag.txt, used at
 *     test/test_mcm.jl:12-50): the even-l mask spectrum is solved from the TT
 *     diagonal and must then reproduce the EE and TE diagonals;
 *   - analytic identities (full-sky mask => identity, completeness, covariance <-> MCM).
 *
 * Two instantiations: double ("pso_*", the reference-shaped Float64 path, also
 * the timed CPU baseline) and long double ("pso_*_ld", used to attribute error).
 */
#include <math.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* 0: the reference computation.  1: every l3 term and every additive term of a covariance
 * formula enters with its absolute value -- the "condition sum" S_abs that scales the
 * Float64 noise floor of a cancelling entry in the parity tests.  Never used when timing. */
static int pso_abs_mode = 0;
void pso_set_abs_mode(int on) { pso_abs_mode = on ? 1 : 0; }

#define PSO_CAT_(a, b) a##b
#define PSO_CAT(a, b) PSO_CAT_(a, b)

/* ---- double ---- */
#define REAL double
#define SQRT sqrt
#define FABS fabs
#define PI_R 3.14159265358979323846
#define FN(x) PSO_CAT(pso_, x)
#include "psoracle_impl.h"
#undef REAL
#undef SQRT
#undef FABS
#undef PI_R
#undef FN

/* ---- long double ---- */
#define REAL long double
#define SQRT sqrtl
#define FABS fabsl
#define PI_R 3.14159265358979323846264338327950288L
#define FN(x) PSO_CAT(PSO_CAT(pso_, x), _ld)
#include "psoracle_impl.h"
#undef REAL
#undef SQRT
#undef FABS
#undef PI_R
#undef FN

int pso_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void pso_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
