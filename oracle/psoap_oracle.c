/*
 * psoap_oracle.c — CPU restatement of PSOAP's GP log-likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (psoap_b200/) may import, link or call this
 * file; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA
 * path against an independent scalar implementation.
 *
 * Parity status: PINNED.  The reference's own tests hold no vector for this path, so the
 * oracle is pinned against outputs of the reference itself: tests/golden/*.npz were produced by importing
 * the unmodified Python/Cython reference (tests/golden/make_golden.py) and tests/test_oracle.py checks
 * every function below against them (bit-exact for the fills, <=1e-12 rel for lnlike).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no fast-math, no FMA contraction, like the reference's
 * `gcc -O2` Cython build) -> oracle/_build/libpsoap_oracle.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* psoap/matrix_functions.pyx:16-17 */
static const double c_kms = 2.99792458e5;
#define C_KMS2 (c_kms * c_kms)

/* psoap/matrix_functions.pyx:21-57 (ncomp=1), :101-144 (ncomp=2), :151-201 (ncomp=3).
 * mat is row-major [N, ld]; both triangles and the diagonal are written.  lwl[c] is the ln-wavelength
 * vector of component c.  Evaluation order of the exponent argument is (p2*r)*r as in the .pyx. */
void oracle_fill_V11(double *mat, long ld, int N, int ncomp, const double *lwl_f, const double *lwl_g,
                     const double *lwl_h, const double *amp, const double *l)
{
    const double *lwl[3] = {lwl_f, lwl_g, lwl_h};
    double amp2[3], p2[3];
    for (int c = 0; c < ncomp; ++c) {
        amp2[c] = amp[c] * amp[c];             /* pyx:28, :108, :111 */
        p2[c] = -0.5 * C_KMS2 / (l[c] * l[c]); /* pyx:29, :109, :112 */
    }
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < i; ++j) {
            double cov = 0.0;
            for (int c = 0; c < ncomp; ++c) {
                double r = lwl[c][j] - lwl[c][i]; /* pyx:47, :133-134 */
                double e = amp2[c] * exp(p2[c] * r * r);
                cov = (c == 0) ? e : cov + e;     /* pyx:49, :136, :193 (left-to-right sum) */
            }
            mat[(long)i * ld + j] = cov;          /* pyx:52-53 */
            mat[(long)j * ld + i] = cov;
        }
    }
    for (int i = 0; i < N; ++i) {                 /* pyx:56-57, :143-144, :200-201 */
        double d = amp2[0];
        for (int c = 1; c < ncomp; ++c) d += amp2[c];
        mat[(long)i * ld + i] = d;
    }
}

/* psoap/matrix_functions.pyx:63-94.  mat is row-major [M, ld], M = len(lwl_f) rows, N = len(lwl_predict)
 * columns; no diagonal special case. */
void oracle_fill_V12_f(double *mat, long ld, int M, int N, const double *lwl_f, const double *lwl_predict,
                       double amp_f, double l_f)
{
    double amp2f = amp_f * amp_f;
    double p2f = -0.5 * C_KMS2 / (l_f * l_f);
    for (int i = 0; i < M; ++i) {
        double lwl_f0 = lwl_f[i];
        for (int j = 0; j < N; ++j) {
            double rf = lwl_predict[j] - lwl_f0;
            mat[(long)i * ld + j] = amp2f * exp(p2f * rf * rf);
        }
    }
}

/* psoap/data.py:25-38 (lredshift) and :40-63 (replicate_wls): out[c][k] = lwl[k] + (-v[c][epoch[k]])/c_kms
 * where epoch[k] is the epoch of masked, row-major-flattened pixel k. */
void oracle_replicate_wls(double *out, const double *lwl, const int *epoch, long N, const double *vel,
                          int ncomp, int n_epochs)
{
    for (int c = 0; c < ncomp; ++c)
        for (long k = 0; k < N; ++k)
            out[(long)c * N + k] = lwl[k] + (-vel[(long)c * n_epochs + epoch[k]]) / c_kms;
}

/* LAPACK dpotrf, UPLO='U', as reached through scipy.linalg.cho_factor(lower=False)
 * (psoap/covariance.py:325,348,370).  The dependency (scipy -> OpenBLAS dpotrf) is not vendored in the
 * reference; this restates the published algorithm (LAPACK dpotf2 'U': for j: ajj = a_jj - u_j^T u_j; fail
 * if ajj <= 0 or NaN; u_jj = sqrt(ajj); row j of U = (a_j,j+1: - U_0:j,j^T U_0:j,j+1:)/u_jj).
 * The matrix is symmetric, so we factor in "lower, row-major" form, which is the same memory as
 * "upper, column-major": a[i*ld + j], j <= i holds L[i][j] = U[j][i].
 * Returns 0, or the 1-based index of the first non-positive pivot (LAPACK info). */
int oracle_potrf(double *a, long ld, int n)
{
    for (int j = 0; j < n; ++j) {
        double *aj = a + (long)j * ld;
        double s = aj[j];
        for (int k = 0; k < j; ++k) s -= aj[k] * aj[k];
        if (!(s > 0.0)) return j + 1;
        double d = sqrt(s);
        aj[j] = d;
        for (int i = j + 1; i < n; ++i) {
            double *ai = a + (long)i * ld;
            double t = ai[j];
            for (int k = 0; k < j; ++k) t -= ai[k] * aj[k];
            ai[j] = t / d;
        }
    }
    return 0;
}

/* LAPACK dpotrs for one right-hand side (scipy.linalg.cho_solve, covariance.py:331,354,376):
 * solve L y = b, then L^T x = y, in place. */
void oracle_potrs(const double *a, long ld, int n, double *b)
{
    for (int i = 0; i < n; ++i) {
        const double *ai = a + (long)i * ld;
        double t = b[i];
        for (int k = 0; k < i; ++k) t -= ai[k] * b[k];
        b[i] = t / ai[i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double t = b[i] / a[(long)i * ld + i];
        b[i] = t;
        for (int k = 0; k < i; ++k) b[k] -= a[(long)i * ld + k] * t;
    }
}

/* psoap/covariance.py:299-331 (lnlike_f), :333-354 (lnlike_f_g), :356-376 (lnlike_f_g_h).
 * V11 is caller-owned scratch [N,N] row-major (left holding the factor).  Returns -inf for negative
 * hyper-parameters (:317,:339,:362) and for a non-positive-definite matrix (:326-327). */
double oracle_lnlike(double *V11, int N, int ncomp, const double *lwl_f, const double *lwl_g,
                     const double *lwl_h, const double *fl, const double *sigma, const double *amp,
                     const double *l, double mu_GP)
{
    for (int c = 0; c < ncomp; ++c)
        if (amp[c] < 0.0 || l[c] < 0.0) return -INFINITY;
    oracle_fill_V11(V11, N, N, ncomp, lwl_f, lwl_g, lwl_h, amp, l);
    for (int i = 0; i < N; ++i) V11[(long)i * N + i] += sigma[i] * sigma[i]; /* :322 */
    if (oracle_potrf(V11, N, N) != 0) return -INFINITY;
    double logdet = 0.0;
    for (int i = 0; i < N; ++i) logdet += 2.0 * log(V11[(long)i * N + i]); /* :329 */
    double *r = (double *)malloc(sizeof(double) * (size_t)N);
    double *x = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; ++i) { r[i] = fl[i] - mu_GP; x[i] = r[i]; }
    oracle_potrs(V11, N, N, x);
    double quad = 0.0;
    for (int i = 0; i < N; ++i) quad += r[i] * x[i];
    free(r); free(x);
    return -0.5 * (quad + logdet); /* :331 */
}
