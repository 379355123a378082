/* fortran_blas.c — netlib-semantics Fortran-ABI BLAS for oracle/_ref.
 *
 * TEST INFRASTRUCTURE ONLY.  The reference links a third-party BLAS (MKL 2022 /
 * BLIS per installation_instructions:7,30-32) that is not vendored and not
 * present here; the symbols its host wrappers reference
 * (src/linearAlgebra/BlasLapackTemplates.h:154-520) are restated below from the
 * published netlib reference algorithms.  Only the double-precision real
 * routines are implemented; the s/c/z entry points exist solely so the shared
 * object loads and abort if ever called.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

typedef void (*orc_dgemm_fn)(const char *, const char *, const int *, const int *, const int *,
                             const double *, const double *, const int *, const double *, const int *,
                             const double *, double *, const int *);
static orc_dgemm_fn g_fast = 0;
void ref_set_dgemm(void *fn) { g_fast = (orc_dgemm_fn)fn; }

static int tr(char t) { return t == 'T' || t == 't' || t == 'C' || t == 'c'; }

void dgemm_(const char *ta, const char *tb, const unsigned *m_, const unsigned *n_, const unsigned *k_,
            const double *alpha_, const double *A, const unsigned *lda_, const double *B,
            const unsigned *ldb_, const double *beta_, double *C, const unsigned *ldc_)
{
    if (g_fast) {
        int m = (int)*m_, n = (int)*n_, k = (int)*k_, la = (int)*lda_, lb = (int)*ldb_, lc = (int)*ldc_;
        g_fast(ta, tb, &m, &n, &k, alpha_, A, &la, B, &lb, beta_, C, &lc);
        return;
    }
    const unsigned m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
    const double alpha = *alpha_, beta = *beta_;
    const int tA = tr(*ta), tB = tr(*tb);
    for (unsigned j = 0; j < n; ++j) {
        double *c = C + (size_t)j * ldc;
        if (beta == 0.0)
            for (unsigned i = 0; i < m; ++i) c[i] = 0.0;
        else if (beta != 1.0)
            for (unsigned i = 0; i < m; ++i) c[i] *= beta;
        for (unsigned l = 0; l < k; ++l) {
            const double b = alpha * (tB ? B[j + (size_t)l * ldb] : B[l + (size_t)j * ldb]);
            if (!tA)
                for (unsigned i = 0; i < m; ++i) c[i] += b * A[i + (size_t)l * lda];
            else
                for (unsigned i = 0; i < m; ++i) c[i] += b * A[l + (size_t)i * lda];
        }
    }
}

void daxpy_(const unsigned *n, const double *alpha, const double *x, const unsigned *incx, double *y,
            const unsigned *incy)
{
    for (unsigned i = 0; i < *n; ++i) y[(size_t)i * *incy] += *alpha * x[(size_t)i * *incx];
}
double dasum_(const unsigned *n, const double *x, const unsigned *incx)
{
    double s = 0.0;
    for (unsigned i = 0; i < *n; ++i) s += fabs(x[(size_t)i * *incx]);
    return s;
}
unsigned idamax_(const unsigned *n, const double *x, const unsigned *incx)
{
    unsigned best = 0;
    double bv = -1.0;
    for (unsigned i = 0; i < *n; ++i)
        if (fabs(x[(size_t)i * *incx]) > bv) { bv = fabs(x[(size_t)i * *incx]); best = i; }
    return *n ? best + 1 : 0;
}
double ddot_(const unsigned *n, const double *x, const unsigned *incx, const double *y, const unsigned *incy)
{
    double s = 0.0;
    for (unsigned i = 0; i < *n; ++i) s += x[(size_t)i * *incx] * y[(size_t)i * *incy];
    return s;
}
double dnrm2_(const unsigned *n, const double *x, const unsigned *incx)
{
    double scale = 0.0, ssq = 1.0;
    for (unsigned i = 0; i < *n; ++i) {
        const double a = fabs(x[(size_t)i * *incx]);
        if (a != 0.0) {
            if (scale < a) { ssq = 1.0 + ssq * (scale / a) * (scale / a); scale = a; }
            else ssq += (a / scale) * (a / scale);
        }
    }
    return scale * sqrt(ssq);
}
void dscal_(const unsigned *n, const double *a, double *x, const unsigned *incx)
{
    for (unsigned i = 0; i < *n; ++i) x[(size_t)i * *incx] *= *a;
}
void dcopy_(const unsigned *n, const double *x, const unsigned *incx, double *y, const unsigned *incy)
{
    for (unsigned i = 0; i < *n; ++i) y[(size_t)i * *incy] = x[(size_t)i * *incx];
}

#define UNUSED_BLAS(name) void name(void) { fprintf(stderr, "oracle/_ref: " #name " is not on the FP64 hot path\n"); abort(); }
UNUSED_BLAS(caxpy_) UNUSED_BLAS(cgemm_) UNUSED_BLAS(dzasum_) UNUSED_BLAS(icamax_) UNUSED_BLAS(isamax_)
UNUSED_BLAS(izamax_) UNUSED_BLAS(sasum_) UNUSED_BLAS(saxpy_) UNUSED_BLAS(scasum_) UNUSED_BLAS(sgemm_)
UNUSED_BLAS(zaxpy_) UNUSED_BLAS(zgemm_) UNUSED_BLAS(sscal_) UNUSED_BLAS(zscal_) UNUSED_BLAS(zdscal_)
UNUSED_BLAS(scopy_) UNUSED_BLAS(zcopy_) UNUSED_BLAS(ccopy_) UNUSED_BLAS(zdotc_) UNUSED_BLAS(dznrm2_)
UNUSED_BLAS(snrm2_) UNUSED_BLAS(sdot_) UNUSED_BLAS(cdotc_) UNUSED_BLAS(scnrm2_) UNUSED_BLAS(cscal_) UNUSED_BLAS(csscal_)
