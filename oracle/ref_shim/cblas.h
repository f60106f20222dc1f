/* TEST INFRASTRUCTURE ONLY (oracle/_ref build). Minimal stand-in for <cblas.h> so the
 * reference headers under /root/reference/src compile without a BLAS install.
 * Only the one routine the reference's hot path names (core/solver.cc:55) is declared. */
#ifndef JB_ORACLE_SHIM_CBLAS_H
#define JB_ORACLE_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
void cblas_daxpy(const int n, const double alpha, const double *x, const int incx, double *y, const int incy);
#ifdef __cplusplus
}
#endif
#endif
