/* TEST INFRASTRUCTURE ONLY (oracle/_ref build). Declares the four Fortran LAPACK symbols
 * that reference containers/mat3.h:147-251 calls for 3x3 determinant/inverse; they are
 * defined in oracle/ref_wrap.cpp as a plain partial-pivot LU (column-major, any small n). */
#ifndef JB_ORACLE_SHIM_LAPACKE_H
#define JB_ORACLE_SHIM_LAPACKE_H
#ifdef __cplusplus
extern "C" {
#endif
void dgetrf_(int *m, int *n, double *a, int *lda, int *ipiv, int *info);
void dgetri_(int *n, double *a, int *lda, int *ipiv, double *work, int *lwork, int *info);
void sgetrf_(int *m, int *n, float *a, int *lda, int *ipiv, int *info);
void sgetri_(int *n, float *a, int *lda, int *ipiv, float *work, int *lwork, int *info);
#ifdef __cplusplus
}
#endif
#endif
