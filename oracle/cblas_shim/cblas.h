/*
 * Minimal cblas.h for building the reference's `make blas` configuration
 * (reference Makefile:48-49) in an image that ships OpenBLAS only as a
 * wheel-bundled shared object without headers.  cblas_sgemm is the single
 * BLAS symbol the reference calls (qwen_asr_kernels.c:181,199,613,657).
 * TEST INFRASTRUCTURE ONLY - part of oracle/, never linked into the product.
 */
#ifndef QASR_ORACLE_CBLAS_SHIM_H
#define QASR_ORACLE_CBLAS_SHIM_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta,
                 const enum CBLAS_TRANSPOSE tb, const int M, const int N, const int K,
                 const float alpha, const float *A, const int lda, const float *B,
                 const int ldb, const float beta, float *C, const int ldc);
void openblas_set_num_threads(int n);
#endif
