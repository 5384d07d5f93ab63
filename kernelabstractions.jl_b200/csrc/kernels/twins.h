// twins.h -- host-side launchers of the tuned kernels and of the literal kernel twins (see registry.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../ka_device.cuh"

namespace b200 {
// tuned fast paths
int copy_bytes(void *dst, const void *src, size_t bytes, cudaStream_t st);
int fill_elems(void *dst, size_t count, int elsize, const void *value, cudaStream_t st);
int transpose_launch(void *a, const void *b, int64_t m, int64_t n, int64_t lda, int64_t ldb, int elsize,
                     cudaStream_t st);
int histogram_launch(int32_t *bins, int64_t nbins, const int32_t *input, int64_t n, cudaStream_t st);
int histogram64_launch(long long *bins, int64_t nbins, const long long *input, int64_t n, cudaStream_t st);
int sgemm_launch(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb,
                 int64_t ldc, int mode, cudaStream_t st);
int saxpy_launch(void *Z, double a, const void *X, const void *Y, size_t n, int is_f64, cudaStream_t st);
// literal twins (KA launches)
int twin_copy(const KaCtx &c, void *A, const void *B, int elsize, cudaStream_t st);
int twin_init(const KaCtx &c, void *arr, const void *value, int elsize, cudaStream_t st);
int twin_transpose(const KaCtx &c, void *a, const void *b, long long rows_a, long long rows_b, int elsize,
                   cudaStream_t st);
int twin_matmul_naive(const KaCtx &c, void *out, const void *a, const void *b, long long M, long long K, int is_f64,
                      cudaStream_t st);
int twin_localmem(const KaCtx &c, int variant, long long *A, long long lenA, cudaStream_t st);
int twin_stmt_form(const KaCtx &c, cudaStream_t st);
int twin_private(const KaCtx &c, long long *A, cudaStream_t st);
int twin_typetest(const KaCtx &c, uint8_t *B, cudaStream_t st);
int twin_forloop(const KaCtx &c, long long *A, long long rowsA, long long N, cudaStream_t st);
int twin_reduce_private(const KaCtx &c, void *out, const void *A, long long n, long long K, int is_f64,
                        cudaStream_t st);
int twin_unroll(const KaCtx &c, int variant, float *a, long long N, cudaStream_t st);
int twin_index_linear(const KaCtx &c, int which, long long *A, cudaStream_t st);
int twin_index_cartesian(const KaCtx &c, int which, long long *A, int ndimA, const long long *dimsA,
                         cudaStream_t st);
int twin_kernel_val(const KaCtx &c, long long *a, long long m, cudaStream_t st);
int twin_kernel_empty(const KaCtx &c, cudaStream_t st);
int twin_saxpy(const KaCtx &c, void *Z, double a, const void *X, const void *Y, int is_f64, cudaStream_t st);
int twin_misc_i64(const KaCtx &c, int which, long long *a, long long len, cudaStream_t st);
int twin_unaliased_accumulate(const KaCtx &c, int local, void *output, const void *input, long long n,
                              long long rows_out, long long rows_in, int is_f64, cudaStream_t st);
int twin_convert(const KaCtx &c, const void *A, void *B, long long rowsB, int is_f64, cudaStream_t st);
int twin_specialfn(const KaCtx &c, int which, float *A, const float *B, cudaStream_t st);
int twin_performance(const KaCtx &c, int which, void *out, const void *in, long long rows_out, long long rows_in,
                     long long bank, int elsize, cudaStream_t st);
// literal twins (KI launches)
int twin_histogram(int32_t *out, long long N, const int32_t *input, long long len, long long gs, long long groups,
                   cudaStream_t st);
int twin_coalesced_matmul(float *output, const float *in1, const float *in2, long long N, long long R, long long M,
                          int TDIM, long long groups_x, long long groups_y, cudaStream_t st);
int twin_test_intrinsics(long long *results, long long len, const long long *groups, const long long *wgs,
                         cudaStream_t st);
int twin_launch_kernel_nd(float *arr, const long long *dims, const long long *groups, const long long *wgs,
                          cudaStream_t st);
}  // namespace b200
