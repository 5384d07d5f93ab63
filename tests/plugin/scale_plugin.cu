// A user-side kernel library for b200_register_kernel: the CUDA twins of
//     @kernel function scale_kernel!(A, s)            I = @index(Global, Linear);  A[I] *= s              end
//     @kernel function rowsum_kernel!(out, @Const(A)) i = @index(Global);  out[i] = sum_j A[i, j]  (j ascending) end
// built by tests/test_plugin_kernels.py with nvcc for sm_100a; not part of libb200ka.so.
#include <cuda_runtime.h>

#include "../../include/b200ka_plugin.cuh"

using b200::KaCtx;

__global__ void scale_kernel(KaCtx c, float *A, float s) {
    if (b200::ka_valid(c)) {
        const long long I = b200::ka_global_linear(c);
        A[I - 1] *= s;
    }
}
__global__ void rowsum_kernel(KaCtx c, double *out, const double *__restrict__ A, long long rows, long long cols) {
    if (b200::ka_valid(c)) {
        const long long i = b200::ka_global_linear(c);
        double s = 0.0;
        for (long long j = 0; j < cols; ++j) s += A[(i - 1) + j * rows];
        out[i - 1] = s;
    }
}

extern "C" int32_t scale_handler(const b200_launch_desc *d, const b200_arg *a, int32_t n, void *stream, void *user) {
    KaCtx c;
    long long groups, items;
    int32_t st = b200_plugin_ctx(d, &c, &groups, &items);
    if (st != B200_OK) return st;
    if (n != 2 || a[0].kind != B200_ARG_ARRAY || a[0].eltype != B200_T_F32 || a[1].kind != B200_ARG_SCALAR) return B200_ERR_BAD_ARGS;
    dim3 grid, block;
    b200::ka_launch_dims(c, &grid, &block);
    if (groups) scale_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(c, (float *)a[0].ptr, (float)a[1].fscalar);
    if (user) ++*(long long *)user;   // call counter owned by the test
    return B200_OK;
}
extern "C" int32_t rowsum_handler(const b200_launch_desc *d, const b200_arg *a, int32_t n, void *stream, void *) {
    KaCtx c;
    long long groups, items;
    int32_t st = b200_plugin_ctx(d, &c, &groups, &items);
    if (st != B200_OK) return st;
    if (n != 2 || a[0].eltype != B200_T_F64 || a[1].eltype != B200_T_F64 || a[1].ndim != 2 || a[0].dims[0] != a[1].dims[0]) return B200_ERR_BAD_ARGS;
    dim3 grid, block;
    b200::ka_launch_dims(c, &grid, &block);
    if (groups) rowsum_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(c, (double *)a[0].ptr, (const double *)a[1].ptr, a[1].dims[0], a[1].dims[1]);
    return B200_OK;
}
