// saxpy.cu -- tuned SAXPY behind saxpy_kernel! / saxpy_kernel:  Z[I] = a * X[I] + Y[I]
//
// Reference: benchmark/benchmarks.jl:29-32 (Z, a, X, Y) and examples/numa_aware.jl:10-13 (in place: Y = a*X + Y), the
// only workloads the reference publishes numbers for.  The product and the sum are rounded separately (Julia does not
// contract `a * x + y` without @fastmath / muladd), so the kernel uses __fmul_rn / __fadd_rn -- never an FMA -- and the
// result is bit-identical to the CPU() backend.  HBM bound: 3 * sizeof(T) bytes per element.
// Layout: grid of 32 x 148 CTAs (tools/sweep.py saxpy: 6.5 TB/s), each thread keeps UNROLL independent 128-bit loads of X and of Y in
// flight before the arithmetic, then streams the 128-bit stores.  Z may alias Y (in-place form) but not X.
#include "../common.cuh"
#include "../tune.h"

namespace b200 {

template <typename V>
__device__ __forceinline__ V ld_nc(const V *p);
template <>
__device__ __forceinline__ float4 ld_nc<float4>(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ double2 ld_nc<double2>(const double2 *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 axpy(float a, const float4 &x, const float4 &y) {
    return make_float4(__fadd_rn(__fmul_rn(a, x.x), y.x), __fadd_rn(__fmul_rn(a, x.y), y.y),
                       __fadd_rn(__fmul_rn(a, x.z), y.z), __fadd_rn(__fmul_rn(a, x.w), y.w));
}
__device__ __forceinline__ double2 axpy(double a, const double2 &x, const double2 &y) {
    return make_double2(__dadd_rn(__dmul_rn(a, x.x), y.x), __dadd_rn(__dmul_rn(a, x.y), y.y));
}
__device__ __forceinline__ float axpy(float a, float x, float y) { return __fadd_rn(__fmul_rn(a, x), y); }
__device__ __forceinline__ double axpy(double a, double x, double y) { return __dadd_rn(__dmul_rn(a, x), y); }

// nv = number of 16-byte vectors; the scalar tail (n % VEC elements) is handled by the last CTA.
template <typename T, typename V, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) saxpy_vec_kernel(T *Z, T a, const T *__restrict__ X, const T *Y, size_t n) {
    constexpr int VEC = sizeof(V) / sizeof(T);
    const size_t nv = n / VEC;
    const size_t tile = (size_t)THREADS * UNROLL;
    const size_t ntiles = nv / tile;
    V *Zv = reinterpret_cast<V *>(Z);
    const V *Xv = reinterpret_cast<const V *>(X);
    const V *Yv = reinterpret_cast<const V *>(Y);
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
        V x[UNROLL], y[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) x[u] = ld_nc<V>(Xv + base + (size_t)u * THREADS);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) y[u] = Yv[base + (size_t)u * THREADS];   // may alias Z: ordinary load
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) __stcs(Zv + base + (size_t)u * THREADS, axpy(a, x[u], y[u]));
    }
    if (blockIdx.x == gridDim.x - 1) {
        for (size_t i = ntiles * tile + threadIdx.x; i < nv; i += THREADS) Zv[i] = axpy(a, Xv[i], Yv[i]);
        for (size_t i = nv * VEC + threadIdx.x; i < n; i += THREADS) Z[i] = axpy(a, X[i], Y[i]);
    }
}
template <typename T>
__global__ void saxpy_scalar_kernel(T *Z, T a, const T *__restrict__ X, const T *Y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        Z[i] = axpy(a, X[i], Y[i]);
}

int saxpy_launch(void *Z, double a, const void *X, const void *Y, size_t n, int is_f64, cudaStream_t st) {
    if (n == 0) return B200_OK;
    const bool aligned = ((((uintptr_t)Z | (uintptr_t)X | (uintptr_t)Y) & 15) == 0);
    const int grid = sm_count() * tune_get("saxpy.ctas_per_sm", 32);
    if (!aligned) {
        if (is_f64)
            saxpy_scalar_kernel<double><<<grid, 256, 0, st>>>((double *)Z, a, (const double *)X, (const double *)Y, n);
        else
            saxpy_scalar_kernel<float><<<grid, 256, 0, st>>>((float *)Z, (float)a, (const float *)X, (const float *)Y, n);
        B200_CHECK_LAUNCH("saxpy_scalar_kernel");
        return B200_OK;
    }
    if (is_f64)
        saxpy_vec_kernel<double, double2, 256, 4><<<grid, 256, 0, st>>>((double *)Z, a, (const double *)X, (const double *)Y, n);
    else
        saxpy_vec_kernel<float, float4, 256, 4><<<grid, 256, 0, st>>>((float *)Z, (float)a, (const float *)X, (const float *)Y, n);
    B200_CHECK_LAUNCH("saxpy_vec_kernel");
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_saxpy(void *Z, double a, const void *X, const void *Y, size_t n, int32_t is_f64) {
    if (n == 0) return B200_OK;
    if (!Z || !X || !Y) return set_error(B200_ERR_INVALID_VALUE, "b200_saxpy: NULL pointer");
    const size_t bytes = n * (is_f64 ? 8 : 4);
    const char *z = (const char *)Z, *x = (const char *)X;
    if (z < x + bytes && x < z + bytes) return set_error(B200_ERR_ALIAS, "b200_saxpy: Z may alias Y but not X");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return saxpy_launch(Z, a, X, Y, n, is_f64, st);
}
