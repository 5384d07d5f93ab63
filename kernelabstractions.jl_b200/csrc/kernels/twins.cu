// twins.cu -- literal CUDA twins of every kernel the reference defines on the hot path.
//
// There is no JIT in this backend, so each `@kernel` / KI kernel of the reference's examples and testsuite
// exists here as a precompiled sm_100a kernel that follows the lowering `transform_gpu!`/`split`/`emit`
// produce (reference src/macros.jl:55-88,113-196): `active = __validindex(ctx)`,
// hoisted @localmem/@private/@uniform allocations, `if active` regions separated by unguarded barriers.
// They honour ANY ndrange / workgroupsize the host passes.  The registry (registry.cu) routes the
// north-star shapes to the tuned kernels in copy.cu / transpose.cu / histogram.cu / matmul_f32.cu and
// everything else here, so results never depend on which one ran.
#include "twins.h"

#include <cstring>

#include "../common.cuh"
#include "../tune.h"

namespace b200 {

// Launch geometry of a KA launch (ka_device.cuh): a native <=3-d grid of <=3-d blocks when the limits allow (group / item ids are
// blockIdx / threadIdx, no division), else the reference's 1-d launch with multiply-high unravelling.  ka_grid x ka_block is
// prod(blocks) groups of prod(wgs) items either way.
#define KA_LAUNCH_1D(ctx, groups, items)                                                               \
    long long groups = 1, items = 1;                                                                   \
    for (int d = 0; d < (ctx).ndim; ++d) {                                                             \
        groups *= (ctx).blocks[d];                                                                     \
        items *= (ctx).wgs[d];                                                                         \
    }                                                                                                  \
    if (groups == 0) return B200_OK; /* reference src/pocl/backend.jl:136-138 */                       \
    if (items > 1024 || items < 1)                                                                     \
        return set_error(B200_ERR_LAUNCH_DIMS, "workgroup of %lld items exceeds 1024", items);         \
    if (groups > 2147483647LL) return set_error(B200_ERR_LAUNCH_DIMS, "%lld groups exceed 2^31-1", groups); \
    dim3 ka_grid, ka_block;                                                                            \
    ka_launch_dims((ctx), &ka_grid, &ka_block);

// ---- 128-bit vectorised lowering of the element-wise kernels on @index(Global, Linear) -------------------------------
// copy_kernel!, init_kernel and saxpy_kernel! are `A[I] = f(B[I], ...)` with I the global linear id, no @localmem and no
// @synchronize: the work-items are independent of their grouping.  When the valid work-items of the launch are exactly
// I = 1..n (a 1-d ndrange, or an N-d one without ragged dimensions -- then @index(Global, Linear) enumerates a contiguous
// range) one CUDA thread executes the 16 / sizeof(T) consecutive work-items that share a 16-byte word, four such words in
// flight per thread: one work-item per thread keeps only 8 KiB of loads in flight per SM, a quarter of what a B200 needs to
// cover its HBM latency (measured: 0.73 TB/s for the reference's div/mod lowering, r05 bench).
// `registry.twin_vec = 0` keeps the scalar one-item-per-thread form.  Same bits either way.
constexpr int KA_VEC_U = 4;   // 16-byte words in flight per thread

static inline bool ka_vectorisable(const KaCtx &c, long long *n_valid) {
    long long total = 1, nd = 1;
    for (int d = 0; d < c.ndim; ++d) {
        total *= c.blocks[d] * c.wgs[d];
        nd *= c.ndrange[d];
    }
    if (c.ndim == 1) {
        *n_valid = c.dynamic_check && c.ndrange[0] < total ? c.ndrange[0] : total;
        return true;
    }
    if (c.ragged != 0) return false;
    *n_valid = c.dynamic_check ? nd : total;
    return true;
}
static inline bool twin_vec_enabled() { return tune_get("registry.twin_vec", 1) != 0; }
__device__ __forceinline__ uint4 ka_ldg16(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void ka_stg16(void *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// bytes [0, nbytes) = work-items 1..n; CTA b owns 16-byte words [b * 256 * U, (b + 1) * 256 * U); the last CTA also moves the
// (< 16) bytes behind the last whole word
__global__ void __launch_bounds__(256) copy_twin_vec_kernel(char *__restrict__ A, const char *__restrict__ B, long long nvec, int tail_bytes) {
    const long long v0 = (long long)blockIdx.x * (256 * KA_VEC_U) + threadIdx.x;
    uint4 w[KA_VEC_U];
#pragma unroll
    for (int u = 0; u < KA_VEC_U; ++u)
        if (v0 + u * 256 < nvec) w[u] = ka_ldg16(B + 16 * (v0 + u * 256));
#pragma unroll
    for (int u = 0; u < KA_VEC_U; ++u)
        if (v0 + u * 256 < nvec) ka_stg16(A + 16 * (v0 + u * 256), w[u]);
    if (blockIdx.x == gridDim.x - 1 && (int)threadIdx.x < tail_bytes) A[16 * nvec + threadIdx.x] = B[16 * nvec + threadIdx.x];
}
__global__ void __launch_bounds__(256) init_twin_vec_kernel(char *__restrict__ A, uint4 pattern, long long nvec, int tail_bytes) {
    const long long v0 = (long long)blockIdx.x * (256 * KA_VEC_U) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < KA_VEC_U; ++u)
        if (v0 + u * 256 < nvec) ka_stg16(A + 16 * (v0 + u * 256), pattern);
    if (blockIdx.x == gridDim.x - 1 && (int)threadIdx.x < tail_bytes)
        A[16 * nvec + threadIdx.x] = reinterpret_cast<const char *>(&pattern)[threadIdx.x];   // tail starts on a 16-byte word: same phase
}

// ---- copy_kernel / copy_kernel!  (src/KernelAbstractions.jl:874-877, examples/memcopy.jl:4-7) ----
template <typename T>
__global__ void copy_twin_kernel(KaCtx c, T *__restrict__ A, const T *__restrict__ B) {
    const bool active = ka_valid(c);
    if (active) {
        const long long I = ka_global_linear(c);
        A[I - 1] = B[I - 1];
    }
}
int twin_copy(const KaCtx &c, void *A, const void *B, int elsize, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    long long n_valid = 0;
    if (twin_vec_enabled() && ka_vectorisable(c, &n_valid) && ((((uintptr_t)A | (uintptr_t)B) & 15) == 0) && 16 % elsize == 0) {
        if (n_valid <= 0) return B200_OK;
        const long long bytes = n_valid * elsize, nvec = bytes / 16;
        const long long ctas = (nvec + 256 * KA_VEC_U - 1) / (256 * KA_VEC_U);
        copy_twin_vec_kernel<<<(unsigned)(ctas < 1 ? 1 : ctas), 256, 0, st>>>((char *)A, (const char *)B, nvec, (int)(bytes - 16 * nvec));
        B200_CHECK_LAUNCH("copy_twin_vec_kernel");
        return B200_OK;
    }
    switch (elsize) {
        case 1: copy_twin_kernel<uint8_t><<<ka_grid, ka_block, 0, st>>>(c, (uint8_t *)A, (const uint8_t *)B); break;
        case 2: copy_twin_kernel<uint16_t><<<ka_grid, ka_block, 0, st>>>(c, (uint16_t *)A, (const uint16_t *)B); break;
        case 4: copy_twin_kernel<uint32_t><<<ka_grid, ka_block, 0, st>>>(c, (uint32_t *)A, (const uint32_t *)B); break;
        case 8: copy_twin_kernel<uint64_t><<<ka_grid, ka_block, 0, st>>>(c, (uint64_t *)A, (const uint64_t *)B); break;
        case 16: copy_twin_kernel<uint4><<<ka_grid, ka_block, 0, st>>>(c, (uint4 *)A, (const uint4 *)B); break;
        default: return set_error(B200_ERR_UNSUPPORTED, "copy_kernel: element size %d", elsize);
    }
    B200_CHECK_LAUNCH("copy_twin_kernel");
    return B200_OK;
}

// ---- init_kernel  (src/KernelAbstractions.jl:869-872) ----
template <typename T>
__global__ void init_twin_kernel(KaCtx c, T *__restrict__ arr, T value) {
    if (ka_valid(c)) arr[ka_global_linear(c) - 1] = value;
}
int twin_init(const KaCtx &c, void *arr, const void *value, int elsize, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    long long n_valid = 0;
    if (twin_vec_enabled() && ka_vectorisable(c, &n_valid) && (((uintptr_t)arr & 15) == 0) && 16 % elsize == 0) {
        if (n_valid <= 0) return B200_OK;
        unsigned char pat[16];
        for (int k = 0; k < 16; ++k) pat[k] = ((const unsigned char *)value)[k % elsize];
        uint4 pattern;
        memcpy(&pattern, pat, 16);
        const long long bytes = n_valid * elsize, nvec = bytes / 16;
        const long long ctas = (nvec + 256 * KA_VEC_U - 1) / (256 * KA_VEC_U);
        init_twin_vec_kernel<<<(unsigned)(ctas < 1 ? 1 : ctas), 256, 0, st>>>((char *)arr, pattern, nvec, (int)(bytes - 16 * nvec));
        B200_CHECK_LAUNCH("init_twin_vec_kernel");
        return B200_OK;
    }
    switch (elsize) {
        case 1: init_twin_kernel<uint8_t><<<ka_grid, ka_block, 0, st>>>(c, (uint8_t *)arr, *(const uint8_t *)value); break;
        case 2: init_twin_kernel<uint16_t><<<ka_grid, ka_block, 0, st>>>(c, (uint16_t *)arr, *(const uint16_t *)value); break;
        case 4: init_twin_kernel<uint32_t><<<ka_grid, ka_block, 0, st>>>(c, (uint32_t *)arr, *(const uint32_t *)value); break;
        case 8: init_twin_kernel<uint64_t><<<ka_grid, ka_block, 0, st>>>(c, (uint64_t *)arr, *(const uint64_t *)value); break;
        default: return set_error(B200_ERR_UNSUPPORTED, "init_kernel: element size %d", elsize);
    }
    B200_CHECK_LAUNCH("init_twin_kernel");
    return B200_OK;
}

// ---- naive_transpose_kernel!  (examples/naive_transpose.jl:4-7) ----
template <typename T>
__global__ void transpose_twin_kernel(KaCtx c, T *__restrict__ a, const T *__restrict__ b, long long rows_a,
                                      long long rows_b) {
    if (ka_valid(c)) {
        long long I[KA_MAX_NDIM];
        ka_expand(c, I);
        const long long i = I[0], j = I[1];
        a[(i - 1) + (j - 1) * rows_a] = b[(j - 1) + (i - 1) * rows_b];
    }
}
int twin_transpose(const KaCtx &c, void *a, const void *b, long long rows_a, long long rows_b, int elsize,
                   cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    switch (elsize) {
        case 1: transpose_twin_kernel<uint8_t><<<ka_grid, ka_block, 0, st>>>(c, (uint8_t *)a, (const uint8_t *)b, rows_a, rows_b); break;
        case 2: transpose_twin_kernel<uint16_t><<<ka_grid, ka_block, 0, st>>>(c, (uint16_t *)a, (const uint16_t *)b, rows_a, rows_b); break;
        case 4: transpose_twin_kernel<uint32_t><<<ka_grid, ka_block, 0, st>>>(c, (uint32_t *)a, (const uint32_t *)b, rows_a, rows_b); break;
        case 8: transpose_twin_kernel<uint64_t><<<ka_grid, ka_block, 0, st>>>(c, (uint64_t *)a, (const uint64_t *)b, rows_a, rows_b); break;
        case 16:
            if ((((uintptr_t)a | (uintptr_t)b) & 15) != 0) return set_error(B200_ERR_UNSUPPORTED, "naive_transpose_kernel!: 16-byte elements need 16-byte aligned arrays");
            transpose_twin_kernel<ulonglong2><<<ka_grid, ka_block, 0, st>>>(c, (ulonglong2 *)a, (const ulonglong2 *)b, rows_a, rows_b);
            break;
        default: return set_error(B200_ERR_UNSUPPORTED, "naive_transpose_kernel!: element size %d not in {1,2,4,8,16}", elsize);
    }
    B200_CHECK_LAUNCH("transpose_twin_kernel");
    return B200_OK;
}

// ---- matmul_kernel!  (examples/matmul.jl:5-15): separate multiply and add roundings, k ascending ----
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
template <typename T>
__global__ void matmul_naive_twin_kernel(KaCtx c, T *__restrict__ out, const T *__restrict__ a,
                                         const T *__restrict__ b, long long M, long long K) {
    if (ka_valid(c)) {
        long long I[KA_MAX_NDIM];
        ka_expand(c, I);
        const long long i = I[0], j = I[1];
        T tmp = (T)0;
        for (long long k = 1; k <= K; ++k) tmp = add_rn(tmp, mul_rn(a[(i - 1) + (k - 1) * M], b[(k - 1) + (j - 1) * K]));
        out[(i - 1) + (j - 1) * M] = tmp;
    }
}
int twin_matmul_naive(const KaCtx &c, void *out, const void *a, const void *b, long long M, long long K, int is_f64,
                      cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    if (is_f64)
        matmul_naive_twin_kernel<double><<<ka_grid, ka_block, 0, st>>>(c, (double *)out, (const double *)a, (const double *)b, M, K);
    else
        matmul_naive_twin_kernel<float><<<ka_grid, ka_block, 0, st>>>(c, (float *)out, (const float *)a, (const float *)b, M, K);
    B200_CHECK_LAUNCH("matmul_naive_twin_kernel");
    return B200_OK;
}

// ---- test/localmem.jl:4-65 ----
__global__ void localmem_twin_kernel(KaCtx c, int variant, long long *__restrict__ A, long long lenA) {
    extern __shared__ __align__(16) long long lmem[];  // @localmem Int (N,)  [x2 for many_localmem]
    const long long N = ka_groupsize_prod(c);           // N = @uniform prod(@groupsize()); N2 likewise
    const long long i = ka_local_linear(c);
    if (variant == 2) {  // unsafe_indices = true: no guard, no region splitting (test/localmem.jl:37-48)
        const long long gI = ka_group_linear(c);
        lmem[i - 1] = i;
        __syncthreads();
        const long long I = (gI - 1) * N + i;
        if (I <= lenA) A[I - 1] = lmem[N - i + 1 - 1];
        return;
    }
    const bool active = ka_valid(c);
    const long long I = ka_global_linear(c);
    if (variant == 0) {
        if (active) lmem[i - 1] = i;
        __syncthreads();
        if (active) A[I - 1] = lmem[N - i + 1 - 1];
    } else if (variant == 1) {
        if (active) lmem[i - 1] = i + 3;
        for (long long j = 1; j <= 2; ++j) {  // loop header on all lanes
            if (active) lmem[i - 1] -= j;
            __syncthreads();
        }
        if (active) A[I - 1] = lmem[N - i + 1 - 1];
    } else {
        long long *lmem2 = lmem + N;
        if (active) {
            lmem[i - 1] = i - 1;
            lmem2[i - 1] = 1;
        }
        __syncthreads();
        if (active) A[I - 1] = lmem[N - i + 1 - 1] + lmem2[N - i + 1 - 1];
    }
}
int twin_localmem(const KaCtx &c, int variant, long long *A, long long lenA, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    size_t smem = (size_t)items * 8 * (variant == 3 ? 2 : 1);
    localmem_twin_kernel<<<ka_grid, ka_block, smem, st>>>(c, variant, A, lenA);
    B200_CHECK_LAUNCH("localmem_twin_kernel");
    return B200_OK;
}

// ---- test/private.jl:4-59 ----
__global__ void stmt_form_twin_kernel(KaCtx c) {
    const long long bs = c.wgs[0];   // @uniform bs = @groupsize()[1]
    volatile long long s = bs / 2;   // @private s = bs / 2
    (void)s;
    __syncthreads();
}
__global__ void private_twin_kernel(KaCtx c, long long *__restrict__ A) {
    const long long N = ka_groupsize_prod(c);
    const bool active = ka_valid(c);
    long long priv[1];
    const long long I = ka_global_linear(c), i = ka_local_linear(c);
    if (active) priv[0] = N - i + 1;
    __syncthreads();
    if (active) A[I - 1] = priv[0];
}
__global__ void typetest_twin_kernel(KaCtx c, uint8_t *__restrict__ B) {
    if (ka_valid(c)) B[ka_global_linear(c) - 1] = 1;  // eltype(priv) === eltype(A)
}
constexpr int FORLOOP_MAX_N = 128;
__global__ void forloop_twin_kernel(KaCtx c, long long *__restrict__ A, long long rowsA, int N) {
    const bool active = ka_valid(c);
    const long long I = ka_global_linear(c), i = ka_local_linear(c);
    long long priv[FORLOOP_MAX_N];  // @private Int (N,)
    if (active) {
        for (int j = 1; j <= N; ++j) priv[j - 1] = A[(I - 1) + (long long)(j - 1) * rowsA];
        A[I - 1] = 0;
    }
    __syncthreads();
    for (int j = 1; j <= N; ++j) {
        if (active) {
            const long long k = (j + i - 1 - 1) % N + 1;  // mod1(j + i - 1, N)
            A[k - 1] += priv[j - 1];
        }
        __syncthreads();
    }
}
template <typename T>
__global__ void reduce_private_twin_kernel(KaCtx c, T *__restrict__ out, const T *__restrict__ A, long long n,
                                           long long K) {
    if (ka_valid(c)) {
        long long I[KA_MAX_NDIM];
        ka_expand(c, I);
        T priv = (T)0;
        for (long long k = 1; k <= K; ++k) priv = add_rn(priv, A[(I[0] - 1) + (k - 1) * n]);
        out[I[0] - 1] = priv;
    }
}
int twin_stmt_form(const KaCtx &c, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    stmt_form_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c);
    B200_CHECK_LAUNCH("stmt_form_twin_kernel");
    return B200_OK;
}
int twin_private(const KaCtx &c, long long *A, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    private_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, A);
    B200_CHECK_LAUNCH("private_twin_kernel");
    return B200_OK;
}
int twin_typetest(const KaCtx &c, uint8_t *B, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    typetest_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, B);
    B200_CHECK_LAUNCH("typetest_twin_kernel");
    return B200_OK;
}
int twin_forloop(const KaCtx &c, long long *A, long long rowsA, long long N, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    if (N < 0 || N > FORLOOP_MAX_N)
        return set_error(B200_ERR_UNSUPPORTED, "forloop: Val(N) with N=%lld > %d private slots", N, FORLOOP_MAX_N);
    forloop_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, A, rowsA, (int)N);
    B200_CHECK_LAUNCH("forloop_twin_kernel");
    return B200_OK;
}
int twin_reduce_private(const KaCtx &c, void *out, const void *A, long long n, long long K, int is_f64,
                        cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    if (is_f64)
        reduce_private_twin_kernel<double><<<ka_grid, ka_block, 0, st>>>(c, (double *)out, (const double *)A, n, K);
    else
        reduce_private_twin_kernel<float><<<ka_grid, ka_block, 0, st>>>(c, (float *)out, (const float *)A, n, K);
    B200_CHECK_LAUNCH("reduce_private_twin_kernel");
    return B200_OK;
}

// ---- test/unroll.jl:5-37 ----
__global__ void unroll_twin_kernel(KaCtx c, int variant, float *__restrict__ a, int N) {
    const bool active = ka_valid(c);
    if (variant == 0) {
        if (active) {
#pragma unroll
            for (int i = 1; i <= 5; ++i) a[i - 1] = (float)i;
        }
    } else if (variant == 1) {
        if (active) {
            const int M = N + 5;
#pragma unroll 4
            for (int i = 6; i <= M; ++i) a[i - 5 - 1] = (float)i;
        }
        __syncthreads();
    } else {
        const float va[3] = {1.f, 2.f, 3.f}, vb[3] = {3.f, 2.f, 1.f};  // @uniform MVectors
        float vc[9];
        const long long I = ka_global_linear(c);
        for (int m = 1; m <= 3; ++m) {
            if (active) {
#pragma unroll
                for (int j = 1; j <= 3; ++j)
#pragma unroll
                    for (int i = 1; i <= 3; ++i) vc[(j - 1) * 3] = (float)m * va[0] * vb[j - 1];
                a[I - 1] = vc[0];
            }
            __syncthreads();  // @synchronize(m % 2 == 0): the condition is ignored (src/KernelAbstractions.jl:319-345)
        }
    }
}
int twin_unroll(const KaCtx &c, int variant, float *a, long long N, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    unroll_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, variant, a, (int)N);
    B200_CHECK_LAUNCH("unroll_twin_kernel");
    return B200_OK;
}

// ---- test/test.jl:46-73 index kernels, :216-230 kernel_val!/kernel_empty ----
__global__ void index_linear_twin_kernel(KaCtx c, int which, long long *__restrict__ A) {
    if (ka_valid(c)) {
        const long long I = ka_global_linear(c);
        A[I - 1] = which == 0 ? I : (which == 1 ? ka_local_linear(c) : ka_group_linear(c));
    }
}
__global__ void index_cartesian_twin_kernel(KaCtx c, int which, long long *__restrict__ A, int ndimA, long long d0,
                                            long long d1, long long d2) {
    if (ka_valid(c)) {
        long long I[KA_MAX_NDIM], v[KA_MAX_NDIM];
        ka_expand(c, I);
        if (which == 0) {
            for (int d = 0; d < c.ndim; ++d) v[d] = I[d];
        } else if (which == 1) {
            ka_local_cartesian(c, v);
        } else {
            ka_group_cartesian(c, v);
        }
        const long long dims[3] = {d0, d1, d2};
        long long lin;
        if (c.ndim == ndimA) {
            lin = 0;
            long long stride = 1;
            for (int d = 0; d < c.ndim; ++d) {
                lin += (I[d] - 1) * stride;
                stride *= dims[d < 3 ? d : 2];
            }
        } else {
            lin = I[0] - 1;
        }
        for (int d = 0; d < c.ndim; ++d) A[lin * c.ndim + d] = v[d];
    }
}
__global__ void kernel_val_twin_kernel(KaCtx c, long long *__restrict__ a, long long m) {
    if (ka_valid(c)) a[ka_global_linear(c) - 1] = m;
}
__global__ void kernel_empty_twin_kernel(KaCtx c) { (void)ka_valid(c); }
int twin_index_linear(const KaCtx &c, int which, long long *A, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    index_linear_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, which, A);
    B200_CHECK_LAUNCH("index_linear_twin_kernel");
    return B200_OK;
}
int twin_index_cartesian(const KaCtx &c, int which, long long *A, int ndimA, const long long *dimsA, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    if (c.ndim > 3) return set_error(B200_ERR_UNSUPPORTED, "index_cartesian: more than 3 dims");
    index_cartesian_twin_kernel<<<ka_grid, ka_block, 0, st>>>(
        c, which, A, ndimA, ndimA > 0 ? dimsA[0] : 1, ndimA > 1 ? dimsA[1] : 1, ndimA > 2 ? dimsA[2] : 1);
    B200_CHECK_LAUNCH("index_cartesian_twin_kernel");
    return B200_OK;
}
int twin_kernel_val(const KaCtx &c, long long *a, long long m, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    kernel_val_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, a, m);
    B200_CHECK_LAUNCH("kernel_val_twin_kernel");
    return B200_OK;
}
int twin_kernel_empty(const KaCtx &c, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    kernel_empty_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c);
    B200_CHECK_LAUNCH("kernel_empty_twin_kernel");
    return B200_OK;
}

// ---- test/test.jl:275-321 context_kernel / gpu_return_kernel!, :339-385 unaliased_accumulate(_local)! ----
// which 0: context_kernel      f(@context, a): a[@index(Global, Linear)] = 1
//       1: gpu_return_kernel!  if i <= length(x) / 2 { x[i] = 1; return }
__global__ void misc_i64_twin_kernel(KaCtx c, int which, long long *__restrict__ a, long long len) {
    if (ka_valid(c)) {
        const long long I = ka_global_linear(c);
        if (which == 0) {
            a[I - 1] = 1;
        } else if (I <= len / 2) {
            a[I - 1] = 1;
            return;
        }
    }
}
int twin_misc_i64(const KaCtx &c, int which, long long *a, long long len, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    misc_i64_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, which, a, len);
    B200_CHECK_LAUNCH("misc_i64_twin_kernel");
    return B200_OK;
}
// LOCAL == false: output[i, j] += input[i, k] for k in j:n (read-modify-write of global memory every step);
// LOCAL == true : sum_val = zero(T); sum_val += input[i, k]; output[i, j] = sum_val.  Plain adds, k ascending.
template <typename T, bool LOCAL>
__global__ void unaliased_accumulate_twin_kernel(KaCtx c, T *output, const T *input, long long n, long long rows_out,
                                                 long long rows_in) {
    if (ka_valid(c)) {
        long long I[KA_MAX_NDIM];
        ka_expand(c, I);
        const long long i = I[0], j = I[1];
        if (LOCAL) {
            T sum_val = (T)0;
            for (long long k = j; k <= n; ++k) sum_val = add_rn(sum_val, input[(i - 1) + (k - 1) * rows_in]);
            output[(i - 1) + (j - 1) * rows_out] = sum_val;
        } else {
            volatile T *o = output + (i - 1) + (j - 1) * rows_out;
            for (long long k = j; k <= n; ++k) *o = add_rn(*o, input[(i - 1) + (k - 1) * rows_in]);
        }
    }
}
int twin_unaliased_accumulate(const KaCtx &c, int local, void *output, const void *input, long long n,
                              long long rows_out, long long rows_in, int is_f64, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    if (is_f64) {
        if (local) unaliased_accumulate_twin_kernel<double, true><<<ka_grid, ka_block, 0, st>>>(c, (double *)output, (const double *)input, n, rows_out, rows_in);
        else unaliased_accumulate_twin_kernel<double, false><<<ka_grid, ka_block, 0, st>>>(c, (double *)output, (const double *)input, n, rows_out, rows_in);
    } else {
        if (local) unaliased_accumulate_twin_kernel<float, true><<<ka_grid, ka_block, 0, st>>>(c, (float *)output, (const float *)input, n, rows_out, rows_in);
        else unaliased_accumulate_twin_kernel<float, false><<<ka_grid, ka_block, 0, st>>>(c, (float *)output, (const float *)input, n, rows_out, rows_in);
    }
    B200_CHECK_LAUNCH("unaliased_accumulate_twin_kernel");
    return B200_OK;
}

// ---- test/convert.jl:5-44 convert_kernel!: ceil / floor / round(ties to even) of A[tid] through ten integer types ----
// Columns 5, 10, 15, 20, 25, 30 (Int128 / UInt128) are left untouched like the commented-out lines of the reference.
template <typename T, typename I>
__device__ __forceinline__ T through_int(T x) { return (T)(I)x; }
template <typename T>
__global__ void convert_twin_kernel(KaCtx c, const T *__restrict__ A, T *__restrict__ B, long long rowsB) {
    if (ka_valid(c)) {
        long long tid = ka_global_linear(c);
        tid = (long long)(int)tid;   // Int64(Int32(tid))
        const T a = A[tid - 1];
        const T v[3] = {(T)ceil((double)a), (T)floor((double)a), (T)rint((double)a)};
#pragma unroll
        for (int op = 0; op < 3; ++op) {
            T *col = B + (tid - 1) + (long long)(10 * op) * rowsB;
            col[0 * rowsB] = through_int<T, signed char>(v[op]);
            col[1 * rowsB] = through_int<T, short>(v[op]);
            col[2 * rowsB] = through_int<T, int>(v[op]);
            col[3 * rowsB] = through_int<T, long long>(v[op]);
            col[5 * rowsB] = through_int<T, unsigned char>(v[op]);
            col[6 * rowsB] = through_int<T, unsigned short>(v[op]);
            col[7 * rowsB] = through_int<T, unsigned int>(v[op]);
            col[8 * rowsB] = through_int<T, unsigned long long>(v[op]);
        }
    }
}
int twin_convert(const KaCtx &c, const void *A, void *B, long long rowsB, int is_f64, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    if (is_f64)
        convert_twin_kernel<double><<<ka_grid, ka_block, 0, st>>>(c, (const double *)A, (double *)B, rowsB);
    else
        convert_twin_kernel<float><<<ka_grid, ka_block, 0, st>>>(c, (const float *)A, (float *)B, rowsB);
    B200_CHECK_LAUNCH("convert_twin_kernel");
    return B200_OK;
}

// ---- test/specialfunctions.jl:5-18 gamma_knl / erf_knl / erfc_knl (Float32; compared with isapprox upstream) ----
__global__ void specialfn_twin_kernel(KaCtx c, int which, float *__restrict__ A, const float *__restrict__ B) {
    if (ka_valid(c)) {
        const long long I = ka_global_linear(c);
        const float x = B[I - 1];
        A[I - 1] = which == 0 ? tgammaf(x) : (which == 1 ? erff(x) : erfcf(x));
    }
}
int twin_specialfn(const KaCtx &c, int which, float *A, const float *B, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    specialfn_twin_kernel<<<ka_grid, ka_block, 0, st>>>(c, which, A, B);
    B200_CHECK_LAUNCH("specialfn_twin_kernel");
    return B200_OK;
}

// ---- saxpy_kernel!  (benchmark/benchmarks.jl:29-32, examples/numa_aware.jl:10-13): Z = a*X + Y, unfused ----
template <typename T>
__global__ void saxpy_twin_kernel(KaCtx c, T *__restrict__ Z, T a, const T *__restrict__ X, const T *__restrict__ Y) {
    if (ka_valid(c)) {
        const long long I = ka_global_linear(c);
        Z[I - 1] = add_rn(mul_rn(a, X[I - 1]), Y[I - 1]);
    }
}
// vectorised form (see copy_twin_vec_kernel): V = 16 / sizeof(T) work-items per 16-byte word, product and sum rounded separately
template <typename T>
__global__ void __launch_bounds__(256) saxpy_twin_vec_kernel(T *__restrict__ Z, T a, const T *__restrict__ X, const T *__restrict__ Y, long long nvec,
                                                             long long n) {
    constexpr int V = 16 / sizeof(T);
    const long long v0 = (long long)blockIdx.x * (256 * KA_VEC_U) + threadIdx.x;
    uint4 x[KA_VEC_U], y[KA_VEC_U];
#pragma unroll
    for (int u = 0; u < KA_VEC_U; ++u)
        if (v0 + u * 256 < nvec) {
            x[u] = ka_ldg16(X + V * (v0 + u * 256));
            y[u] = ka_ldg16(Y + V * (v0 + u * 256));
        }
#pragma unroll
    for (int u = 0; u < KA_VEC_U; ++u)
        if (v0 + u * 256 < nvec) {
            uint4 z;
            const T *xs = reinterpret_cast<const T *>(&x[u]), *ys = reinterpret_cast<const T *>(&y[u]);
            T *zs = reinterpret_cast<T *>(&z);
#pragma unroll
            for (int k = 0; k < V; ++k) zs[k] = add_rn(mul_rn(a, xs[k]), ys[k]);
            ka_stg16(Z + V * (v0 + u * 256), z);
        }
    if (blockIdx.x == gridDim.x - 1) {      // the (< V) work-items behind the last whole word
        const long long I = V * nvec + threadIdx.x;
        if (I < n) Z[I] = add_rn(mul_rn(a, X[I]), Y[I]);
    }
}
int twin_saxpy(const KaCtx &c, void *Z, double a, const void *X, const void *Y, int is_f64, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    long long n_valid = 0;
    // in place (Z == Y, examples/numa_aware.jl:10-13) is fine: a work-item reads its own element before writing it; Z may not
    // overlap X or Y in any other way (the registry checks X)
    if (twin_vec_enabled() && ka_vectorisable(c, &n_valid) && ((((uintptr_t)Z | (uintptr_t)X | (uintptr_t)Y) & 15) == 0)) {
        if (n_valid <= 0) return B200_OK;
        const int V = is_f64 ? 2 : 4;
        const long long nvec = n_valid / V;
        const long long ctas = (nvec + 256 * KA_VEC_U - 1) / (256 * KA_VEC_U);
        const unsigned g = (unsigned)(ctas < 1 ? 1 : ctas);
        if (is_f64)
            saxpy_twin_vec_kernel<double><<<g, 256, 0, st>>>((double *)Z, a, (const double *)X, (const double *)Y, nvec, n_valid);
        else
            saxpy_twin_vec_kernel<float><<<g, 256, 0, st>>>((float *)Z, (float)a, (const float *)X, (const float *)Y, nvec, n_valid);
        B200_CHECK_LAUNCH("saxpy_twin_vec_kernel");
        return B200_OK;
    }
    if (is_f64)
        saxpy_twin_kernel<double><<<ka_grid, ka_block, 0, st>>>(c, (double *)Z, a, (const double *)X, (const double *)Y);
    else
        saxpy_twin_kernel<float><<<ka_grid, ka_block, 0, st>>>(c, (float *)Z, (float)a, (const float *)X, (const float *)Y);
    B200_CHECK_LAUNCH("saxpy_twin_kernel");
    return B200_OK;
}

// ---- examples/performance.jl:17-138: simple / local-memory / multi-element copy and transpose ----
// WHICH: 0 simple_copy, 1 simple_transpose (bounds-checked lanes); 2 lmem_copy, 3 lmem_transpose, 4 coalesced_copy,
// 5 coalesced_transpose (`unsafe_indices = true`: no lane guard, plain barrier -- src/macros.jl:72-76).
// tile = @localmem eltype(output) (N + BANK, M)  resp.  (TILE_DIM + BANK, TILE_DIM), column-major.
template <typename T, int WHICH>
__global__ void performance_twin_kernel(KaCtx c, T *__restrict__ output, const T *__restrict__ input,
                                        long long rows_out, long long rows_in, long long bank) {
    extern __shared__ __align__(16) unsigned char perf_tile_raw[];
    T *tile = reinterpret_cast<T *>(perf_tile_raw);
    if (WHICH <= 1) {
        if (ka_valid(c)) {
            long long I[KA_MAX_NDIM];
            ka_expand(c, I);
            if (WHICH == 0)
                output[(I[0] - 1) + (I[1] - 1) * rows_out] = input[(I[0] - 1) + (I[1] - 1) * rows_in];
            else
                output[(I[1] - 1) + (I[0] - 1) * rows_out] = input[(I[0] - 1) + (I[1] - 1) * rows_in];
        }
        return;
    }
    long long g[KA_MAX_NDIM], l[KA_MAX_NDIM];
    ka_group_cartesian(c, g);
    ka_local_cartesian(c, l);
    const long long gi = g[0], gj = g[1], i = l[0], j = l[1];
    const long long N = c.wgs[0], M = c.wgs[1];   // @uniform @groupsize()[1], [2]
    const long long trows = N + bank;
    if (WHICH == 2 || WHICH == 3) {
        long long I = (gi - 1) * N + i, J = (gj - 1) * M + j;
        tile[(i - 1) + (j - 1) * trows] = input[(I - 1) + (J - 1) * rows_in];
        __syncthreads();
        if (WHICH == 2) {
            output[(I - 1) + (J - 1) * rows_out] = tile[(i - 1) + (j - 1) * trows];
        } else {
            I = (gj - 1) * M + i;   // pivot the group index
            J = (gi - 1) * N + j;
            output[(I - 1) + (J - 1) * rows_out] = tile[(j - 1) + (i - 1) * trows];
        }
    } else {
        const long long TILE = N, ROWS = M;       // TILE_DIM, BLOCK_ROWS
        long long I = (gi - 1) * TILE + i, J = (gj - 1) * TILE + j;
        for (long long k = 0; k <= TILE - 1; k += ROWS)
            tile[(i - 1) + (j + k - 1) * trows] = input[(I - 1) + (J + k - 1) * rows_in];
        __syncthreads();
        if (WHICH == 4) {
            for (long long k = 0; k <= TILE - 1; k += ROWS)
                output[(I - 1) + (J + k - 1) * rows_out] = tile[(i - 1) + (j + k - 1) * trows];
        } else {
            I = (gj - 1) * TILE + i;   // transpose block offsets
            J = (gi - 1) * TILE + j;
            for (long long k = 0; k <= TILE - 1; k += ROWS)
                output[(I - 1) + (J + k - 1) * rows_out] = tile[(j + k - 1) + (i - 1) * trows];
        }
    }
}
template <typename T>
static int launch_performance(const KaCtx &c, int which, T *out, const T *in, long long rows_out, long long rows_in,
                              long long bank, long long groups, long long items, cudaStream_t st) {
    dim3 ka_grid, ka_block;
    ka_launch_dims(c, &ka_grid, &ka_block);
    size_t smem = 0;
    if (which >= 2) smem = (size_t)(c.wgs[0] + bank) * (size_t)(which >= 4 ? c.wgs[0] : c.wgs[1]) * sizeof(T);
    if (smem > 48 * 1024) return set_error(B200_ERR_LAUNCH_DIMS, "performance kernel: %zu bytes of @localmem", smem);
#define B200_PERF_CASE(W)                                                                                          \
    case W: performance_twin_kernel<T, W><<<ka_grid, ka_block, smem, st>>>(c, out, in, rows_out,    \
                                                                                           rows_in, bank); break;
    switch (which) {
        B200_PERF_CASE(0) B200_PERF_CASE(1) B200_PERF_CASE(2) B200_PERF_CASE(3) B200_PERF_CASE(4) B200_PERF_CASE(5)
        default: return set_error(B200_ERR_INVALID_VALUE, "performance kernel: variant %d", which);
    }
#undef B200_PERF_CASE
    B200_CHECK_LAUNCH("performance_twin_kernel");
    return B200_OK;
}
int twin_performance(const KaCtx &c, int which, void *out, const void *in, long long rows_out, long long rows_in,
                     long long bank, int elsize, cudaStream_t st) {
    KA_LAUNCH_1D(c, groups, items);
    if (elsize == 4)
        return launch_performance<uint32_t>(c, which, (uint32_t *)out, (const uint32_t *)in, rows_out, rows_in, bank, groups, items, st);
    if (elsize == 8)
        return launch_performance<uint64_t>(c, which, (uint64_t *)out, (const uint64_t *)in, rows_out, rows_in, bank, groups, items, st);
    return set_error(B200_ERR_UNSUPPORTED, "performance kernel: element size %d", elsize);
}

// ==================================================================================================
// KI twins: true <=3-D launches, no implicit bounds check (src/intrinsics.jl:300-373, pocl/backend.jl:156-205)
// ==================================================================================================

// histogram_kernel!  (examples/histogram.jl:18-58), literal: multi-pass over bins when N > gs
__global__ void histogram_twin_kernel(int32_t *__restrict__ histogram_output, long long N,
                                      const int32_t *__restrict__ input, long long len, int gs) {
    extern __shared__ int32_t shared_histogram[];  // KI.localmemory(eltype(input), gs)
    const long long gid = (long long)blockIdx.x + 1, lid = (long long)threadIdx.x + 1;
    const long long tid = (gid - 1) * gs + lid;
    for (long long min_element = 1; min_element <= N; min_element += gs) {
        shared_histogram[lid - 1] = 0;
        __syncthreads();
        long long max_element = min_element + gs;
        if (max_element > N) max_element = N + 1;
        long long bin = tid <= len ? (long long)input[tid - 1] : 0;
        if (bin >= min_element && bin < max_element) {
            bin -= min_element - 1;
            atomicAdd(&shared_histogram[bin - 1], 1);
        }
        __syncthreads();
        if (lid + min_element - 1 <= N) atomicAdd(&histogram_output[lid + min_element - 1 - 1], shared_histogram[lid - 1]);
    }
}
int twin_histogram(int32_t *out, long long N, const int32_t *input, long long len, long long gs, long long groups,
                   cudaStream_t st) {
    if (groups == 0) return B200_OK;
    if (gs < 1 || gs > 1024) return set_error(B200_ERR_LAUNCH_DIMS, "workgroup of %lld items exceeds 1024", gs);
    if (groups > 2147483647LL) return set_error(B200_ERR_LAUNCH_DIMS, "too many groups");
    histogram_twin_kernel<<<(unsigned)groups, (unsigned)gs, (size_t)gs * 4, st>>>(out, N, input, len, (int)gs);
    B200_CHECK_LAUNCH("histogram_twin_kernel");
    return B200_OK;
}

// coalesced_matmul_kernel!  (examples/performant_matmul.jl:15-77), literal 16x16-tile version (TDIM <= 32)
__global__ void coalesced_matmul_twin_kernel(float *__restrict__ output, const float *__restrict__ input1,
                                             const float *__restrict__ input2, long long N, long long R, long long M,
                                             int TDIM) {
    extern __shared__ float tiles[];  // two (TDIM+1, TDIM) column-major tiles
    float *tile1 = tiles, *tile2 = tiles + (TDIM + 1) * TDIM;
    const long long gi = blockIdx.x + 1, gj = blockIdx.y + 1;
    const int i = threadIdx.x + 1, j = threadIdx.y + 1;
    float outval = -0.0f;
    const long long NUM_TILES = (R + TDIM - 1) / TDIM;
    const long long I = (gi - 1) * TDIM + i, J = (gj - 1) * TDIM + j;
    for (long long t = 0; t < NUM_TILES; ++t) {
        tile1[(i - 1) + (j - 1) * (TDIM + 1)] =
            (I <= N && t * TDIM + j <= R) ? input1[(I - 1) + (t * TDIM + j - 1) * N] : 0.0f;
        tile2[(i - 1) + (j - 1) * (TDIM + 1)] =
            (t * TDIM + i <= R && J <= M) ? input2[(t * TDIM + i - 1) + (J - 1) * R] : 0.0f;
        __syncthreads();
        float out = 0.0f;
        for (int k = 1; k <= TDIM; ++k)
            out = fmaf(tile1[(i - 1) + (k - 1) * (TDIM + 1)], tile2[(k - 1) + (j - 1) * (TDIM + 1)], out);
        outval += out;
        __syncthreads();
    }
    if (I <= N && J <= M) output[(I - 1) + (J - 1) * N] = outval;
}
int twin_coalesced_matmul(float *output, const float *in1, const float *in2, long long N, long long R, long long M,
                          int TDIM, long long groups_x, long long groups_y, cudaStream_t st) {
    if (groups_x == 0 || groups_y == 0) return B200_OK;
    if (TDIM < 1 || TDIM > 32) return set_error(B200_ERR_LAUNCH_DIMS, "coalesced_matmul_kernel!: TDIM %d not in 1..32", TDIM);
    if (groups_y > 65535 || groups_x > 2147483647LL) return set_error(B200_ERR_LAUNCH_DIMS, "too many groups");
    dim3 grid((unsigned)groups_x, (unsigned)groups_y), block(TDIM, TDIM);
    size_t smem = (size_t)2 * (TDIM + 1) * TDIM * sizeof(float);
    coalesced_matmul_twin_kernel<<<grid, block, smem, st>>>(output, in1, in2, N, R, M, TDIM);
    B200_CHECK_LAUNCH("coalesced_matmul_twin_kernel");
    return B200_OK;
}

// test_intrinsics_kernel  (test/intrinsics.jl:11-25): six Int64 per element, .x components, 1-based
__global__ void test_intrinsics_twin_kernel(long long *__restrict__ results, long long len) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i <= len) {
        long long *r = results + (i - 1) * 6;
        r[0] = (long long)gridDim.x * blockDim.x;
        r[1] = i;
        r[2] = blockDim.x;
        r[3] = threadIdx.x + 1;
        r[4] = gridDim.x;
        r[5] = blockIdx.x + 1;
    }
}
// launch_kernel1d/2d/3d  (test/intrinsics.jl:31-72)
__global__ void launch_kernel_nd_twin_kernel(float *__restrict__ arr, long long d0, long long d1) {
    const long long i = threadIdx.x + 1, j = threadIdx.y + 1, k = threadIdx.z + 1;
    const long long gi = blockIdx.x + 1, gj = blockIdx.y + 1, gk = blockIdx.z + 1;
    const long long ngi = gridDim.x, ngj = gridDim.y, ngk = gridDim.z;
    const long long x = (gi - 1) * ngi + i, y = (gj - 1) * ngj + j, z = (gk - 1) * ngk + k;
    arr[(x - 1) + (y - 1) * d0 + (z - 1) * d0 * d1] = 1.0f;
}
int twin_test_intrinsics(long long *results, long long len, const long long *groups, const long long *wgs,
                         cudaStream_t st) {
    if (groups[0] * groups[1] * groups[2] == 0) return B200_OK;
    dim3 grid((unsigned)groups[0], (unsigned)groups[1], (unsigned)groups[2]);
    dim3 block((unsigned)wgs[0], (unsigned)wgs[1], (unsigned)wgs[2]);
    test_intrinsics_twin_kernel<<<grid, block, 0, st>>>(results, len);
    B200_CHECK_LAUNCH("test_intrinsics_twin_kernel");
    return B200_OK;
}
int twin_launch_kernel_nd(float *arr, const long long *dims, const long long *groups, const long long *wgs,
                          cudaStream_t st) {
    if (groups[0] * groups[1] * groups[2] == 0) return B200_OK;
    dim3 grid((unsigned)groups[0], (unsigned)groups[1], (unsigned)groups[2]);
    dim3 block((unsigned)wgs[0], (unsigned)wgs[1], (unsigned)wgs[2]);
    launch_kernel_nd_twin_kernel<<<grid, block, 0, st>>>(arr, dims[0], dims[1]);
    B200_CHECK_LAUNCH("launch_kernel_nd_twin_kernel");
    return B200_OK;
}

}  // namespace b200
