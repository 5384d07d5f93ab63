// arrayops.cu -- what a KA user gets from Base/LinearAlgebra for free on the CPU() backend, for device arrays:
//   maximum / minimum / sum   (examples/histogram.jl:10,88-90: the histogram is sized by `maximum(input)`)
//   A == B                    (examples/memcopy.jl:23, memcopy_static.jl:23, naive_transpose.jl:32, test/copyto.jl:17-19)
//   isapprox(A, B)            (examples/matmul.jl:36, performant_matmul.jl:92, histogram.jl:97-99): norm(A-B) <= rtol*max(norm A, norm B)
//   rand-like fills           (the deterministic input streams of SURVEY.md 8(d), the same counter hash the test oracle uses)
// so the examples run on B200Arrays without a round trip through host memory.
//
// All of them stream their operands once: HBM-bound, n * elsize bytes per operand.  One persistent grid (8 CTAs of 256
// threads per SM), 16-byte ld.global.nc loads with four in flight per thread, warp-shuffle + shared-memory block
// reduction, per-block partials in a per-thread scratch buffer; the block that takes the last ticket folds the partials
// and stores the result (<= 24 bytes) straight into a pinned host landing pad (zero-copy), so a call is ONE launch plus
// one stream wait -- these calls return a VALUE, so they are the one place where the library waits for the stream.
#include <cmath>
#include <cstring>
#include <type_traits>

#include "../common.cuh"

namespace b200 {
namespace ao {

constexpr int THREADS = 256, CTAS_PER_SM = 8, MAX_BLOCKS = 148 * CTAS_PER_SM * 2;

enum { OP_SUM = 0, OP_MAX = 1, OP_MIN = 2 };

template <typename V>
__device__ __forceinline__ V ld_nc16(const V *p) {
    static_assert(sizeof(V) == 16, "16-byte vector");
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return *reinterpret_cast<V *>(&r);
}

// Unaligned views (B200Array.view, Base.view): `head` leading elements are handled one by one, the 128-bit loop starts at the
// first 16-byte boundary; two operands vectorise together only when they are equally misaligned, else everything is scalar.
template <typename T>
__device__ __forceinline__ size_t head_elems(const T *a, const T *b, size_t n, bool *vec_ok) {
    const uintptr_t ua = (uintptr_t)a, ub = b ? (uintptr_t)b : ua;
    *vec_ok = ((ua ^ ub) & 15) == 0 && (ua % sizeof(T)) == 0;
    if (!*vec_ok) return n;
    const size_t h = ((16 - (ua & 15)) & 15) / sizeof(T);
    return h < n ? h : n;
}

// ---- combine rules (Julia semantics: NaN propagates through maximum/minimum; max(-0.0, 0.0) == 0.0) ----
template <typename T>
__device__ __forceinline__ T jl_max(T a, T b) { return a > b ? a : b; }
template <typename T>
__device__ __forceinline__ T jl_min(T a, T b) { return a < b ? a : b; }
template <>
__device__ __forceinline__ float jl_max<float>(float a, float b) {   // one instruction: NaN-propagating, -0.0 < +0.0
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <>
__device__ __forceinline__ double jl_max<double>(double a, double b) {
    if (isnan(a) || isnan(b)) return a + b;
    return a == b ? (signbit(a) ? b : a) : (a > b ? a : b);
}
template <>
__device__ __forceinline__ float jl_min<float>(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <>
__device__ __forceinline__ double jl_min<double>(double a, double b) {
    if (isnan(a) || isnan(b)) return a + b;
    return a == b ? (signbit(a) ? a : b) : (a < b ? a : b);
}

template <typename T> struct SumAcc { using type = long long; };      // sum(::Array{Int32}) widens to Int64 in Julia
template <> struct SumAcc<float> { using type = double; };            // floating sums are carried in Float64
template <> struct SumAcc<double> { using type = double; };
template <> struct SumAcc<unsigned int> { using type = unsigned long long; };
template <> struct SumAcc<unsigned long long> { using type = unsigned long long; };

// identities: 0 for sum, the lowest / highest value of the type for maximum / minimum (-Inf / +Inf for floats, so NaN and signed
// zeros still come out by Julia's rules)
template <typename T> __device__ __forceinline__ T lowest_of() { return std::is_signed<T>::value ? (T)((T)1 << (sizeof(T) * 8 - 1)) : (T)0; }
template <> __device__ __forceinline__ float lowest_of<float>() { return -__int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double lowest_of<double>() { return -__longlong_as_double(0x7ff0000000000000LL); }
template <typename T> __device__ __forceinline__ T highest_of() { return std::is_signed<T>::value ? (T)~((T)1 << (sizeof(T) * 8 - 1)) : (T)~(T)0; }
template <> __device__ __forceinline__ float highest_of<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double highest_of<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

template <typename T, int OP>
struct Red {
    using Acc = typename std::conditional<OP == OP_SUM, typename SumAcc<T>::type, T>::type;
    __device__ static Acc identity() {
        if (OP == OP_SUM) return (Acc)0;
        if (OP == OP_MAX) return (Acc)lowest_of<T>();
        return (Acc)highest_of<T>();
    }
    __device__ static Acc lift(T v) { return (Acc)v; }
    __device__ static Acc comb(Acc a, Acc b) {
        if (OP == OP_SUM) return a + b;
        if (OP == OP_MAX) return jl_max<Acc>(a, b);
        return jl_min<Acc>(a, b);
    }
};

template <typename Acc, typename F>
__device__ __forceinline__ Acc block_reduce(Acc v, F comb, Acc *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Acc other;
        if (sizeof(Acc) == 8) {
            long long t = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<long long *>(&v), o);
            other = *reinterpret_cast<Acc *>(&t);
        } else {
            int t = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<int *>(&v), o);
            other = *reinterpret_cast<Acc *>(&t);
        }
        v = comb(v, other);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = sh[lane < (int)(blockDim.x >> 5) ? lane : 0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Acc other;
            if (sizeof(Acc) == 8) {
                long long t = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<long long *>(&v), o);
                other = *reinterpret_cast<Acc *>(&t);
            } else {
                int t = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<int *>(&v), o);
                other = *reinterpret_cast<Acc *>(&t);
            }
            v = comb(v, other);   // lanes >= nwarps hold a copy of sh[0]: harmless for max/min, fixed up for sum below
        }
    }
    __syncthreads();
    return v;
}


// Grid-wide finish: every block publishes Q partial values, the block that draws the last ticket folds all of them and
// writes the Q results to `result` (a device-visible pinned host buffer), then re-arms the ticket for the next call.
template <typename Acc, int Q, typename F>
__device__ __forceinline__ void finish(const Acc (&v)[Q], Acc *partials, unsigned int *ticket, Acc *result, F comb, Acc identity) {
    __shared__ bool last;
    __shared__ Acc fold[Q][THREADS];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < Q; ++q) partials[Q * blockIdx.x + q] = v[q];
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    Acc acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = identity;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[q] = comb(acc[q], __ldcg(partials + Q * i + q));
#pragma unroll
    for (int q = 0; q < Q; ++q) fold[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
#pragma unroll
            for (int q = 0; q < Q; ++q) fold[q][threadIdx.x] = comb(fold[q][threadIdx.x], fold[q][threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < Q; ++q) result[q] = fold[q][0];
        __threadfence_system();
        *ticket = 0;
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(THREADS) reduce_kernel(const T *__restrict__ x, size_t n, typename Red<T, OP>::Acc *partials, unsigned int *ticket,
                                                         typename Red<T, OP>::Acc *result) {
    using R = Red<T, OP>;
    using Acc = typename R::Acc;
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) V { T e[VEC]; };
    bool vec_ok;
    const size_t head = head_elems<T>(x, nullptr, n, &vec_ok);
    const size_t nvec = (n - head) / VEC, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    Acc acc = R::identity();
    const V *xv = reinterpret_cast<const V *>(x + head);
    for (size_t j = tid; j < head; j += stride) acc = R::comb(acc, R::lift(x[j]));
    size_t i = tid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        V v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_nc16(xv + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc = R::comb(acc, R::lift(v[u].e[e]));
    }
    for (; i < nvec; i += stride) {
        V v = ld_nc16(xv + i);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc = R::comb(acc, R::lift(v.e[e]));
    }
    for (size_t j = head + nvec * VEC + tid; j < n; j += stride) acc = R::comb(acc, R::lift(x[j]));
    __shared__ Acc sh[32];
    Acc blk[1];
    if (OP == OP_SUM) {   // lanes beyond the warp count must not contribute twice: plain two-level sum
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long t = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<long long *>(&acc), o);
            acc = acc + *reinterpret_cast<Acc *>(&t);
        }
        if (lane == 0) sh[warp] = acc;
        __syncthreads();
        Acc sum = 0;
        if (threadIdx.x == 0)
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum = sum + sh[w];
        blk[0] = sum;
    } else {
        blk[0] = block_reduce<Acc>(acc, [](Acc a, Acc b) { return R::comb(a, b); }, sh);
    }
    finish<Acc, 1>(blk, partials, ticket, result, [](Acc a, Acc b) { return R::comb(a, b); }, R::identity());
}
// ---- A == B: number of positions where !(a == b) (so NaN never equals, -0.0 == 0.0), bit patterns for opaque types ----
template <typename T>
__global__ void __launch_bounds__(THREADS) mismatch_kernel(const T *__restrict__ a, const T *__restrict__ b, size_t n,
                                                           unsigned long long *partials, unsigned int *ticket,
                                                           unsigned long long *result) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) V { T e[VEC]; };
    bool vec_ok;
    const size_t head = head_elems<T>(a, b, n, &vec_ok);
    const size_t nvec = (n - head) / VEC, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    const V *av = reinterpret_cast<const V *>(a + head), *bv = reinterpret_cast<const V *>(b + head);
    for (size_t j = tid; j < head; j += stride) bad += !(a[j] == b[j]);
    size_t i = tid;
    for (; i + stride < nvec; i += 2 * stride) {
        V x0 = ld_nc16(av + i), y0 = ld_nc16(bv + i), x1 = ld_nc16(av + i + stride), y1 = ld_nc16(bv + i + stride);
#pragma unroll
        for (int e = 0; e < VEC; ++e) bad += !(x0.e[e] == y0.e[e]) + !(x1.e[e] == y1.e[e]);
    }
    for (; i < nvec; i += stride) {
        V x0 = ld_nc16(av + i), y0 = ld_nc16(bv + i);
#pragma unroll
        for (int e = 0; e < VEC; ++e) bad += !(x0.e[e] == y0.e[e]);
    }
    for (size_t j = head + nvec * VEC + tid; j < n; j += stride) bad += !(a[j] == b[j]);
    __shared__ unsigned long long sh[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = bad;
    __syncthreads();
    unsigned long long blk[1] = {0};
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) blk[0] += sh[w];
    finish<unsigned long long, 1>(blk, partials, ticket, result, [](unsigned long long x, unsigned long long y) { return x + y; }, 0ull);
}

// ---- isapprox: sum (x-y)^2, sum x^2, sum y^2 in Float64 ----
template <typename T>
__global__ void __launch_bounds__(THREADS) norms_kernel(const T *__restrict__ a, const T *__restrict__ b, size_t n, double *partials,
                                                        unsigned int *ticket, double *result) {
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) V { T e[VEC]; };
    bool vec_ok;
    const size_t head = head_elems<T>(a, b, n, &vec_ok);
    const size_t nvec = (n - head) / VEC, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double sd = 0, sx = 0, sy = 0;
    const V *av = reinterpret_cast<const V *>(a + head), *bv = reinterpret_cast<const V *>(b + head);
    auto add = [&](T xe, T ye) {
        const double x = (double)xe, y = (double)ye, d = x - y;
        sd = fma(d, d, sd);
        sx = fma(x, x, sx);
        sy = fma(y, y, sy);
    };
    size_t i = tid;
    for (; i + stride < nvec; i += 2 * stride) {
        V x0 = ld_nc16(av + i), y0 = ld_nc16(bv + i), x1 = ld_nc16(av + i + stride), y1 = ld_nc16(bv + i + stride);
#pragma unroll
        for (int e = 0; e < VEC; ++e) add(x0.e[e], y0.e[e]);
#pragma unroll
        for (int e = 0; e < VEC; ++e) add(x1.e[e], y1.e[e]);
    }
    for (; i < nvec; i += stride) {
        V x0 = ld_nc16(av + i), y0 = ld_nc16(bv + i);
#pragma unroll
        for (int e = 0; e < VEC; ++e) add(x0.e[e], y0.e[e]);
    }
    for (size_t j = tid; j < head; j += stride) add(a[j], b[j]);
    for (size_t j = head + nvec * VEC + tid; j < n; j += stride) add(a[j], b[j]);
    __shared__ double sh[3][32];
    double v[3] = {sd, sx, sy};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    double blk[3] = {0, 0, 0};
    if (threadIdx.x == 0)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) blk[q] += sh[q][w];
    finish<double, 3>(blk, partials, ticket, result, [](double x, double y) { return x + y; }, 0.0);
}
// ---- deterministic input streams: a splitmix-style counter hash (mix32), four elements per 16-byte store ----
__device__ __forceinline__ uint32_t mix32(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (uint32_t)(x >> 32);
}
template <bool BINS>
__global__ void __launch_bounds__(THREADS) gen_kernel(uint32_t *__restrict__ out, size_t n, unsigned long long base, uint32_t nbins, uint32_t recip) {
    const size_t head0 = ((16 - ((uintptr_t)out & 15)) & 15) / 4, head = head0 < n ? head0 : n;
    const size_t nvec = (n - head) / 4, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    auto one = [&](size_t i) -> uint32_t {
        const uint32_t h = mix32(base + (unsigned long long)i);
        if (BINS && nbins == 1u) return 1u;
        if (BINS) {   // h % nbins without a divide: recip = floor(2^32 / nbins) underestimates the quotient by at most 2
            uint32_t r = h - __umulhi(h, recip) * nbins;
            r -= r >= nbins ? nbins : 0u;
            r -= r >= nbins ? nbins : 0u;
            return r + 1u;
        }
        return __float_as_uint((float)(h >> 8) * (1.0f / 16777216.0f));
    };
    for (size_t j = tid; j < head; j += stride) out[j] = one(j);
    for (size_t v = tid; v < nvec; v += stride) {
        const size_t e = head + 4 * v;
        reinterpret_cast<uint4 *>(out + head)[v] = make_uint4(one(e), one(e + 1), one(e + 2), one(e + 3));
    }
    for (size_t j = head + nvec * 4 + tid; j < n; j += stride) out[j] = one(j);
}

// ---- per-thread, per-device scratch: MAX_BLOCKS * 24 bytes of partials, a ticket, and a pinned 32-byte landing pad that
// ---- the device writes directly (UVA: the pinned host pointer is valid on the device) ----
struct Scratch {
    void *dev = nullptr;             // partials
    unsigned int *ticket = nullptr;  // zero between calls (re-armed by the finishing block)
    void *pinned = nullptr;
};
static thread_local Scratch g_scratch[kMaxDevices];

static int scratch(Scratch **out) {
    B200_TRY(ensure_device());
    Scratch &s = g_scratch[tls().device];
    if (!s.dev) {
        B200_CUDA(cudaMalloc(&s.dev, (size_t)MAX_BLOCKS * 24 + 16));
        s.ticket = (unsigned int *)((char *)s.dev + (size_t)MAX_BLOCKS * 24);
        B200_CUDA(cudaMemset(s.ticket, 0, 16));
        B200_CUDA(cudaHostAlloc(&s.pinned, 32, cudaHostAllocMapped | cudaHostAllocPortable));
    }
    *out = &s;
    return B200_OK;
}
static int grid_for(size_t nvec) {
    int64_t blocks = cdiv((int64_t)(nvec ? nvec : 1), THREADS * 4);
    int64_t cap = (int64_t)sm_count() * CTAS_PER_SM;
    if (cap > MAX_BLOCKS) cap = MAX_BLOCKS;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}
static int fetch(Scratch *s, void *out_host, size_t bytes, cudaStream_t st) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaMemset(s->ticket, 0, 16);   // a failed launch must not leave a half-drawn ticket behind
        return set_error(B200_ERR_CUDA, "array reduction failed: %s", cudaGetErrorString(e));
    }
    memcpy(out_host, s->pinned, bytes);
    return B200_OK;
}

template <typename T, int OP>
static int reduce_typed(const void *x, size_t n, void *out_host, cudaStream_t st) {
    using Acc = typename Red<T, OP>::Acc;
    Scratch *s;
    B200_TRY(scratch(&s));
    const int grid = grid_for(n / (16 / sizeof(T)));
    reduce_kernel<T, OP><<<grid, THREADS, 0, st>>>((const T *)x, n, (Acc *)s->dev, s->ticket, (Acc *)s->pinned);
    B200_CHECK_LAUNCH("reduce_kernel");
    return fetch(s, out_host, sizeof(Acc), st);
}
template <typename T>
static int reduce_op(int op, const void *x, size_t n, void *out_host, cudaStream_t st) {
    if (op == OP_SUM) return reduce_typed<T, OP_SUM>(x, n, out_host, st);
    if (op == OP_MAX) return reduce_typed<T, OP_MAX>(x, n, out_host, st);
    return reduce_typed<T, OP_MIN>(x, n, out_host, st);
}

}  // namespace ao
}  // namespace b200

using namespace b200;

extern "C" b200_status b200_reduce(int32_t op, int32_t eltype, const void *x, size_t n, void *result_host) {
    if (!result_host) return set_error(B200_ERR_INVALID_VALUE, "b200_reduce: result pointer is NULL");
    if (op < ao::OP_SUM || op > ao::OP_MIN) return set_error(B200_ERR_INVALID_VALUE, "b200_reduce: unknown op %d", op);
    if (n == 0) {
        if (op != ao::OP_SUM)   // Julia: maximum(Int32[]) throws ArgumentError("reducing over an empty collection is not allowed")
            return set_error(B200_ERR_INVALID_VALUE, "reducing over an empty collection is not allowed");
        memset(result_host, 0, 8);
        return B200_OK;
    }
    if (!x) return set_error(B200_ERR_INVALID_VALUE, "b200_reduce: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    switch (eltype) {
        case B200_T_I32: return ao::reduce_op<int>(op, x, n, result_host, st);
        case B200_T_U32: return ao::reduce_op<unsigned int>(op, x, n, result_host, st);
        case B200_T_I64: return ao::reduce_op<long long>(op, x, n, result_host, st);
        case B200_T_U64: return ao::reduce_op<unsigned long long>(op, x, n, result_host, st);
        case B200_T_F32: return ao::reduce_op<float>(op, x, n, result_host, st);
        case B200_T_F64: return ao::reduce_op<double>(op, x, n, result_host, st);
        default: return set_error(B200_ERR_UNSUPPORTED, "b200_reduce: element type %d not supported", eltype);
    }
}

extern "C" b200_status b200_count_mismatch(const void *a, const void *b, size_t n, int32_t eltype, int32_t elsize,
                                           int64_t *count_host) {
    if (!count_host) return set_error(B200_ERR_INVALID_VALUE, "b200_count_mismatch: result pointer is NULL");
    *count_host = 0;
    if (n == 0) return B200_OK;
    if (!a || !b) return set_error(B200_ERR_INVALID_VALUE, "b200_count_mismatch: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    ao::Scratch *s;
    B200_TRY(ao::scratch(&s));
    unsigned long long *part = (unsigned long long *)s->dev;
    int grid;
    if (eltype == B200_T_F32) {
        grid = ao::grid_for(n / 4);
        ao::mismatch_kernel<float><<<grid, ao::THREADS, 0, st>>>((const float *)a, (const float *)b, n, part, s->ticket, (unsigned long long *)s->pinned);
    } else if (eltype == B200_T_F64) {
        grid = ao::grid_for(n / 2);
        ao::mismatch_kernel<double><<<grid, ao::THREADS, 0, st>>>((const double *)a, (const double *)b, n, part, s->ticket, (unsigned long long *)s->pinned);
    } else {   // integers, Bool, opaque isbits: equality of the bit patterns, compared in the widest unit that divides elsize
        if (elsize <= 0) return set_error(B200_ERR_INVALID_VALUE, "b200_count_mismatch: bad element size");
        const size_t bytes = n * (size_t)elsize;
        if (elsize % 8 == 0) {
            grid = ao::grid_for(bytes / 16);
            ao::mismatch_kernel<unsigned long long><<<grid, ao::THREADS, 0, st>>>((const unsigned long long *)a, (const unsigned long long *)b, bytes / 8, part, s->ticket, (unsigned long long *)s->pinned);
        } else if (elsize % 4 == 0) {
            grid = ao::grid_for(bytes / 16);
            ao::mismatch_kernel<unsigned int><<<grid, ao::THREADS, 0, st>>>((const unsigned int *)a, (const unsigned int *)b, bytes / 4, part, s->ticket, (unsigned long long *)s->pinned);
        } else {
            grid = ao::grid_for(bytes / 16);
            ao::mismatch_kernel<unsigned char><<<grid, ao::THREADS, 0, st>>>((const unsigned char *)a, (const unsigned char *)b, bytes, part, s->ticket, (unsigned long long *)s->pinned);
        }
    }
    B200_CHECK_LAUNCH("mismatch_kernel");
    (void)grid;
    return ao::fetch(s, count_host, 8, st);
}

extern "C" b200_status b200_norms(const void *x, const void *y, size_t n, int32_t is_f64, double *out3_host) {
    if (!out3_host) return set_error(B200_ERR_INVALID_VALUE, "b200_norms: result pointer is NULL");
    out3_host[0] = out3_host[1] = out3_host[2] = 0.0;
    if (n == 0) return B200_OK;
    if (!x || !y) return set_error(B200_ERR_INVALID_VALUE, "b200_norms: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    ao::Scratch *s;
    B200_TRY(ao::scratch(&s));
    const int grid = ao::grid_for(n / (is_f64 ? 2 : 4));
    if (is_f64)
        ao::norms_kernel<double><<<grid, ao::THREADS, 0, st>>>((const double *)x, (const double *)y, n, (double *)s->dev, s->ticket, (double *)s->pinned);
    else
        ao::norms_kernel<float><<<grid, ao::THREADS, 0, st>>>((const float *)x, (const float *)y, n, (double *)s->dev, s->ticket, (double *)s->pinned);
    B200_CHECK_LAUNCH("norms_kernel");
    return ao::fetch(s, out3_host, 24, st);
}

extern "C" b200_status b200_rand_uniform_f32(float *out, size_t n, uint64_t seed) {
    if (n == 0) return B200_OK;
    if (!out) return set_error(B200_ERR_INVALID_VALUE, "b200_rand_uniform_f32: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    ao::gen_kernel<false><<<ao::grid_for(n / 4), ao::THREADS, 0, st>>>((uint32_t *)out, n, seed * 0x100000001B3ull, 1u, 0u);
    B200_CHECK_LAUNCH("gen_kernel<uniform>");
    return B200_OK;
}

extern "C" b200_status b200_rand_bins_i32(int32_t *out, size_t n, uint64_t seed, int32_t nbins) {
    if (nbins <= 0) return set_error(B200_ERR_INVALID_VALUE, "b200_rand_bins_i32: nbins must be positive");
    if (n == 0) return B200_OK;
    if (!out) return set_error(B200_ERR_INVALID_VALUE, "b200_rand_bins_i32: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    ao::gen_kernel<true><<<ao::grid_for(n / 4), ao::THREADS, 0, st>>>((uint32_t *)out, n, seed * 0x100000001B3ull, (uint32_t)nbins,
                                                                 (uint32_t)(0x100000000ull / (uint32_t)nbins));
    B200_CHECK_LAUNCH("gen_kernel<bins>");
    return B200_OK;
}
