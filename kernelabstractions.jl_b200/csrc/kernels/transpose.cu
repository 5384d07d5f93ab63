// transpose.cu -- b200_transpose: a[i,j] = b[j,i]  (reference examples/naive_transpose.jl:4-7).
//
// The reference kernel is "naive": with workgroup (256,1) the stores to `a` are coalesced and the loads
// of `b` are strided by a whole column (64 KiB at 16384^2), i.e. one 32-byte sector per 4-byte element.
// This file computes exactly the same result (a pure permutation of 4/8-byte words: bit-exact) with a
// shared-memory tile so that BOTH sides move 128-bit, fully coalesced:
//
//   fast path (4-byte elements, m,n multiples of 64, 16-byte aligned rows):
//     64x64 tile per CTA, 256 threads.  Each thread loads four float4 along b's contiguous dimension from
//     four consecutive columns -> a 4x4 block in registers, transposed for free by renaming registers,
//     written as four float4 into a [64][64] shared tile whose 16-byte chunks are XOR-swizzled by (row>>2)
//     (conflict-free for both the write and the read pattern), then read back as float4 along a's
//     contiguous dimension and stored.  Per 16 bytes moved: 1 LDG.128 + 1 STS.128 + 1 LDS.128 + 1 STG.128.
//   any-shape path (4-byte elements, any m,n; 16-byte aligned bases, leading dimensions multiples of 4 words): the same
//     scheme; edge tiles are predicated, interior tiles are not (6.2 TB/s at 10000 x 7000, 6.0 at 16388^2).
//   generic path (1-, 2-, 8-, 16-byte elements, unaligned rows): 32x32 padded tile, scalar accesses, predicated edges (6.6 TB/s for
//     Float64: 32 lanes x 8 bytes are full 256-byte requests already; 4.6 TB/s for 4-byte words).
//
// HBM bound: 2*elsize bytes of traffic per element; algorithmic bytes for 16384^2 F32 = 2^31.
#include "../common.cuh"
#include "../tma_pipe.cuh"
#include "../tune.h"

namespace b200 {

__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f4(float *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// a: m x n (lda), b: n x m (ldb).  Tile (ti, tj) covers i in [64 ti, 64 ti + 64), j in [64 tj, 64 tj + 64).
// ORDER 0: consecutive CTAs walk j (b's contiguous dimension) first; ORDER 1: i (a's contiguous dim) first.
template <int ORDER>
__global__ void __launch_bounds__(256) transpose64_f32_kernel(float *__restrict__ a, const float *__restrict__ b,
                                                              int tiles_i, int tiles_j, int64_t lda, int64_t ldb, int prefetch) {
    __shared__ __align__(16) float tile[64 * 64];  // [j][i], 16-byte chunk index XOR ((j >> 2) & 15)
    const int t = threadIdx.x;
    // Programmatic dependent launch (launch attribute set by transpose_launch): the next PDL-aware kernel of the stream may
    // become resident while this grid drains; nothing here touches global memory before the previous grid has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int ti, tj;
    if (ORDER == 0) {
        tj = blockIdx.x % tiles_j;
        ti = blockIdx.x / tiles_j;
    } else {
        ti = blockIdx.x % tiles_i;
        tj = blockIdx.x / tiles_i;
    }
    const int64_t i0 = (int64_t)ti * 64, j0 = (int64_t)tj * 64;
    // The first `prefetch` CTAs -- the ones that become resident while the previous kernel of the stream drains -- pull their tile
    // (64 columns x 256 bytes) into L2 before the dependency wait; a prefetch has no architectural effect (L2 is the point of
    // coherence).  Measured (profiles/r05_slab_probe.txt, leading 16 x 148 CTAs): 44.3 -> 43.0 us on the 8-GPU slab (8192 tiles),
    // 85.7 -> 84.7 on the 4-GPU slab, but 327.8 -> 335.5 us at 16384^2 (and 338 when every CTA does it): a long launch keeps HBM
    // busy to its very end, so the prefetches only compete.  Hence: launches of <= 16384 tiles only (`transpose.l2_prefetch`).
    if (blockIdx.x < (unsigned)prefetch && t < 128) {   // `prefetch` = number of leading CTAs that do it (0: none)
        const float *line = b + j0 + (i0 + (t >> 1)) * ldb + (t & 1) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- load: thread owns j = 4*jq .. 4*jq+3 and i = 4*iq .. 4*iq+3
    const int jq = t & 15, iq = t >> 4;
    const float *src = b + (j0 + 4 * jq) + (i0 + 4 * iq) * ldb;
    float4 v0 = ldg_stream_f4(src);
    float4 v1 = ldg_stream_f4(src + ldb);
    float4 v2 = ldg_stream_f4(src + 2 * ldb);
    float4 v3 = ldg_stream_f4(src + 3 * ldb);
    // v_r.{x,y,z,w} = b[j+{0,1,2,3}, i+r]  ->  row (j+c) of the tile gets (v0.c, v1.c, v2.c, v3.c)
    float4 *tile4 = reinterpret_cast<float4 *>(tile);
    const int chunk = iq ^ jq;  // (row >> 2) == jq for the four rows 4*jq + c
    tile4[(4 * jq + 0) * 16 + chunk] = make_float4(v0.x, v1.x, v2.x, v3.x);
    tile4[(4 * jq + 1) * 16 + chunk] = make_float4(v0.y, v1.y, v2.y, v3.y);
    tile4[(4 * jq + 2) * 16 + chunk] = make_float4(v0.z, v1.z, v2.z, v3.z);
    tile4[(4 * jq + 3) * 16 + chunk] = make_float4(v0.w, v1.w, v2.w, v3.w);
    __syncthreads();
    // ---- store: thread owns i-chunk ic (i = 4*ic..4*ic+3) of rows j = jr, jr+16, jr+32, jr+48
    const int ic = t & 15, jr = t >> 4;
    float *dst = a + (i0 + 4 * ic) + (j0 + jr) * lda;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int j = jr + 16 * s;
        float4 w = tile4[j * 16 + (ic ^ ((j >> 2) & 15))];
        stg_stream_f4(dst + (int64_t)(16 * s) * lda, w);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// TMA variant of the fast path (`transpose.impl = 1`; A/B against the kernel above in profiles/r05_transpose_ab.txt).
// The load side goes to the TMA engine: a 3-d tensor map over b with the columns i of a tile PERMUTED on the way in --
// dimensions (j: contiguous, i_hi = i / 4: stride 4 ldb, i_lo = i % 4: stride ldb), box 32 x 16 x 4 -- so that shared-memory
// row r = i_lo * 16 + i_hi holds the 128-byte segment b[j0 .. j0+32, i0 + 4 i_hi + i_lo], with the hardware's 128-byte swizzle
// (16-byte chunk ^= row & 7).  Lane (ic = t & 15, jq = t >> 4) then reads its 4 x 4 block as four LDS.128 at rows i_lo * 16 + ic:
// sixteen CONSECUTIVE rows per instruction, i.e. all eight swizzle phases -- conflict free (with the natural row order the
// rows of a 4 x 4 register block are 4 apart and the hardware swizzle, which only knows row & 7, leaves an 8-way conflict).
// After the register transpose the same lane mapping makes the stores coalesced: 16 lanes x float4 = 256 contiguous bytes of
// one column of a.  Per 16 bytes: TMA + 1 LDS.128 + 1 STG.128 (the SIMT kernel: LDG + STS + LDS + STG), no second pass
// through shared memory; persistent CTAs keep STAGES tiles of loads in flight behind an mbarrier ring.
// ---------------------------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(256) transpose64_tma_kernel(const __grid_constant__ CUtensorMap map_b, float *__restrict__ a, int tiles_i,
                                                              int tiles_j, int64_t lda) {
    extern __shared__ __align__(1024) unsigned char tma_smem_raw[];   // STAGES x (2 boxes x 64 rows x 128 B) + 1 KiB of slack
    __shared__ __align__(8) uint64_t full[STAGES];
    // the 128-byte swizzle is a function of the shared-memory ADDRESS: boxes must start on 1 KiB boundaries
    unsigned char *tma_smem = tma_smem_raw + ((1024u - (pipe::smem_u32(tma_smem_raw) & 1023u)) & 1023u);
    const int t = threadIdx.x;
    const unsigned ntiles = (unsigned)tiles_i * (unsigned)tiles_j;
    const unsigned mine = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_of = [&](unsigned k, int &ti, int &tj) {           // consecutive tiles walk j (b's contiguous dimension) first
        const unsigned id = blockIdx.x + k * gridDim.x;
        tj = (int)(id % (unsigned)tiles_j);
        ti = (int)(id / (unsigned)tiles_j);
    };
    auto issue = [&](unsigned k) {                                // thread 0: both 32-column boxes of tile k
        int ti, tj;
        tile_of(k, ti, tj);
        const int s = (int)(k % STAGES);
        unsigned char *dst = tma_smem + (size_t)s * 16384;
        pipe::mbar_expect_tx(&full[s], 16384);
#pragma unroll
        for (int h = 0; h < 2; ++h)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                             "r"(pipe::smem_u32(dst + h * 8192)), "l"(&map_b), "r"(tj * 64 + h * 32), "r"(ti * 16), "r"(0),
                         "r"(pipe::smem_u32(&full[s]))
                         : "memory");
    };
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (t == 0) {
        pipe::tma_prefetch_desc(&map_b);
        for (int s = 0; s < STAGES; ++s) pipe::mbar_init(&full[s], 1);
        pipe::mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (t == 0)
        for (unsigned k = 0; k < mine && k < (unsigned)STAGES; ++k) issue(k);

    const int ic = t & 15, jq = t >> 4;                           // i-chunk (4 rows of a) and j-chunk (4 columns of a) of the lane
    const int half = jq >> 3, chunk = (jq & 7) ^ (ic & 7);        // box and swizzled 16-byte chunk inside the 128-byte row
    for (unsigned k = 0; k < mine; ++k) {
        const int s = (int)(k % STAGES);
        pipe::mbar_wait(&full[s], (k / STAGES) & 1);
        const float4 *box = reinterpret_cast<const float4 *>(tma_smem + (size_t)s * 16384 + half * 8192);
        const float4 v0 = box[(0 * 16 + ic) * 8 + chunk];         // b[j0 + 4 jq .. +3, i0 + 4 ic + 0]
        const float4 v1 = box[(1 * 16 + ic) * 8 + chunk];
        const float4 v2 = box[(2 * 16 + ic) * 8 + chunk];
        const float4 v3 = box[(3 * 16 + ic) * 8 + chunk];
        __syncthreads();                                          // every lane has read stage s: it may be refilled
        if (t == 0 && k + STAGES < mine) issue(k + STAGES);
        int ti, tj;
        tile_of(k, ti, tj);
        float *dst = a + ((int64_t)ti * 64 + 4 * ic) + ((int64_t)tj * 64 + 4 * jq) * lda;
        stg_stream_f4(dst, make_float4(v0.x, v1.x, v2.x, v3.x));
        stg_stream_f4(dst + lda, make_float4(v0.y, v1.y, v2.y, v3.y));
        stg_stream_f4(dst + 2 * lda, make_float4(v0.z, v1.z, v2.z, v3.z));
        stg_stream_f4(dst + 3 * lda, make_float4(v0.w, v1.w, v2.w, v3.w));
    }
}

static int transpose_tma_launch(float *a, const float *b, int64_t m, int64_t n, int64_t lda, int64_t ldb, cudaStream_t st) {
    // b is n x m (ld ldb): dimensions (j, i_hi, i_lo) = (n, m / 4, 4), strides (4 ldb, ldb) elements
    pipe::EncodeTiledFn enc = nullptr;
    B200_TRY(pipe::get_encode_fn(&enc));
    CUtensorMap map;
    cuuint64_t gdim[3] = {(cuuint64_t)n, (cuuint64_t)(m / 4), 4};
    cuuint64_t gstride[2] = {(cuuint64_t)ldb * 16, (cuuint64_t)ldb * 4};
    cuuint32_t box[3] = {32, 16, 4};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(b), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B200_ERR_CUDA, "transpose: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    const int tiles_i = (int)(m / 64), tiles_j = (int)(n / 64);
    const int stages = tune_get("transpose.tma_stages", 3);
    const int cps = tune_get("transpose.tma_ctas_per_sm", 4);
    int64_t grid = (int64_t)sm_count() * cps;
    if (grid > (int64_t)tiles_i * tiles_j) grid = (int64_t)tiles_i * tiles_j;
    const size_t smem = (size_t)stages * 16384 + 1024;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tune_get("transpose.pdl", 1) != 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define B200_TT_CASE(S)                                                                                                         \
    case S:                                                                                                                     \
        B200_CUDA(cudaFuncSetAttribute(transpose64_tma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        B200_CUDA(cudaLaunchKernelEx(&cfg, transpose64_tma_kernel<S>, map, a, tiles_i, tiles_j, lda));                           \
        break;
    switch (stages) {
        B200_TT_CASE(2)
        B200_TT_CASE(3)
        B200_TT_CASE(4)
        B200_TT_CASE(6)
        default: return set_error(B200_ERR_INVALID_VALUE, "transpose.tma_stages must be 2, 3, 4 or 6");
    }
#undef B200_TT_CASE
    B200_CHECK_LAUNCH("transpose64_tma_kernel");
    return B200_OK;
}

// The same tile scheme for 4- AND 8-byte words and for ANY m, n (T = uint32_t / uint64_t, V = 16 / sizeof(T) words per
// 128-bit vector): interior tiles run the vector path above; edge tiles take a predicated path in which a vector is used
// when all V of its words are in range and single words otherwise.  Needs 16-byte aligned bases and lda, ldb multiples of V.
template <typename T>
struct alignas(16) Vec16 {
    T e[16 / sizeof(T)];
};
template <typename T>
__device__ __forceinline__ Vec16<T> ldg_stream_v(const T *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return *reinterpret_cast<Vec16<T> *>(&r);
}
template <typename T>
__device__ __forceinline__ void stg_stream_v(T *p, const Vec16<T> &v) {
    const uint4 r = *reinterpret_cast<const uint4 *>(&v);
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}
template <typename T, int ORDER>
__global__ void __launch_bounds__(256) transpose64_any_kernel(T *__restrict__ a, const T *__restrict__ b, int64_t m, int64_t n,
                                                              int tiles_i, int tiles_j, int64_t lda, int64_t ldb) {
    constexpr int V = 16 / sizeof(T);        // words per vector
    constexpr int CH = 64 / V;               // 16-byte chunks per tile row
    constexpr int PASSES = CH * CH / 256;    // (jq, iq) blocks per thread in the load phase
    constexpr int ROWS_PER_PASS = 256 / CH;  // tile rows covered per store pass
    __shared__ Vec16<T> tile[64 * CH];       // [j][chunk of i], chunk index XOR (j / V)
    const int t = threadIdx.x;
    int ti, tj;
    if (ORDER == 0) {
        tj = blockIdx.x % tiles_j;
        ti = blockIdx.x / tiles_j;
    } else {
        ti = blockIdx.x % tiles_i;
        tj = blockIdx.x / tiles_i;
    }
    const int64_t i0 = (int64_t)ti * 64, j0 = (int64_t)tj * 64;
    const bool interior = i0 + 64 <= m && j0 + 64 <= n;

    const int jq = t % CH;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
        const int iq = t / CH + (256 / CH) * p;
        const int64_t j = j0 + V * jq, i = i0 + V * iq;
        Vec16<T> blk[V];                     // blk[r].e[c] = b[j + c, i + r]
        if (interior) {
#pragma unroll
            for (int r = 0; r < V; ++r) blk[r] = ldg_stream_v(b + j + (i + r) * ldb);
        } else {
#pragma unroll
            for (int r = 0; r < V; ++r) {
                if (i + r < m && j + V <= n) {
                    blk[r] = ldg_stream_v(b + j + (i + r) * ldb);
                } else {
#pragma unroll
                    for (int c = 0; c < V; ++c) blk[r].e[c] = (i + r < m && j + c < n) ? b[j + c + (i + r) * ldb] : T(0);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < V; ++c) {
            Vec16<T> w;
#pragma unroll
            for (int r = 0; r < V; ++r) w.e[r] = blk[r].e[c];
            tile[(V * jq + c) * CH + (iq ^ jq)] = w;   // row V jq + c, chunk iq, swizzled by row / V == jq
        }
    }
    __syncthreads();
    const int ic = t % CH, jr = t / CH;
#pragma unroll
    for (int sidx = 0; sidx < 64 / ROWS_PER_PASS; ++sidx) {
        const int jl = jr + ROWS_PER_PASS * sidx;
        const Vec16<T> w = tile[jl * CH + (ic ^ ((jl / V) & (CH - 1)))];
        const int64_t i = i0 + V * ic, j = j0 + jl;
        if (interior || (j < n && i + V <= m)) {
            stg_stream_v(a + i + j * lda, w);
        } else if (j < n) {
#pragma unroll
            for (int r = 0; r < V; ++r)
                if (i + r < m) a[i + r + j * lda] = w.e[r];
        }
    }
}

// Generic tile transpose with bounds checks; T is a 1-, 2-, 4-, 8- or 16-byte word type.
struct alignas(16) Word16 {
    uint64_t lo, hi;
};
template <typename T>
__global__ void __launch_bounds__(256) transpose32_generic_kernel(T *__restrict__ a, const T *__restrict__ b,
                                                                  int64_t m, int64_t n, int64_t lda, int64_t ldb,
                                                                  int64_t tiles_j) {
    __shared__ T tile[32][33];
    const int64_t tj = blockIdx.x % tiles_j, ti = blockIdx.x / tiles_j;
    const int64_t i0 = ti * 32, j0 = tj * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        int64_t j = j0 + tx, i = i0 + ty + r;
        if (j < n && i < m) tile[ty + r][tx] = b[j + i * ldb];  // tile[i_local][j_local]
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        int64_t i = i0 + tx, j = j0 + ty + r;
        if (i < m && j < n) a[i + j * lda] = tile[tx][ty + r];
    }
}

int transpose_launch(void *a, const void *b, int64_t m, int64_t n, int64_t lda, int64_t ldb, int elsize,
                     cudaStream_t st) {
    if (m == 0 || n == 0) return B200_OK;
    const bool fast = elsize == 4 && (m % 64 == 0) && (n % 64 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) &&
                      (((uintptr_t)a | (uintptr_t)b) & 15) == 0 && (m / 64) * (n / 64) < (int64_t)INT32_MAX &&
                      tune_get("transpose.force_generic", 0) == 0;
    if (fast && tune_get("transpose.impl", 0) == 1 && n < (1LL << 31) && m < (1LL << 31))
        return transpose_tma_launch((float *)a, (const float *)b, m, n, lda, ldb, st);
    if (fast) {
        int tiles_i = (int)(m / 64), tiles_j = (int)(n / 64);
        unsigned grid = (unsigned)((int64_t)tiles_i * tiles_j);
        // `transpose.ctas_per_sm` (0 = as many as fit: 8) caps the resident CTAs per SM with unused dynamic shared memory: a
        // CTA lives for (bytes in flight per SM) / (bandwidth per SM), and a launch pays about one CTA lifetime of ramp + drain
        const int cps = tune_get("transpose.ctas_per_sm", 0);
        size_t dyn = 0;
        if (cps > 0 && cps < 8) dyn = (size_t)(227 * 1024) / (size_t)cps - 16 * 1024 - 1024;
        auto kern = tune_get("transpose.order", 0) == 0 ? transpose64_f32_kernel<0> : transpose64_f32_kernel<1>;
        if (dyn > 32 * 1024) B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = dyn;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = tune_get("transpose.pdl", 1) != 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // leading CTAs that pull their tile into L2 before the dependency wait: only worth it for SHORT launches (see the kernel)
        int pf = tune_get("transpose.l2_prefetch", -1);
        if (pf < 0) pf = grid <= 16384u ? 16 : 0;
        B200_CUDA(cudaLaunchKernelEx(&cfg, kern, (float *)a, (const float *)b, tiles_i, tiles_j, lda, ldb, pf * sm_count()));
        B200_CHECK_LAUNCH("transpose64_f32_kernel");
        return B200_OK;
    }
    // any shape, 4-byte words: vector tile kernel when the rows can be addressed with 128-bit accesses.  (8-byte words stay
    // on the scalar tile kernel below: 32 lanes x 8 bytes already make full 256-byte requests -- measured 6.6 TB/s against
    // 6.2 for the 128-bit variant, which is therefore not instantiated.)
    if (elsize == 4 && (lda % 4 == 0) && (ldb % 4 == 0) && (((uintptr_t)a | (uintptr_t)b) & 15) == 0 &&
        cdiv(m, 64) * cdiv(n, 64) < (int64_t)INT32_MAX && tune_get("transpose.force_generic", 0) == 0) {
        const int tiles_i = (int)cdiv(m, 64), tiles_j = (int)cdiv(n, 64);
        const unsigned grid = (unsigned)((int64_t)tiles_i * tiles_j);
        if (tune_get("transpose.order", 0) == 0)
            transpose64_any_kernel<uint32_t, 0><<<grid, 256, 0, st>>>((uint32_t *)a, (const uint32_t *)b, m, n, tiles_i, tiles_j, lda, ldb);
        else
            transpose64_any_kernel<uint32_t, 1><<<grid, 256, 0, st>>>((uint32_t *)a, (const uint32_t *)b, m, n, tiles_i, tiles_j, lda, ldb);
        B200_CHECK_LAUNCH("transpose64_any_kernel");
        return B200_OK;
    }
    int64_t tiles_i = cdiv(m, 32), tiles_j = cdiv(n, 32);
    if (tiles_i * tiles_j >= (int64_t)INT32_MAX)
        return set_error(B200_ERR_UNSUPPORTED, "transpose: %lld x %lld too large for the generic path", (long long)m,
                         (long long)n);
    unsigned grid = (unsigned)(tiles_i * tiles_j);
    if (elsize == 16 && ((((uintptr_t)a | (uintptr_t)b) & 15) != 0))
        return set_error(B200_ERR_UNSUPPORTED, "transpose: 16-byte elements need 16-byte aligned arrays");
#define B200_T32_CASE(BYTES, T)                                                                                       \
    case BYTES:                                                                                                       \
        transpose32_generic_kernel<T><<<grid, 256, 0, st>>>((T *)a, (const T *)b, m, n, lda, ldb, tiles_j);           \
        break;
    switch (elsize) {   // a pure permutation of elsize-byte words: any isbits element type of these sizes
        B200_T32_CASE(1, uint8_t)
        B200_T32_CASE(2, uint16_t)
        B200_T32_CASE(4, uint32_t)
        B200_T32_CASE(8, uint64_t)
        B200_T32_CASE(16, Word16)
        default: return set_error(B200_ERR_UNSUPPORTED, "transpose: element size %d not in {1,2,4,8,16}", elsize);
    }
#undef B200_T32_CASE
    B200_CHECK_LAUNCH("transpose32_generic_kernel");
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_transpose(void *a, const void *b, int64_t m, int64_t n, int64_t lda, int64_t ldb,
                                      int32_t elsize) {
    if (m < 0 || n < 0 || lda < m || ldb < n)
        return set_error(B200_ERR_INVALID_VALUE, "b200_transpose: bad shape m=%lld n=%lld lda=%lld ldb=%lld",
                         (long long)m, (long long)n, (long long)lda, (long long)ldb);
    if (m == 0 || n == 0) return B200_OK;
    if (!a || !b) return set_error(B200_ERR_INVALID_VALUE, "b200_transpose: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return transpose_launch(a, b, m, n, lda, ldb, elsize, st);
}
