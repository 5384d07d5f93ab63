// matmul_f32.cu -- b200_matmul_f32: C(MxN) = A(MxK) * B(KxN), column-major FP32 on the SIMT FMA pipe.
//
// Reference: coalesced_matmul_kernel! (examples/performant_matmul.jl:15-77): 16x16 work-groups, two
// (17x16) shared tiles, two barriers per 16 FMAs per item.  Its arithmetic per output element is
//     outval = -0.0;  for each 16-wide k tile: out = 0; for k in tile (ascending): out += a*b;  outval += out
// This kernel keeps exactly that arithmetic (mode B200_MATMUL_TILE16: 16-wide partials, k ascending, fused
// multiply-add, zero-padded ragged tiles) -- so it is bit-identical to the oracle's restatement -- but
// re-tiles the work for the B200 SM: 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile per thread
// (64 independent FMA chains), operands staged in double-buffered shared memory with the next tile's global
// loads in flight during the current tile's math, 128-bit shared loads (A: 8 distinct 16-byte chunks per
// warp, B: 4), one barrier per k tile.  Bound: FP32 FMA pipe, 2*M*N*K flop.
// Mode B200_MATMUL_RUNNING keeps one running accumulator per output (k ascending, fused), saving the
// 64 extra accumulators and the add per tile.
#include "../common.cuh"
#include "../tune.h"

namespace b200 {

constexpr int BM = 128, BN = 128, BK = 16;

template <bool TILE16, bool ALIGNED, int V>
__global__ void __launch_bounds__(256, 1)
    sgemm128_kernel(float *__restrict__ C, const float *__restrict__ A, const float *__restrict__ B, int M, int K,
                    int N, int64_t lda, int64_t ldb, int64_t ldc, int tiles_m, int tiles_n, int group_m) {
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    // grouped rasterisation: consecutive CTAs share A row-panels / B column-panels in L2
    int pid = blockIdx.x;
    int width = group_m * tiles_n;
    int group = pid / width;
    int first_m = group * group_m;
    int gsz = min(tiles_m - first_m, group_m);
    int tm = first_m + (pid % width) % gsz;
    int tn = (pid % width) / gsz;
    const int m0 = tm * BM, n0 = tn * BN;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wm = warp & 1, wn = warp >> 1;        // 2 x 4 warps -> warp tile 64 x 32
    // Lane -> fragment mapping chosen from tools/micro/lds_patterns.cu (profiles/r02_lds_wavefront_patterns.csv): an
    // LDS.128 costs 2 wavefronts when its address ignores lane bit 0 or lane bit 1, 4 otherwise.  A rows use lane
    // bits 1-3, B columns use lane bits 0 and 4, so all four fragment loads of a k step take the 2-wavefront path.
    const int lx = (V & 2) ? ((lane >> 1) & 7) : (lane & 7);
    const int ly = (V & 2) ? ((lane & 1) | ((lane >> 4) << 1)) : (lane >> 3);   // 8 x 4 lanes
    const int mrow = wm * 64 + 4 * lx;              // + {0, 32}
    const int ncol = wn * 32 + 4 * ly;              // + {0, 16}

    // Accumulators are float2 pairs along m so that the inner product issues packed FFMA2 (a pair from one LDS.128,
    // b as the scalar-broadcast operand): tools/micro/fma_peak.cu measures 65 TFLOP/s for FFMA2 against 46 for
    // scalar 3-register FFMA on this part.  Each lane of an FFMA2 is an IEEE fused multiply-add -> same bits.
    float2 acc[4][8];
    float2 tot[TILE16 ? 4 : 1][TILE16 ? 8 : 1];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.0f, 0.0f);
    if (TILE16) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) tot[TILE16 ? i : 0][TILE16 ? j : 0] = make_float2(-0.0f, -0.0f);
    }

    // global -> register staging of one k tile
    float4 ra[2], rb[2];
    const int a_k = t >> 5, a_m4 = t & 31;          // A: idx = t + 256 r -> k = a_k + 8 r, m = 4 a_m4
    const int b_n = t & 127, b_half = t >> 7;       // B: n = b_n, k = 8 b_half + 4 r
    auto load_global = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = k0 + a_k + 8 * r;
            const int m = m0 + 4 * a_m4;
            if (ALIGNED) {
                ra[r] = *reinterpret_cast<const float4 *>(A + m + (int64_t)k * lda);
            } else {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (m + e < M && k < K) ? A[(m + e) + (int64_t)k * lda] : 0.0f;
                ra[r] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = k0 + 8 * b_half + 4 * r;
            const int n = n0 + b_n;
            if (ALIGNED) {
                rb[r] = *reinterpret_cast<const float4 *>(B + k + (int64_t)n * ldb);
            } else {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (k + e < K && n < N) ? B[(k + e) + (int64_t)n * ldb] : 0.0f;
                rb[r] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    };
    auto store_shared = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) *reinterpret_cast<float4 *>(&As[buf][a_k + 8 * r][4 * a_m4]) = ra[r];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int k = 8 * b_half + 4 * r;
            Bs[buf][k + 0][b_n] = rb[r].x;
            Bs[buf][k + 1][b_n] = rb[r].y;
            Bs[buf][k + 2][b_n] = rb[r].z;
            Bs[buf][k + 3][b_n] = rb[r].w;
        }
    };

    const int ntiles = (K + BK - 1) / BK;
    load_global(0);
    store_shared(0);
    __syncthreads();

    for (int kt = 0; kt < ntiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ntiles) load_global((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float b[8];
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][mrow]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][mrow + 32]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][ncol]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][ncol + 16]);
            const float2 a[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y),
                                 make_float2(a1.z, a1.w)};
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 bb = make_float2(b[j], b[j]);
                    const float2 c = (TILE16 && kk == 0) ? make_float2(0.0f, 0.0f) : acc[i][j];  // out = zero(T)
                    if (V & 1)
                        acc[i][j] = __ffma2_rn(a[i], bb, c);                                      // out += a*b
                    else
                        acc[i][j] = make_float2(fmaf(a[i].x, b[j], c.x), fmaf(a[i].y, b[j], c.y));
                }
        }
        if (TILE16) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                {
                    float2 &tt = tot[TILE16 ? i : 0][TILE16 ? j : 0];                // outval += out
                    tt = (V & 1) ? __fadd2_rn(tt, acc[i][j]) : make_float2(tt.x + acc[i][j].x, tt.y + acc[i][j].y);
                }
        }
        if (kt + 1 < ntiles) store_shared(buf ^ 1);
        __syncthreads();
    }

    // epilogue: rows {mrow..mrow+3, mrow+32..+35}, cols {ncol..+3, ncol+16..+19}
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + ncol + (j & 3) + (j >> 2) * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + mrow + 32 * h;
            const float2 lo = TILE16 ? tot[TILE16 ? 2 * h : 0][TILE16 ? j : 0] : acc[2 * h][j];
            const float2 hi = TILE16 ? tot[TILE16 ? 2 * h + 1 : 0][TILE16 ? j : 0] : acc[2 * h + 1][j];
            const float v[4] = {lo.x, lo.y, hi.x, hi.y};
            if (ALIGNED) {
                *reinterpret_cast<float4 *>(C + m + (int64_t)n * ldc) = make_float4(v[0], v[1], v[2], v[3]);
            } else if (n < N) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (m + e < M) C[(m + e) + (int64_t)n * ldc] = v[e];
            }
        }
    }
}

int sgemm_tma_applicable(const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb);
int sgemm_tma_launch(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb,
                     int64_t ldc, int mode, cudaStream_t st);

int sgemm_launch(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb,
                 int64_t ldc, int mode, cudaStream_t st) {
    if (M == 0 || N == 0) return B200_OK;
    if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return set_error(B200_ERR_UNSUPPORTED, "matmul: dimension exceeds 2^31-1");
    if (mode != B200_MATMUL_TILE16 && mode != B200_MATMUL_RUNNING && mode != B200_MATMUL_UNFUSED)
        return set_error(B200_ERR_INVALID_VALUE, "matmul: unknown mode %d", mode);
    if (mode == B200_MATMUL_UNFUSED && !(tune_get("sgemm.engine", 1) == 1 && sgemm_tma_applicable(A, B, M, K, N, lda, ldb)))
        return set_error(B200_ERR_UNSUPPORTED, "matmul: the unfused mode needs TMA-addressable operands (16-byte aligned, ld %% 4 == 0)");
    // engine 1 (default): TMA-fed warp-specialised kernel (matmul_f32_tma.cu); engine 0 or operands TMA cannot address
    // (leading dimension / base not 16-byte aligned, K == 0): the register-staged kernel below.  Same bits either way.
    if (tune_get("sgemm.engine", 1) == 1 && sgemm_tma_applicable(A, B, M, K, N, lda, ldb))
        return sgemm_tma_launch(C, A, B, M, K, N, lda, ldb, ldc, mode, st);
    int tiles_m = (int)cdiv(M, BM), tiles_n = (int)cdiv(N, BN);
    if ((int64_t)tiles_m * tiles_n > INT32_MAX) return set_error(B200_ERR_UNSUPPORTED, "matmul: too many tiles");
    unsigned grid = (unsigned)(tiles_m * tiles_n);
    int group_m = tune_get("sgemm.group_m", 8);
    bool aligned = (M % BM == 0) && (N % BN == 0) && (K % BK == 0) && K > 0 && (lda % 4 == 0) && (ldb % 4 == 0) &&
                   (ldc % 4 == 0) && ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) == 0);
    const int variant = tune_get("sgemm.variant", 2) & 3;
#define B200_SGEMM_LAUNCH(T16, AL, VV)                                                                              \
    sgemm128_kernel<T16, AL, VV><<<grid, 256, 0, st>>>(C, A, B, (int)M, (int)K, (int)N, lda, ldb, ldc, tiles_m,      \
                                                       tiles_n, group_m)
#define B200_SGEMM_VARIANT(T16, AL)                                                                                 \
    switch (variant) {                                                                                              \
        case 0: B200_SGEMM_LAUNCH(T16, AL, 0); break;                                                               \
        case 1: B200_SGEMM_LAUNCH(T16, AL, 1); break;                                                               \
        case 2: B200_SGEMM_LAUNCH(T16, AL, 2); break;                                                               \
        default: B200_SGEMM_LAUNCH(T16, AL, 3); break;                                                              \
    }
    if (mode == B200_MATMUL_TILE16) {
        if (aligned) { B200_SGEMM_VARIANT(true, true) } else { B200_SGEMM_VARIANT(true, false) }
    } else {
        if (aligned) { B200_SGEMM_VARIANT(false, true) } else { B200_SGEMM_VARIANT(false, false) }
    }
#undef B200_SGEMM_VARIANT
#undef B200_SGEMM_LAUNCH
    B200_CHECK_LAUNCH("sgemm128_kernel");
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_matmul_f32(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N,
                                       int64_t lda, int64_t ldb, int64_t ldc, int32_t mode) {
    if (M < 0 || N < 0 || K < 0 || lda < M || ldb < K || ldc < M)
        return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32: bad shape");
    if (M == 0 || N == 0) return B200_OK;
    if (!C || (K > 0 && (!A || !B))) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return sgemm_launch(C, A, B, M, K, N, lda, ldb, ldc, mode, st);
}
