// matmul_f32_tma.cu -- FP32 SIMT GEMM fed by TMA: the default engine behind b200_matmul_f32.
//
// Reference: coalesced_matmul_kernel! (examples/performant_matmul.jl:15-77).  Arithmetic per output element (kept
// bit for bit in mode B200_MATMUL_TILE16, see matmul_f32.cu and the parity tests):
//     outval = -0.0;  per 16-wide k tile: out = 0; for k ascending: out = fma(a, b, out);  outval += out
// with ragged edge tiles zero-padded (:40-49) -- which is what TMA's out-of-bounds zero fill produces for free.
//
// Why a second engine (measured in profiles/r02_*): the register-staged kernel of matmul_f32.cu spends issue slots and
// ~16 registers on LDG -> STS staging, transposes B through 4-byte shared stores, and leaves the position of the
// global loads to ptxas (which sank them next to the stores when the FMAs became FFMA2, exposing DRAM latency).  Here
//   * a dedicated producer warp group (warps 8-11, one active lane) issues two cp.async.bulk.tensor.2d per 16-k stage (A box
//     128 m x 16 k, B box 16 k x 16 nj n) into a STAGES-deep ring, completion on `full` mbarriers, slot release on `empty`
//     mbarriers.  Registers are handed to warps in groups of four, so the group gives its registers back with `setmaxnreg.dec`
//     and the consumers take them with `setmaxnreg.inc` (TILE16 needs 206).  Against thread 0 of a consumer warp producing
//     (PW = 0, kept for comparison): 58.6 -> 61.2 TFLOP/s in the running-sum mode;
//   * the consumers software-pipeline the whole k stream: operands of the next k step are in flight while the current
//     step's FFMA2s issue, across tile boundaries too (the next `full` barrier is polled before the last step of a tile).
//     tools/micro/sgemm_inner.cu isolates the inner loop: 64.1 TFLOP/s with both operand streams from shared memory,
//     68.4 without loads -- the pipeline around it now costs 4.5 %;
//   * 8 consumer warps: 2 x 4 warps, warp tile 64 x 4 nj, thread tile 8 x nj held as float2 pairs along m, inner
//     product issued as packed FFMA2 (a pair = half of one LDS.128, b = scalar-broadcast operand).  FFMA2 peaks at
//     65 TFLOP/s on this part against 46 for 3-register scalar FFMA (tools/micro/fma_peak.cu);
//   * B stays k-contiguous in shared memory exactly as it is in HBM (no transpose): a thread reads two consecutive k of
//     one column with one LDS.64 and walks its 8 columns;
//   * lane -> (row, column) mapping takes the cheap shared-memory path measured by tools/micro/lds_patterns.cu: a
//     wavefront serves a whole request only if its address ignores lane bit 0 or lane bit 1, so A rows use lane bits
//     1-3 and B columns lane bits 0 and 4: 8 wavefronts per k step per warp instead of 13.
// Bound: FP32 FMA pipe, 2*M*N*K flop.
#include "../common.cuh"
#include "../tma_pipe.cuh"
#include "../tune.h"

namespace b200 {
namespace sg {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 8;                    // measured with the producer warp group: 6 stages 57.2, 8 stages 57.3, 12 stages 56.5 TFLOP/s
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
// MI = float4 row groups per thread: 2 -> 8 warps (2 x 4), thread tile 8 x nj; 1 -> 16 warps (4 x 4), thread tile 4 x nj
constexpr int consumer_warps(int MI) { return 16 / MI; }
constexpr int LAG = 2;                       // a slot is refilled LAG k tiles after it was consumed
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;

// NJ = columns per thread (8, 7 or 6): the CTA tile is 128 x 16 NJ.  The narrower tiles exist only to fill whole waves of
// the 148-SM grid (pick_nj below); the arithmetic per output element -- and therefore every bit of C -- does not change.
// PW = 1: a dedicated producer warp group (warps CONSUMER_WARPS .. +3; registers are allocated to warps in groups of four, so a
// single extra warp would cost as much) issues the TMA loads, so no consumer warp ever diverges into producer code or blocks
// on an `empty` barrier.  The kernel is compiled for 384 threads (168 registers each); the producer group gives its
// registers back (`setmaxnreg.dec 40`) and the consumer groups take them (`setmaxnreg.inc 232`): 8*32*232 + 4*32*40 = 64512.
// PW = 0: thread 0 of consumer warp 0 produces with a lag of LAG tiles.
// UNFUSED mode: round(a*b) and round(tmp + product) as TWO packed instructions.  ptxas 12.9 contracts `mul.rn.f32x2` followed by
// `add.rn.f32x2` into one FFMA2 (scalar `mul.rn.f32` / `add.rn.f32` are left alone, as the PTX manual promises; the packed forms
// are not -- the first two builds of this mode ran at FFMA2 speed and failed bit parity, /tmp experiment recorded in DESIGN.md).
// Two FFMA2 with RUN-TIME constants cannot be folded or contracted and round identically:
//     fma(a, b, -0.0) == round(a*b)   (adding -0.0 changes neither value nor sign of zero),   fma(p, 1.0, c) == round(p + c).
// `one` and `negzero` arrive as kernel arguments so ptxas cannot see their values.
__device__ __forceinline__ float2 mul_then_add(float2 a, float2 b, float2 c, float one, float negzero) {
    const float2 p = __ffma2_rn(a, b, make_float2(negzero, negzero));
    return __ffma2_rn(p, make_float2(one, one), c);
}

// MODE: B200_MATMUL_TILE16 (reference coalesced_matmul_kernel! order), B200_MATMUL_RUNNING (one fused running sum) or
// B200_MATMUL_UNFUSED (matmul_kernel!, examples/matmul.jl:5-15: tmp = zero(T); tmp += a*b with the product and the sum rounded
// separately, k ascending -- packed FMUL2 + FADD2, two pipe instructions per pair of MACs).
template <int MODE, int NJ, int MI, int PW>
__global__ void __launch_bounds__((consumer_warps(MI) + 4 * PW) * 32, 1)
    sgemm_tma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     float *__restrict__ C, int M, int N, int64_t ldc, int ntiles, int tiles_m, int tiles_n, int group_m,
                     int vec_ok, float one, float negzero) {
    using namespace pipe;
    constexpr bool TILE16 = MODE == B200_MATMUL_TILE16, UNFUSED = MODE == B200_MATMUL_UNFUSED;
    constexpr int CONSUMER_WARPS = consumer_warps(MI);
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = (uint64_t *)(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;

    // grouped rasterisation: consecutive CTAs share A row-panels / B column-panels in L2
    const int pid = blockIdx.x;
    const int width = group_m * tiles_n;
    const int group = pid / width;
    const int first_m = group * group_m;
    const int gsz = min(tiles_m - first_m, group_m);
    const int tm = first_m + (pid % width) % gsz;
    const int tn = (pid % width) / gsz;
    constexpr int BN_ = 16 * NJ;
    const int m0 = tm * BM, n0 = tn * BN_;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // 32-bit shared-window addresses, computed once (every smem_u32() inside the loop would re-read SR_CgaCtaId)
    uint32_t smem_a = smem_u32(smem), full_a = smem_u32(full), empty_a = smem_u32(empty);
    asm volatile("" : "+r"(smem_a), "+r"(full_a), "+r"(empty_a));
    auto wait_a = [](uint32_t bar, uint32_t parity) {
        asm volatile(
            "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
    };
    auto produce = [&](int kt) {   // thread 0 only
        const int s = kt % STAGES;
        if (kt >= STAGES) wait_a(empty_a + 8 * s, ((kt / STAGES) - 1) & 1);
        const uint32_t bar = full_a + 8 * s, stage = smem_a + s * STAGE_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(A_BYTES + BN_ * BK * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(stage),
                     "l"(&tmap_a), "r"(m0), "r"(kt * BK), "r"(bar)
                     : "memory");                                               // As[k][m]: 16 rows of 512 B
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(stage + A_BYTES),
                     "l"(&tmap_b), "r"(kt * BK), "r"(n0), "r"(bar)
                     : "memory");                                               // Bs[n][k]: 16 NJ rows of 64 B
    };
    if (PW) {
        if (warp >= CONSUMER_WARPS) {          // producer warp group: one lane runs ahead as far as the ring allows
            asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
            if (t == CONSUMER_WARPS * 32)
                for (int kt = 0; kt < ntiles; ++kt) produce(kt);
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    } else if (t == 0) {
        for (int kt = 0; kt < STAGES - LAG && kt < ntiles; ++kt) produce(kt);
    }

    // ---------------------------------------------------------------------- consumers
    constexpr int WARPS_M = 4 / MI;                                 // (4 / MI) x 4 warps -> warp tile 32 MI x 4 NJ
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int lx = (lane >> 1) & 7;                                 // lane bits 1-3 -> rows 4 lx .. 4 lx + 3 (+32)
    const int ly = (lane & 1) | ((lane >> 4) << 1);                 // lane bits 0,4 -> columns ly + 4 j
    const int mrow = wm * (32 * MI) + 4 * lx;
    const int ncol = wn * (4 * NJ) + ly;

    float2 acc[2 * MI][NJ];
    float2 tot[TILE16 ? 2 * MI : 1][TILE16 ? NJ : 1];
#pragma unroll
    for (int i = 0; i < 2 * MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = make_float2(0.0f, 0.0f);
    if (TILE16) {
#pragma unroll
        for (int i = 0; i < 2 * MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) tot[TILE16 ? i : 0][TILE16 ? j : 0] = make_float2(-0.0f, -0.0f);
    }

    // Software pipeline over the whole k stream: while the FFMA2s of k step (kt, k2) issue, the operands of the NEXT step are
    // already in flight -- including across the tile boundary, where the next step is (kt+1, 0) and its `full` barrier is
    // polled before the last step of tile kt is computed.  tools/micro/sgemm_inner.cu: the inner loop alone sustains
    // 64 TFLOP/s; what the kernel lost against it was the exposed barrier poll + shared-memory latency at every tile start.
    float2 b[2][NJ];
    float4 av[2][2][MI];
    auto load_frags = [&](int buf, int stage, int k2) {
        const float *As = reinterpret_cast<const float *>(smem + stage * STAGE_BYTES) + mrow;
        const float *Bs = reinterpret_cast<const float *>(smem + stage * STAGE_BYTES + A_BYTES) + ncol * BK;
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[buf][j] = *reinterpret_cast<const float2 *>(Bs + (4 * j) * BK + k2);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int g = 0; g < MI; ++g) av[buf][h][g] = *reinterpret_cast<const float4 *>(As + (k2 + h) * BM + 32 * g);
    };
    int s = 0;
    uint32_t phase = 0;
    wait_a(full_a, 0);
    load_frags(0, 0, 0);
    for (int kt = 0; kt < ntiles; ++kt) {
        if (!PW && t == 0 && kt + STAGES - LAG < ntiles) produce(kt + STAGES - LAG);
        const int s_next = s + 1 == STAGES ? 0 : s + 1;
        const uint32_t phase_next = s + 1 == STAGES ? phase ^ 1 : phase;
#pragma unroll
        for (int k2 = 0; k2 < BK; k2 += 2) {
            const int cur = (k2 >> 1) & 1;
            if (k2 + 2 < BK) {
                load_frags(cur ^ 1, s, k2 + 2);
            } else if (kt + 1 < ntiles) {
                wait_a(full_a + 8 * s_next, phase_next);
                load_frags(cur ^ 1, s_next, 0);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = k2 + h;
                float2 a[2 * MI];
#pragma unroll
                for (int g = 0; g < MI; ++g) {
                    a[2 * g] = make_float2(av[cur][h][g].x, av[cur][h][g].y);
                    a[2 * g + 1] = make_float2(av[cur][h][g].z, av[cur][h][g].w);
                }
#pragma unroll
                for (int i = 0; i < 2 * MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const float bj = h ? b[cur][j].y : b[cur][j].x;
                        const float2 c = (TILE16 && kk == 0) ? make_float2(0.0f, 0.0f) : acc[i][j];  // out = zero(T)
                        if (UNFUSED)
                            acc[i][j] = mul_then_add(a[i], make_float2(bj, bj), c, one, negzero);     // tmp += (a*b), two roundings
                        else
                            acc[i][j] = __ffma2_rn(a[i], make_float2(bj, bj), c);                     // out += a*b
                    }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a + 8 * s) : "memory");   // stage s released
        s = s_next;
        phase = phase_next;
        if (TILE16) {
#pragma unroll
            for (int i = 0; i < 2 * MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    float2 &tt = tot[TILE16 ? i : 0][TILE16 ? j : 0];
                    tt = __fadd2_rn(tt, acc[i][j]);                                                  // outval += out
                }
        }
    }

    // epilogue: rows mrow + 32 h .. + 3 (h < MI) of columns ncol + 4 j
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int n = n0 + ncol + 4 * j;
#pragma unroll
        for (int h = 0; h < MI; ++h) {
            const int m = m0 + mrow + 32 * h;
            const float2 lo = TILE16 ? tot[TILE16 ? 2 * h : 0][TILE16 ? j : 0] : acc[2 * h][j];
            const float2 hi = TILE16 ? tot[TILE16 ? 2 * h + 1 : 0][TILE16 ? j : 0] : acc[2 * h + 1][j];
            if (n < N) {
                float *dst = C + m + (int64_t)n * ldc;
                if (vec_ok && m + 3 < M) {
                    *reinterpret_cast<float4 *>(dst) = make_float4(lo.x, lo.y, hi.x, hi.y);
                } else {
                    const float v[4] = {lo.x, lo.y, hi.x, hi.y};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (m + e < M) dst[e] = v[e];
                }
            }
        }
    }
}

}  // namespace sg

// Returns B200_OK after launching, B200_ERR_UNSUPPORTED (without touching the error string's meaning for the caller)
// when the operands do not satisfy TMA's alignment rules -- the caller then uses the register-staged engine.
int sgemm_tma_applicable(const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb) {
    return M > 0 && N > 0 && K > 0 && (lda % 4 == 0) && (ldb % 4 == 0) && ((((uintptr_t)A | (uintptr_t)B) & 15) == 0) &&
           M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31);
}

template <int MODE, int NJ, int MI, int PW>
static int sgemm_tma_launch_nj(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda,
                               int64_t ldb, int64_t ldc, cudaStream_t st) {
    using namespace sg;
    constexpr int BN_ = 16 * NJ;
    CUtensorMap ta, tb;
    B200_TRY(pipe::make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, M, K, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    B200_TRY(pipe::make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, B, K, N, ldb, BK, BN_, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
    const int tiles_m = (int)cdiv(M, BM), tiles_n = (int)cdiv(N, BN_);
    if ((int64_t)tiles_m * tiles_n > INT32_MAX) return set_error(B200_ERR_UNSUPPORTED, "matmul: too many tiles");
    const unsigned grid = (unsigned)(tiles_m * tiles_n);
    const int ntiles = (int)cdiv(K, BK);
    const int group_m = tune_get("sgemm.group_m", 8);
    const int vec_ok = (ldc % 4 == 0) && (((uintptr_t)C & 15) == 0);
    B200_CUDA(cudaFuncSetAttribute(sgemm_tma_kernel<MODE, NJ, MI, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    sgemm_tma_kernel<MODE, NJ, MI, PW><<<grid, (consumer_warps(MI) + 4 * PW) * 32, SMEM_BYTES, st>>>(ta, tb, C, (int)M, (int)N, ldc, ntiles, tiles_m,
                                                                         tiles_n, group_m, vec_ok, 1.0f, -0.0f);
    B200_CHECK_LAUNCH("sgemm_tma_kernel");
    return B200_OK;
}

// Columns per thread: CTAs run one per SM in rounds of `sms`, a round lasts ~NJ units, so minimise rounds * NJ (ties keep
// the widest tile).  A 1024 x 8192 row panel (the 8-GPU shard of 8192^3) is 512 tiles of 128 x 128 = 3.46 rounds -> 4,
// but 592 tiles of 128 x 112 = exactly 4 rounds of 7/8 the length.  `sgemm.nj` pins a value.
static int pick_nj(int64_t M, int64_t N, int sms) {
    const int forced = tune_get("sgemm.nj", 0);
    if (forced >= 6 && forced <= 8) return forced;
    int best = 8;
    int64_t best_cost = INT64_MAX;
    for (int nj = 8; nj >= 6; --nj) {
        const int64_t cost = cdiv(cdiv(M, sg::BM) * cdiv(N, 16 * nj), sms) * nj;
        if (cost < best_cost) {
            best_cost = cost;
            best = nj;
        }
    }
    return best;
}

int sgemm_tma_launch(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda,
                     int64_t ldb, int64_t ldc, int mode, cudaStream_t st) {
    const int nj = pick_nj(M, N, sm_count());
    const int mi = 2;   // MI = 1 (16 warps, thread tile 4 x nj) measured no faster (54.3 vs 54.5 TFLOP/s) and is not instantiated
    const int pw = tune_get("sgemm.producer_warp", 1) != 0 || mode == B200_MATMUL_UNFUSED;   // UNFUSED is built for PW = 1 only
#define B200_SGEMM_CASE(MODE_, NJ_, MI_, PW_) \
    if (mode == MODE_ && nj == NJ_ && mi == MI_ && pw == PW_) \
        return sgemm_tma_launch_nj<MODE_, NJ_, MI_, PW_>(C, A, B, M, K, N, lda, ldb, ldc, st);
    B200_SGEMM_CASE(0, 8, 2, 1) B200_SGEMM_CASE(0, 7, 2, 1) B200_SGEMM_CASE(0, 6, 2, 1)
    B200_SGEMM_CASE(1, 8, 2, 1) B200_SGEMM_CASE(1, 7, 2, 1) B200_SGEMM_CASE(1, 6, 2, 1)
    B200_SGEMM_CASE(2, 8, 2, 1) B200_SGEMM_CASE(2, 7, 2, 1) B200_SGEMM_CASE(2, 6, 2, 1)
    B200_SGEMM_CASE(0, 8, 2, 0) B200_SGEMM_CASE(0, 7, 2, 0) B200_SGEMM_CASE(0, 6, 2, 0)
    B200_SGEMM_CASE(1, 8, 2, 0) B200_SGEMM_CASE(1, 7, 2, 0) B200_SGEMM_CASE(1, 6, 2, 0)
#undef B200_SGEMM_CASE
    return set_error(B200_ERR_UNSUPPORTED, "sgemm: no kernel for nj=%d mi=%d pw=%d", nj, mi, pw);
}

}  // namespace b200
