// matmul_tc.cu -- opt-in tensor-core GEMMs: tcgen05.mma (BF16, TF32 or split-TF32 operands, FP32 accumulate in TMEM), TMA operand
// staging, mbarrier pipelines, warp-specialised persistent CTAs and CTA pairs.
//
// Reference: there is none -- KernelAbstractions' only dense contraction is the FP32 SIMT kernel of
// examples/performant_matmul.jl:15-77.  north_star asks for an opt-in BF16 tensor-core path for it; parity is
// therefore "unpinned"  and defined as: FP32-oracle GEMM on BF16-rounded inputs, rtol 1e-2.
//
// Shape of the computation: C(MxN, column-major FP32) = A * B with both operands stored K-contiguous in BF16
//   Abf: M rows x K (row-major)      Bbf: N rows x K (== column-major KxN as Julia stores B)
// so both are "K-major" for UMMA and one SWIZZLE_128B TMA box per operand per stage lands them in shared
// memory in exactly the canonical layout the UMMA shared-memory descriptor expects.
//
// TF32 variant (b200_matmul_tf32, template parameter TF32): the SAME pipeline fed with the FP32 matrices exactly as they
// lie in HBM -- no conversion pass at all.  B (K x N column-major) is K-major as before (box 32 k x 256 n, 128-byte rows);
// A (M x K column-major) is M-contiguous, i.e. "MN-major" for UMMA: four TMA boxes of 32 m x 32 k per stage land as four
// 4 KiB slabs [k][32 m].  For 32-bit MN-major operands the tensor core accepts ONE shared-memory layout only, the 128-byte
// swizzle with a 32-byte atom (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> descriptor layout type 1): 32-byte chunk c of
// k-row r is stored at chunk c ^ (r & 3); canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units with
// LBO = 4096 B (next 32 rows of M) and SBO = 512 B (next 4 k).  With the ordinary SWIZZLE_128B image the MMA silently
// produces zeros (tools/micro/tf32_probe.cu tries the variants).  The instruction descriptor sets a_major = MN and
// formats = TF32; each tcgen05.mma.kind::tf32 consumes K = 8 (two 4-k atoms, +1024 B per step).  The tensor core reads the FP32 words and
// uses their top 19 bits.  Edge tiles: TMA zero-fills out-of-range rows/columns/k, the epilogue predicates its stores.
//
// CTA (192 threads, 1 per SM, persistent over its share of the output tiles):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor.2d operand boxes per stage into a ring (completion on `full` mbarriers)
//   warp 1   : TMEM allocator + MMA issuer -- one elected lane issues 4 tcgen05.mma per stage (12 in the 3xTF32 mode),
//              tcgen05.commit releases the stage / publishes the accumulator
//   warps 2-5: epilogue       -- tcgen05.ld 32 lanes x 32 columns -> registers -> coalesced st.global
//   TMEM     : 2 accumulator buffers of 256 columns, so the epilogue of one (tile, k chunk) overlaps the MMAs of the next.
// Variants (template parameters): NCTA = 2 runs CTA pairs (cta_group::2, 256 x bn tiles, B split across the pair, 6 stages instead
// of 4); X3 = 1 / 2 is the split-operand FP32-grade mode with its correction accumulator and K chunks (see PipeCfg below; 2 = 256-wide
// tiles, one TMEM buffer, 8 epilogue warps = 320 threads); the tile width bn and the k-chunk length are run-time arguments.
#include <cuda.h>
#include <cuda_bf16.h>

#include "../common.cuh"
#include "../tma_pipe.cuh"
#include "../tune.h"

namespace b200 {

namespace tc {
constexpr int BM = 128, BN = 256, BK = 64;   // BF16: 64 k per stage, 16 per MMA (4 MMAs per stage)
constexpr int BK_TF32 = 32;                  // TF32: 32 k per stage,  8 per MMA (same bytes, 4 MMAs per stage)
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
static_assert(BM * BK_TF32 * 4 == A_BYTES && BN * BK_TF32 * 4 == B_BYTES, "both variants use 128-byte operand rows");
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// MN-major, SWIZZLE_128B_BASE32B (layout type 1) canonical layout of a 128 (M) x 32 (k) FP32 operand staged as four
// [k][32 m] slabs: LBO = 4096 B between 32-row groups of M, SBO = 512 B between groups of 4 k, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// K-major, SWIZZLE_128B canonical layout: rows 128 B apart, 8-row atoms 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D=F32, A=B=BF16, both K-major, N=BN, M=BM
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10);   // | (bn >> 3) << 17 | (M >> 4) << 24 at run time
// kind::tf32 instruction descriptor: D=F32, A=B=TF32, A MN-major (bit 15), B K-major, N=BN, M=BM
constexpr uint32_t kIdescTf32 =
    (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA-pair (cta_group::2) plumbing -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are credited to the mbarrier at `bar_cluster_addr`
// (the leader CTA's `full` barrier): the cta_group::2 form allows the barrier to live in the peer CTA.
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *tmap, int c0, int c1,
                                                 uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
        : "memory");
}
// tcgen05.commit of a CTA pair: arrives on the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (TF32)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                     "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                     : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                     "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                     : "memory");
}

// NCTA = 1: one CTA per 128 x bn tile.  NCTA = 2: a CTA pair (cluster of 2, one TPC) per 256 x bn tile -- each CTA stages
// its own 128 rows of A and HALF of the tile's B columns; the leader's tcgen05.mma.cta_group::2 (M = 256) reads both
// halves, so every SM pulls 16 + 16 KiB per k block through L2 instead of 16 + 32, and six stages fit instead of four.
// X3 = 1 ("3xTF32", b200_matmul_f32_tf32x3): every operand is split into its TF32 head (what the tensor core reads from the FP32
// word anyway) and the TF32 head of the remainder, A = A_hi + A_lo, B = B_hi + B_lo, and each k step issues A_lo*B_hi, A_hi*B_lo
// and A_hi*B_hi into the same accumulator (the lo*lo term is below FP32 resolution).  A stage then holds four operand tiles.
// The tensor core truncates when it accumulates, an error that grows linearly with k, so K is processed in chunks of
// kb_chunk k blocks: each chunk is accumulated in TMEM and the epilogue adds it to C in FP32 with round-to-nearest.
// X3 = 2 is the same arithmetic on 256-wide tiles (CTA pairs only): main accumulator in TMEM columns [0, 256), correction
// accumulator in [256, 512), i.e. ONE buffer -- the MMAs of a chunk wait for the epilogue of the previous one -- and eight epilogue
// warps (two per TMEM lane group, 128 columns each) because a thread can carry 128 chunk sums, not 256.  Why: with 128-wide tiles
// shared memory is the limit -- per k block TMA writes 48 KiB and the twelve MMAs read 12 x (4 + 2) KiB, 120 KiB in the 768 clocks the
// tensor core needs = 156 B/clk against 128 B/clk (plain TF32 at bn = 128 measures the same 0.70 of its bn = 256 rate); 256-wide tiles
// move 64 + 96 KiB per 1536 clocks = 104 B/clk.
template <int NCTA, int X3>
struct PipeCfg {
    static constexpr int SPLIT = X3 ? 1 : 0;
    static constexpr int B_STAGE = (X3 == 1 ? B_BYTES / 2 : B_BYTES) / NCTA;   // X3 = 1 tiles are at most BN / 2 wide
    static constexpr int STAGES = X3 == 2 ? 3 : X3 ? (NCTA == 2 ? 4 : 3) : (NCTA == 2 ? 6 : 4);
    static constexpr int STAGE_BYTES = (1 + SPLIT) * (A_BYTES + B_STAGE);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int EPI_WARPS = X3 == 2 ? 8 : 4;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int CORR_OFF = X3 == 2 ? BN : BN / 2;   // TMEM column offset of the correction accumulator
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool TF32, int NCTA, int X3>
__global__ void __launch_bounds__(PipeCfg<NCTA, X3>::THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ CUtensorMap tmap_alo, const __grid_constant__ CUtensorMap tmap_blo,
                        float *__restrict__ C, int64_t ldc, int M, int N, int tiles_m, int tiles_n, int num_kb,
                        int group_m, int bn, int kb_chunk) {
    static_assert(!X3 || TF32, "the split-operand mode is a TF32 mode");
    // bn = columns per output tile (multiple of 16 NCTA, <= BN): chosen by the host so that the tile count fills whole
    // rounds of the persistent grid (pick_bn); shared-memory stages and TMEM accumulators stay sized for BN.
    constexpr int KB = TF32 ? BK_TF32 : BK;   // k elements per stage
    constexpr int STAGES = PipeCfg<NCTA, X3>::STAGES, STAGE_BYTES = PipeCfg<NCTA, X3>::STAGE_BYTES;
    constexpr int LO_OFF = A_BYTES + PipeCfg<NCTA, X3>::B_STAGE;   // X3: the lo tiles follow the hi tiles inside a stage
    constexpr int SPLIT = PipeCfg<NCTA, X3>::SPLIT, CORR_OFF = PipeCfg<NCTA, X3>::CORR_OFF;
    constexpr bool WIDE = X3 == 2;                                 // one accumulator buffer, eight epilogue warps
    static_assert(!WIDE || NCTA == 2, "the 256-wide split-operand tiles are a CTA-pair layout");
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned stage bases (the dynamic window starts at the same offset in both
    // CTAs of a pair, so the aligned bases coincide as the pair MMA requires)
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *full = (uint64_t *)(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = (uint32_t *)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = tiles_m * tiles_n;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs)
    const int unit = blockIdx.x / NCTA, num_units = gridDim.x / NCTA;
    const int bnh = bn / NCTA;                                      // B rows staged by this CTA

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);    // leader's expect_tx arrive; bytes from both CTAs' TMA loads
            mbar_init(&empty[s], 1);   // one tcgen05.commit (multicast to both CTAs of a pair)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], PipeCfg<NCTA, X3>::EPI_WARPS * NCTA);   // one arrive per epilogue warp of every CTA of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int tile, int &tm, int &tn) {
        int width = group_m * tiles_n;
        int group = tile / width;
        int first_m = group * group_m;
        int gsz = min(tiles_m - first_m, group_m);
        tm = first_m + (tile % width) % gsz;
        tn = (tile % width) / gsz;
    };

    if (warp == 0) {
        // ===== TMA producer (every CTA: own A rows, own share of the B columns) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                int tm, tn;
                tile_coords(tile, tm, tn);
                const int m_base = (tm * NCTA + (int)rank) * BM, n_base = tn * bn + (int)rank * bnh;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *sa = smem + stage * STAGE_BYTES;
                    unsigned char *sb = sa + A_BYTES;
                    if (NCTA == 2) {
                        const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
                        if (rank == 0) mbar_expect_tx(&full[stage], (1 + SPLIT) * 2 * (A_BYTES + bnh * 128));
                        if (TF32) {
#pragma unroll
                            for (int g = 0; g < BM / 32; ++g)
                                tma_load_2d_pair(sa + g * 4096, &tmap_a, m_base + 32 * g, kb * KB, lead_full);
                            if (X3) {
#pragma unroll
                                for (int g = 0; g < BM / 32; ++g)
                                    tma_load_2d_pair(sa + LO_OFF + g * 4096, &tmap_alo, m_base + 32 * g, kb * KB, lead_full);
                                tma_load_2d_pair(sb + LO_OFF, &tmap_blo, kb * KB, n_base, lead_full);
                            }
                        } else {
                            tma_load_2d_pair(sa, &tmap_a, kb * KB, m_base, lead_full);
                        }
                        tma_load_2d_pair(sb, &tmap_b, kb * KB, n_base, lead_full);
                    } else {
                        mbar_expect_tx(&full[stage], (1 + SPLIT) * (A_BYTES + bn * 128));
                        if (X3) {
#pragma unroll
                            for (int g = 0; g < BM / 32; ++g)
                                tma_load_2d(sa + LO_OFF + g * 4096, &tmap_alo, m_base + 32 * g, kb * KB, &full[stage]);
                            tma_load_2d(sb + LO_OFF, &tmap_blo, kb * KB, n_base, &full[stage]);
                        }
                        if (TF32) {
#pragma unroll
                            for (int g = 0; g < BM / 32; ++g)   // A slabs [k][32 m]: coordinates (m, k) of column-major A
                                tma_load_2d(sa + g * 4096, &tmap_a, m_base + 32 * g, kb * KB, &full[stage]);
                        } else {
                            tma_load_2d(sa, &tmap_a, kb * KB, m_base, &full[stage]);
                        }
                        tma_load_2d(sb, &tmap_b, kb * KB, n_base, &full[stage]);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = (TF32 ? kIdescTf32 : kIdesc) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = unit; tile < num_tiles; tile += num_units) {
              for (int kb0 = 0; kb0 < num_kb; kb0 += kb_chunk, ++it) {   // one TMEM accumulator per (tile, k chunk)
                const int acc = WIDE ? 0 : (it & 1);
                const uint32_t acc_phase = (uint32_t)(WIDE ? it : (it >> 1)) & 1;
                // WIDE: only the MAIN accumulator is handed to the epilogue at a chunk boundary, so the correction MMAs around the boundary
                // keep the tensor core busy while the epilogue drains -- the last k block of a chunk issues main, commit, corrections; the
                // first block of the next chunk corrections, wait for the drain, main.  (First chunk of a tile: the epilogue of the previous
                // tile still reads the correction accumulator, so wait before anything.)
                if (!WIDE || kb0 == 0) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                }
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                const int kb1 = min(kb0 + kb_chunk, num_kb);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t db = make_desc_k128(sa + A_BYTES);
                    const uint64_t da = TF32 ? make_desc_mn128(sa) : make_desc_k128(sa);
                    // TF32: A's next group of 8 k = +1024 B = 64 units; K-major operands: +32 B inside the 128-byte
                    // swizzle row = 2 units per MMA (UMMA_K * 2 bytes for BF16, UMMA_K_TF32 * 4 bytes for TF32)
                    constexpr int a_step = TF32 ? 64 : 2;
                    const uint32_t first = (uint32_t)(kb - kb0);
                    if (WIDE) {
                        const uint64_t dal = make_desc_mn128(sa + LO_OFF), dbl = make_desc_k128(sa + LO_OFF + A_BYTES);
                        const uint32_t tmem_s = tmem_d + (uint32_t)CORR_OFF;
                        const bool lead = kb0 != 0 && kb == kb0, tail = kb1 != num_kb && kb == kb1 - 1;
                        auto corr = [&](int k) {   // A_lo * B_hi + A_hi * B_lo; restarted once per TILE, not per chunk (~2^-11 of the main term)
                            umma_pair<true>(tmem_s, dal + (uint64_t)(a_step * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                            umma_pair<true>(tmem_s, da + (uint64_t)(a_step * k), dbl + (uint64_t)(2 * k), idesc, 1u);
                        };
                        auto mainm = [&](int k) {
                            umma_pair<true>(tmem_d, da + (uint64_t)(a_step * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((first | (uint32_t)k) != 0));
                        };
                        if (lead) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) corr(k);
                            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                            tc_fence_after();
#pragma unroll
                            for (int k = 0; k < 4; ++k) mainm(k);
                            if (tail) umma_commit_pair(&tmem_full[acc]);
                        } else if (tail) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) mainm(k);
                            umma_commit_pair(&tmem_full[acc]);   // covers every main MMA of the chunk; the corrections below run under the drain
#pragma unroll
                            for (int k = 0; k < 4; ++k) corr(k);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                corr(k);
                                mainm(k);
                            }
                        }
                    } else
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t dak = da + (uint64_t)(a_step * k), dbk = db + (uint64_t)(2 * k);
                        if (X3) {
                            // The tensor core truncates on every accumulation, relative to the accumulator's magnitude.  The two
                            // correction terms therefore get an accumulator of their own (columns +128 of the same buffer; bn <= 128
                            // in this mode): their sum is ~2^-11 of the main term, so their truncation is invisible, and the
                            // main accumulator sees one truncation per k step instead of three.
                            const uint64_t dalk = make_desc_mn128(sa + LO_OFF) + (uint64_t)(a_step * k);
                            const uint64_t dblk = make_desc_k128(sa + LO_OFF + A_BYTES) + (uint64_t)(2 * k);
                            const uint32_t tmem_s = tmem_d + (uint32_t)CORR_OFF;
                            // WIDE: the correction accumulator is NOT restarted per k chunk -- it is ~2^-11 of the main one, so its own
                            // truncation stays ~2^-25 of the result over any K -- and the epilogue reads it once per tile.
                            if (NCTA == 2) {
                                umma_pair<true>(tmem_s, dalk, dbk, idesc, WIDE ? (uint32_t)((kb | k) != 0) : (uint32_t)((first | k) != 0));
                                umma_pair<true>(tmem_s, dak, dblk, idesc, 1u);
                                umma_pair<true>(tmem_d, dak, dbk, idesc, (first | k) != 0);
                            } else {
                                umma_tf32(tmem_s, dalk, dbk, idesc, (first | k) != 0);
                                umma_tf32(tmem_s, dak, dblk, idesc, 1u);
                                umma_tf32(tmem_d, dak, dbk, idesc, (first | k) != 0);
                            }
                        } else if (NCTA == 2) {
                            umma_pair<TF32>(tmem_d, dak, dbk, idesc, (first | k) != 0);
                        } else if (TF32) {
                            umma_tf32(tmem_d, dak, dbk, idesc, (first | k) != 0);
                        } else {
                            umma_bf16(tmem_d, dak, dbk, idesc, (first | k) != 0);
                        }
                    }
                    if (NCTA == 2) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);  // stage reusable once read
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (!WIDE || kb1 == num_kb) {   // (WIDE: inner chunks committed after their last main MMA above)
                    if (NCTA == 2) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);  // accumulator complete
                }
              }
            }
        }
    } else {
        // ===== epilogue warps 2..5 (2..9 when WIDE): TMEM lane group = warp % 4 (own 128 rows of the tile); WIDE: warps 2..5 take
        // columns [0, 128) of the tile, warps 6..9 columns [128, 256) =====
        const int lane_grp = warp & 3;
        const int col_off = WIDE ? ((warp - 2) >> 2) * (BN / 2) : 0;
        const uint32_t lead_empty0 = NCTA == 2 ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u;
        float accr[X3 ? BN / 2 : 1];   // X3: this thread's row of the tile, carried across the k chunks
        int it = 0;
        for (int tile = unit; tile < num_tiles; tile += num_units) {
          int tm, tn;
          tile_coords(tile, tm, tn);
          for (int kb0 = 0; kb0 < num_kb; kb0 += kb_chunk, ++it) {
            const int acc = WIDE ? 0 : (it & 1);
            const uint32_t acc_phase = (uint32_t)(WIDE ? it : (it >> 1)) & 1;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int64_t m = (int64_t)(tm * NCTA + (int)rank) * BM + lane_grp * 32 + lane;
            const int n0 = tn * bn;
            float *crow = C + m + (int64_t)n0 * ldc;
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BN);
            const bool full_tile = (int64_t)(tm * NCTA + (int)rank + 1) * BM <= M && n0 + bn <= N && (bn & 31) == 0;
            const bool add = kb0 != 0;   // later k chunks are added to what the first one produced (FP32, round to nearest)
            if (WIDE) {
                // 256-wide 3xTF32 tiles: this warp's 128 columns of the main accumulator per k chunk (the MMAs of the next chunk wait for
                // this drain, so it is four 32-column loads, not sixteen dependent ld + wait pairs), the correction accumulator once.
                const bool last = kb0 + kb_chunk >= num_kb;
#pragma unroll
                for (int c = 0; c < BN / 2; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + (uint32_t)(col_off + c), r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) accr[X3 ? c + j : 0] = add ? accr[X3 ? c + j : 0] + __uint_as_float(r[j]) : __uint_as_float(r[j]);
                }
                if (last) {
#pragma unroll
                    for (int c = 0; c < BN / 2; c += 32) {
                        uint32_t r[32];
                        tmem_ld32(taddr + (uint32_t)(CORR_OFF + col_off + c), r);
#pragma unroll
                        for (int j = 0; j < 32; ++j) accr[X3 ? c + j : 0] += __uint_as_float(r[j]);
                    }
                    if (m < M) {   // the only write of C
                        int64_t ld = ldc;
                        asm volatile("" : "+l"(ld));   // opaque here: keeps the 128 store addresses from being hoisted (and spilled)
                        float *pc = crow + (int64_t)col_off * ld;
                        const int ncols = min(bn, N - n0) - col_off;
#pragma unroll
                        for (int c = 0; c < BN / 2; ++c) {
                            if (c < ncols) *pc = accr[X3 ? c : 0];
                            pc += ld;
                        }
                    }
                }
            } else if (X3) {
                // 3xTF32: bn <= BN / 2.  The chunk sums live in registers (128 per thread) until the last chunk of the tile -- adding
                // them to C in global memory instead cost 8 GB of read-modify-write traffic per 8192^3 product (219 -> 236 TFLOP/s).
#pragma unroll
                for (int c = 0; c < BN / 2; c += 16) {   // 16 columns at a time: 128 carried sums + 32 in flight fit the register file
                    uint32_t r[16], rs[16];
                    tmem_ld16(taddr + (uint32_t)(col_off + c), r);
                    tmem_ld16(taddr + (uint32_t)(CORR_OFF + col_off + c), rs);   // the correction accumulator
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v = __uint_as_float(r[j]) + __uint_as_float(rs[j]);
                        accr[X3 ? c + j : 0] = add ? accr[X3 ? c + j : 0] + v : v;
                    }
                }
                if (kb0 + kb_chunk >= num_kb && m < M) {   // last chunk: the only write of C
                    int64_t ld = ldc;
                    asm volatile("" : "+l"(ld));   // opaque here: keeps the 128 store addresses from being hoisted out of the chunk loop (and spilled)
                    float *pc = crow + (int64_t)col_off * ld;
                    const int ncols = min(bn, N - n0) - col_off;
#pragma unroll
                    for (int c = 0; c < BN / 2; ++c) {
                        if (c < ncols) *pc = accr[X3 ? c : 0];
                        pc += ld;
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < bn; c += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + (uint32_t)c, r);   // warp-collective: every lane takes part, stores are predicated
                if (full_tile) {
                    if (add) {   // all 32 loads first: with a run-time ldc the compiler must assume the stores alias the next load
                        float old[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) old[j] = __ldcg(crow + (int64_t)(c + j) * ldc);
#pragma unroll
                        for (int j = 0; j < 32; ++j) crow[(int64_t)(c + j) * ldc] = old[j] + __uint_as_float(r[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) crow[(int64_t)(c + j) * ldc] = __uint_as_float(r[j]);
                    }
                } else if (m < M) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c + j < bn && n0 + c + j < N) {
                            float *pc = crow + (int64_t)(c + j) * ldc;
                            *pc = add ? *pc + __uint_as_float(r[j]) : __uint_as_float(r[j]);
                        }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {   // this warp has drained its quarter of the accumulator
                if (NCTA == 2) mbar_arrive_cluster(lead_empty0 + (uint32_t)(acc * 8)); else mbar_arrive(&tmem_empty[acc]);
            }
          }
        }
    }
    __syncwarp();   // role branches diverged inside the warp; the cluster barrier is warp-aligned
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if (NCTA == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}
}  // namespace tc

// ---- FP32 -> BF16 operand preparation (round to nearest even) ----------------------------------------
// transpose == 0: out[r + c*rows] = bf16(in[r + c*ld])        (B: K x N column-major is already K-contiguous)
// transpose != 0: out[c + r*cols] = bf16(in[r + c*ld])        (A: M x K column-major -> M rows of K)
__device__ __forceinline__ float4 ld_nc_v4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
// generic form: any rows / ld / alignment
__global__ void convert_bf16_kernel(__nv_bfloat16 *__restrict__ out, const float *__restrict__ in, int64_t rows,
                                    int64_t cols, int64_t ld) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i % rows, c = i / rows;
        out[i] = __float2bfloat16_rn(in[r + c * ld]);
    }
}
// contiguous form (ld == rows, 16-byte aligned, total % 8 == 0): two 128-bit loads -> one 128-bit store per item,
// four items in flight per thread
__global__ void __launch_bounds__(256) convert_bf16_vec_kernel(uint4 *__restrict__ out, const float4 *__restrict__ in,
                                                               int64_t n8) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n8; i += 4 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[2 * u] = ld_nc_v4(in + 2 * (i + u * stride));
            v[2 * u + 1] = ld_nc_v4(in + 2 * (i + u * stride) + 1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            out[i + u * stride] = make_uint4(pack_bf16x2(v[2 * u].x, v[2 * u].y), pack_bf16x2(v[2 * u].z, v[2 * u].w),
                                             pack_bf16x2(v[2 * u + 1].x, v[2 * u + 1].y), pack_bf16x2(v[2 * u + 1].z, v[2 * u + 1].w));
    }
    for (; i < n8; i += stride) {
        const float4 a = ld_nc_v4(in + 2 * i), b = ld_nc_v4(in + 2 * i + 1);
        out[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
}
// generic transposing form: 32 x 32 tiles, scalar accesses
__global__ void __launch_bounds__(256) convert_bf16_transpose_kernel(__nv_bfloat16 *__restrict__ out,
                                                                     const float *__restrict__ in, int64_t rows,
                                                                     int64_t cols, int64_t ld, int64_t tiles_r) {
    __shared__ float tile[32][33];
    const int64_t tr = blockIdx.x % tiles_r, tcn = blockIdx.x / tiles_r;
    const int64_t r0 = tr * 32, c0 = tcn * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 32; q += 8) {
        int64_t r = r0 + tx, c = c0 + ty + q;
        if (r < rows && c < cols) tile[ty + q][tx] = in[r + c * ld];  // tile[c_local][r_local]
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 32; q += 8) {
        int64_t c = c0 + tx, r = r0 + ty + q;
        if (r < rows && c < cols) out[c + r * cols] = __float2bfloat16_rn(tile[tx][ty + q]);
    }
}
// aligned transposing form (rows % 64 == 0, cols % 64 == 0, ld % 4 == 0, 16-byte aligned): 64 x 64 tiles; 128-bit
// coalesced loads along r into tile[c][r] (row pitch 68 words: 16-byte aligned rows, conflict-free column reads), then
// every thread converts 16 consecutive c of one r and writes one full 32-byte sector.
__global__ void __launch_bounds__(256) convert_bf16_transpose64_kernel(__nv_bfloat16 *__restrict__ out,
                                                                       const float *__restrict__ in, int64_t rows,
                                                                       int64_t cols, int64_t ld, int64_t tiles_r) {
    __shared__ __align__(16) float tile[64][68];
    const int64_t r0 = (blockIdx.x % tiles_r) * 64, c0 = (blockIdx.x / tiles_r) * 64;
    const int t = threadIdx.x;
    {
        const int rq = (t & 15) * 4, cl = t >> 4;   // 16 lanes x float4 = 64 r of one column; 16 columns per pass
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = ld_nc_v4(reinterpret_cast<const float4 *>(in + (r0 + rq) + (c0 + cl + 16 * q) * ld));
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4 *>(&tile[cl + 16 * q][rq]) = v[q];
    }
    __syncthreads();
    const int r = t & 63, cg = (t >> 6) * 16;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(tile[cg + 2 * j][r], tile[cg + 2 * j + 1][r]);
    uint4 *dst = reinterpret_cast<uint4 *>(out + (c0 + cg) + (r0 + r) * cols);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Columns per output tile: the persistent grid works through ceil(tiles / units) rounds of tiles whose duration is
// proportional to bn, so choose the bn that minimises rounds * bn -- e.g. a 1024 x 8192 row panel on 74 CTA pairs takes
// 2 rounds either way, 148 tiles of 256 x 224 instead of 128 tiles of 256 x 256.  Ties keep the widest tile (fewest
// re-reads of A); `tcgemm.bn` pins a value.  bn is a multiple of 16 per CTA of the unit.
static int pick_bn(int64_t M, int64_t N, int units, int ncta) {
    const int max_bn = tc::BN;
    const int step = 16 * ncta;
    const int forced = tune_get("tcgemm.bn", 0);
    if (forced >= step && forced <= max_bn && forced % step == 0) return forced;
    const int64_t tiles_m = cdiv(M, tc::BM * ncta);
    int best = max_bn;
    int64_t best_cost = INT64_MAX;
    for (int bn = max_bn; bn >= max_bn / 2; bn -= step) {
        const int64_t tiles = tiles_m * cdiv(N, bn);
        const int64_t cost = cdiv(tiles, units) * (bn + 8);   // + 8: per-tile fixed cost (pipeline refill, epilogue hand-over)
        if (cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

// One persistent CTA per SM; NCTA = 2 launches clusters of two (a CTA pair shares a TPC) -- chosen when the matrix has
// more than one 128-row tile (`tcgemm.cta2` = 0 forces single CTAs).
template <bool TF32, int NCTA, int X3>
static int launch_tcgen05(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &talo, const CUtensorMap &tblo, float *C, int64_t ldc,
                          int64_t M, int64_t N, int num_kb, int bn, int kb_chunk, cudaStream_t st) {
    using namespace tc;
    const int tiles_m = (int)cdiv(M, BM * NCTA), tiles_n = (int)cdiv(N, bn);
    int grid = sm_count() / NCTA * NCTA;
    if ((int64_t)grid > (int64_t)tiles_m * tiles_n * NCTA) grid = tiles_m * tiles_n * NCTA;
    auto kfn = gemm_tcgen05_kernel<TF32, NCTA, X3>;
    B200_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, PipeCfg<NCTA, X3>::SMEM_BYTES));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(PipeCfg<NCTA, X3>::THREADS);
    cfg.dynamicSmemBytes = PipeCfg<NCTA, X3>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kfn, ta, tb, talo, tblo, C, ldc, (int)M, (int)N, tiles_m, tiles_n, num_kb,
                                 tune_get("tcgemm.group_m", NCTA == 2 ? 8 : 16), bn, kb_chunk > 0 ? kb_chunk : num_kb));
    B200_CHECK_LAUNCH(X3 ? "gemm_tcgen05_kernel<tf32x3>" : (TF32 ? "gemm_tcgen05_kernel<tf32>" : "gemm_tcgen05_kernel<bf16>"));
    return B200_OK;
}
static int use_cta_pairs(int64_t M) { return tune_get("tcgemm.cta2", 1) != 0 && M > tc::BM; }

// [rows x K] bf16, K contiguous; box = 64 elements (128 bytes) x box_rows with 128B swizzle
static int make_tmap_bf16_kmajor(CUtensorMap *tm, const void *ptr, int64_t rows, int64_t K, int box_rows) {
    return pipe::make_tmap_2d(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, K, rows, K, tc::BK, box_rows,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

int gemm_bf16_launch(float *C, const uint16_t *A, const uint16_t *B, int64_t M, int64_t K, int64_t N, int64_t ldc,
                     cudaStream_t st) {
    using namespace tc;
    if (M <= 0 || N <= 0 || K <= 0 || K % 8 != 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return set_error(B200_ERR_UNSUPPORTED,
                         "b200_matmul_bf16: M, N, K must be positive and K a multiple of 8 (TMA row pitch) (got %lld, %lld, %lld)",
                         (long long)M, (long long)N, (long long)K);
    if ((((uintptr_t)A | (uintptr_t)B) & 15) != 0)
        return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_bf16: operands must be 16-byte aligned");
    CUtensorMap ta, tb;
    const int ncta = use_cta_pairs(M) ? 2 : 1;
    const int bn = pick_bn(M, N, sm_count() / ncta, ncta);
    B200_TRY(make_tmap_bf16_kmajor(&ta, A, M, K, BM));
    B200_TRY(make_tmap_bf16_kmajor(&tb, B, N, K, bn / ncta));
    if (ncta == 2) return launch_tcgen05<false, 2, 0>(ta, tb, ta, tb, C, ldc, M, N, (int)cdiv(K, BK), bn, 0, st);
    return launch_tcgen05<false, 1, 0>(ta, tb, ta, tb, C, ldc, M, N, (int)cdiv(K, BK), bn, 0, st);
}

// FP32 operands straight from HBM: A (M x K, lda) column-major -> boxes of 32 m x 32 k; B (K x N, ldb) -> 32 k x 256 n
int gemm_tf32_launch(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb,
                     int64_t ldc, cudaStream_t st) {
    using namespace tc;
    if (M <= 0 || N <= 0 || K <= 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return set_error(B200_ERR_UNSUPPORTED, "b200_matmul_tf32: M, N, K must be positive (got %lld, %lld, %lld)",
                         (long long)M, (long long)N, (long long)K);
    if ((lda % 4) || (ldb % 4) || ((((uintptr_t)A | (uintptr_t)B) & 15) != 0))
        return set_error(B200_ERR_UNSUPPORTED, "b200_matmul_tf32: TMA needs 16-byte aligned operands and leading dimensions that are multiples of 4");
    CUtensorMap ta, tb;
    const int ncta = use_cta_pairs(M) ? 2 : 1;
    const int bn = pick_bn(M, N, sm_count() / ncta, ncta);
    B200_TRY(pipe::make_tmap_2d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, M, K, lda, 32, BK_TF32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    B200_TRY(pipe::make_tmap_2d(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, B, K, N, ldb, BK_TF32, bn / ncta, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    if (ncta == 2) return launch_tcgen05<true, 2, 0>(ta, tb, ta, tb, C, ldc, M, N, (int)cdiv(K, BK_TF32), bn, 0, st);
    return launch_tcgen05<true, 1, 0>(ta, tb, ta, tb, C, ldc, M, N, (int)cdiv(K, BK_TF32), bn, 0, st);
}

// ---- 3xTF32: FP32-grade products on the tensor cores -------------------------------------------------------------------
// lo = x - (x with the 13 low mantissa bits cleared): the part of x the tensor core drops when it reads x as TF32; exact in FP32.
__global__ void __launch_bounds__(256) split_tf32_lo_kernel(float4 *__restrict__ lo, const float4 *__restrict__ x, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = ld_nc_v4(x + i);
        auto lo_of = [](float x) {   // Inf / NaN keep a zero remainder, so they propagate through the head term only (Inf - Inf would be NaN)
            const float d = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            return (__float_as_uint(x) & 0x7F800000u) == 0x7F800000u ? 0.0f : d;
        };
        float4 r;
        r.x = lo_of(v.x);
        r.y = lo_of(v.y);
        r.z = lo_of(v.z);
        r.w = lo_of(v.w);
        lo[i] = r;
    }
}

int gemm_tf32x3_launch(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb,
                       int64_t ldc, float *workspace, cudaStream_t st) {
    using namespace tc;
    if (M <= 0 || N <= 0 || K <= 0 || M > INT32_MAX || N > INT32_MAX || K > INT32_MAX)
        return set_error(B200_ERR_UNSUPPORTED, "b200_matmul_f32_tf32x3: M, N, K must be positive (got %lld, %lld, %lld)", (long long)M,
                         (long long)K, (long long)N);
    if ((lda % 4) || (ldb % 4) || ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)workspace) & 15) != 0))
        return set_error(B200_ERR_UNSUPPORTED, "b200_matmul_f32_tf32x3: TMA needs 16-byte aligned operands and leading dimensions that are multiples of 4");
    float *Alo = workspace, *Blo = workspace + lda * K;   // same shapes and leading dimensions as A and B
    const int sgrid = sm_count() * 8;
    split_tf32_lo_kernel<<<sgrid, 256, 0, st>>>((float4 *)Alo, (const float4 *)A, lda * K / 4);
    B200_CHECK_LAUNCH("split_tf32_lo_kernel");
    split_tf32_lo_kernel<<<sgrid, 256, 0, st>>>((float4 *)Blo, (const float4 *)B, ldb * N / 4);
    B200_CHECK_LAUNCH("split_tf32_lo_kernel");
    CUtensorMap ta, tb, talo, tblo;
    const int ncta = use_cta_pairs(M) ? 2 : 1;
    // CTA pairs: 256-wide tiles, one (main + correction) accumulator pair filling the 512 TMEM columns (X3 = 2; `tcgemm.x3_wide` = 0 / 1
    // pins the narrow / wide layout).  Single CTAs (M <= 128): two buffers of 128 + 128 columns -> tiles at most 128 wide (X3 = 1);
    // narrower wave-filling widths do not pay there (1024 x 8192 panel: bn 64 -> 113 TFLOP/s against 175 with 128): `tcgemm.bn` only.
    // (too few 256-wide tiles to occupy the CTA pairs -> the narrow layout: 256 x 8192 x 8192 measures 131 against 112 TFLOP/s)
    const int wide_knob = tune_get("tcgemm.x3_wide", -1);
    const bool wide = ncta == 2 && (wide_knob < 0 ? 2 * cdiv(M, 2 * BM) * cdiv(N, BN) >= sm_count() / 2 : wide_knob != 0);
    const int forced_bn = tune_get("tcgemm.bn", 0);
    const int max_bn = wide ? BN : BN / 2;
    const int bn = (forced_bn >= 16 * ncta && forced_bn <= max_bn && forced_bn % (16 * ncta) == 0) ? forced_bn
                   : wide                                                                         ? pick_bn(M, N, sm_count() / ncta, ncta)
                                                                                                  : max_bn;
    for (int lo = 0; lo < 2; ++lo) {
        B200_TRY(pipe::make_tmap_2d(lo ? &talo : &ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, lo ? Alo : A, M, K, lda, 32, BK_TF32,
                                    CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
        B200_TRY(pipe::make_tmap_2d(lo ? &tblo : &tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, lo ? Blo : B, K, N, ldb, BK_TF32, bn / ncta,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
    }
    // k blocks (of 32) per accumulation chunk: the tensor core's truncating accumulation is restarted every chunk
    // measured at 8192^3, uniform [0,1) inputs (max element-wise error vs Float64 / TFLOP/s): 256 -> 2e-6 / 230, 512 -> 3.6e-6 / 236,
    // 1024 -> 7e-6 / 239; the main accumulator loses ~7.4e-9 per k (one truncation of half an ulp per 8 k)
    int kb_chunk = tune_get("tcgemm.x3_chunk_k", 256) / BK_TF32;
    if (kb_chunk < 1) kb_chunk = 1;
    if (wide) return launch_tcgen05<true, 2, 2>(ta, tb, talo, tblo, C, ldc, M, N, (int)cdiv(K, BK_TF32), bn, kb_chunk, st);
    if (ncta == 2) return launch_tcgen05<true, 2, 1>(ta, tb, talo, tblo, C, ldc, M, N, (int)cdiv(K, BK_TF32), bn, kb_chunk, st);
    return launch_tcgen05<true, 1, 1>(ta, tb, talo, tblo, C, ldc, M, N, (int)cdiv(K, BK_TF32), bn, kb_chunk, st);
}

int convert_bf16_launch(uint16_t *out, const float *in, int64_t rows, int64_t cols, int64_t ld, int transpose,
                        cudaStream_t st) {
    if (rows == 0 || cols == 0) return B200_OK;
    const bool al16 = ((((uintptr_t)out | (uintptr_t)in) & 15) == 0);
    if (!transpose) {
        const int64_t total = rows * cols;
        if (ld == rows && al16 && total % 8 == 0) {
            int64_t blocks = cdiv(total / 8, 256 * 4);
            int grid = (int)(blocks < (int64_t)sm_count() * 8 ? blocks : (int64_t)sm_count() * 8);
            convert_bf16_vec_kernel<<<grid, 256, 0, st>>>((uint4 *)out, (const float4 *)in, total / 8);
            B200_CHECK_LAUNCH("convert_bf16_vec_kernel");
        } else {
            int grid = sm_count() * 8;
            convert_bf16_kernel<<<grid, 256, 0, st>>>((__nv_bfloat16 *)out, in, rows, cols, ld);
            B200_CHECK_LAUNCH("convert_bf16_kernel");
        }
    } else if (al16 && rows % 64 == 0 && cols % 64 == 0 && ld % 4 == 0 && (rows / 64) * (cols / 64) <= 2147483647LL) {
        convert_bf16_transpose64_kernel<<<(unsigned)((rows / 64) * (cols / 64)), 256, 0, st>>>((__nv_bfloat16 *)out, in, rows,
                                                                                             cols, ld, rows / 64);
        B200_CHECK_LAUNCH("convert_bf16_transpose64_kernel");
    } else {
        int64_t tiles_r = cdiv(rows, 32), tiles_c = cdiv(cols, 32);
        if (tiles_r * tiles_c > 2147483647LL) return set_error(B200_ERR_UNSUPPORTED, "convert: matrix too large");
        convert_bf16_transpose_kernel<<<(unsigned)(tiles_r * tiles_c), 256, 0, st>>>((__nv_bfloat16 *)out, in, rows, cols,
                                                                                     ld, tiles_r);
        B200_CHECK_LAUNCH("convert_bf16_transpose_kernel");
    }
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_matmul_bf16(float *C, const uint16_t *Abf, const uint16_t *Bbf, int64_t M, int64_t K,
                                        int64_t N, int64_t ldc) {
    if (!C || !Abf || !Bbf) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_bf16: NULL pointer");
    if (ldc < M) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_bf16: ldc < M");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return gemm_bf16_launch(C, Abf, Bbf, M, K, N, ldc, st);
}

extern "C" b200_status b200_matmul_tf32(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N,
                                        int64_t lda, int64_t ldb, int64_t ldc) {
    if (M < 0 || N < 0 || K < 0 || lda < M || ldb < K || ldc < M)
        return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_tf32: bad shape");
    if (M == 0 || N == 0) return B200_OK;
    if (!C || !A || !B) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_tf32: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return gemm_tf32_launch(C, A, B, M, K, N, lda, ldb, ldc, st);
}

extern "C" b200_status b200_matmul_f32_tf32x3(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N,
                                              int64_t lda, int64_t ldb, int64_t ldc, void *workspace) {
    if (M < 0 || N < 0 || K < 0 || lda < M || ldb < K || ldc < M)
        return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_tf32x3: bad shape");
    if (M == 0 || N == 0) return B200_OK;
    if (!C || !A || !B || !workspace) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_tf32x3: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return gemm_tf32x3_launch(C, A, B, M, K, N, lda, ldb, ldc, (float *)workspace, st);
}

extern "C" b200_status b200_convert_f32_bf16(uint16_t *out, const float *in, int64_t rows, int64_t cols, int64_t ld,
                                             int32_t transpose) {
    if (rows < 0 || cols < 0 || ld < rows) return set_error(B200_ERR_INVALID_VALUE, "b200_convert_f32_bf16: bad shape");
    if (rows == 0 || cols == 0) return B200_OK;
    if (!out || !in) return set_error(B200_ERR_INVALID_VALUE, "b200_convert_f32_bf16: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return convert_bf16_launch(out, in, rows, cols, ld, transpose, st);
}

extern "C" b200_status b200_matmul_f32_via_bf16(float *C, const float *A, const float *B, int64_t M, int64_t K,
                                                int64_t N, int64_t lda, int64_t ldb, int64_t ldc, void *workspace) {
    if (!C || !A || !B || !workspace) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_via_bf16: NULL pointer");
    if (lda < M || ldb < K || ldc < M) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_via_bf16: bad ld");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    uint16_t *abf = (uint16_t *)workspace;
    uint16_t *bbf = abf + M * K;
    B200_TRY(convert_bf16_launch(abf, A, M, K, lda, 1, st));  // A: M x K col-major -> M rows of K
    B200_TRY(convert_bf16_launch(bbf, B, K, N, ldb, 0, st));  // B: K x N col-major == N rows of K
    return gemm_bf16_launch(C, abf, bbf, M, K, N, ldc, st);
}

// The same product with B already converted and RESIDENT (b200_convert_f32_bf16(Bbf, B, K, N, ldb, 0), once): what a sharded
// run does -- B is replicated at set-up and only the rank's A row panel changes per call.  workspace: M * K
// bf16 for the converted panel.  Same bits as b200_matmul_f32_via_bf16 on the same operands.
extern "C" b200_status b200_matmul_f32_bf16_resident_b(float *C, const float *A, const uint16_t *Bbf, int64_t M, int64_t K,
                                                       int64_t N, int64_t lda, int64_t ldc, void *workspace) {
    if (!C || !A || !Bbf || !workspace) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_bf16_resident_b: NULL pointer");
    if (lda < M || ldc < M) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_bf16_resident_b: bad ld");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    uint16_t *abf = (uint16_t *)workspace;
    B200_TRY(convert_bf16_launch(abf, A, M, K, lda, 1, st));  // A: M x K col-major -> M rows of K
    return gemm_bf16_launch(C, abf, Bbf, M, K, N, ldc, st);
}
