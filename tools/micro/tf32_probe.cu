// Probe for the tcgen05 kind::tf32 MMA with an MN-major A operand staged by TMA (SWIZZLE_128B).
// One CTA: TMA loads A (128 m x 32 k, column-major FP32 => four [k][32 m] slabs) and B (32 k x 256 n => [n][32 k]),
// dumps the shared-memory image, then runs 4 x tcgen05.mma 128x256x8 with the descriptor variant given on the
// command line and dumps the TMEM accumulator.  The host checks D against A*B and prints which variants match.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tf32_probe tf32_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant {
    uint32_t a_lbo, a_sbo, a_kstep;   // bytes
    uint32_t idesc;
    uint32_t a_swz, a_slab;           // layout_type field for A, bytes between TMA slabs
    uint32_t use32;                   // A staged with the 32B-atom swizzle tensor map
};

__global__ void __launch_bounds__(128, 1)
    probe(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap ta32, const __grid_constant__ CUtensorMap tb,
          float *smem_dump, float *D,
          Variant v) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char *sa = smem, *sb = smem + A_BYTES;
    uint64_t *bar = (uint64_t *)(smem + A_BYTES + B_BYTES);
    uint64_t *bar2 = bar + 1;
    uint32_t *tslot = (uint32_t *)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(A_BYTES + B_BYTES) : "memory");
        for (int g = 0; g < 4; ++g)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                    "r"(smem_u32(sa + g * v.a_slab)),
                "l"(v.use32 ? &ta32 : &ta), "r"(32 * g), "r"(0), "r"(smem_u32(bar))
                : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(smem_u32(sb)),
            "l"(&tb), "r"(0), "r"(0), "r"(smem_u32(bar))
            : "memory");
    }
    // everyone waits for the data
    asm volatile(
        "{\n.reg .pred p;\nW1_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1_%=;\nbra W1_%=;\nD1_%=:\n}\n" ::"r"(
            smem_u32(bar))
        : "memory");
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) smem_dump[i] = ((float *)smem)[i];
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint64_t da = (uint64_t)((smem_u32(sa) >> 4) & 0x3FFF) | ((uint64_t)(v.a_lbo >> 4) << 16) |
                            ((uint64_t)(v.a_sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)v.a_swz << 61);
        const uint64_t db = (uint64_t)((smem_u32(sb) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                            ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        for (int k = 0; k < 4; ++k) {
            uint32_t acc = k != 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                "l"(da + (uint64_t)((v.a_kstep >> 4) * k)), "l"(db + (uint64_t)(2 * k)), "r"(v.idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar2)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW2_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2_%=;\nbra W2_%=;\nD2_%=:\n}\n" ::"r"(
            smem_u32(bar2))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = warp * 32 + lane;
    for (int c = 0; c < BN; c += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[m + (size_t)(c + j) * BM] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)p;
    const int M = BM, K = BK, N = BN;
    std::vector<float> A(M * K), B(K * N), Dh(M * N), dump((A_BYTES + B_BYTES) / 4);
    for (int k = 0; k < K; ++k)
        for (int m = 0; m < M; ++m) A[m + k * M] = (float)((m * 7 + k * 3) % 31 - 15);   // column-major M x K
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) B[k + n * K] = (float)((k * 5 + n * 11) % 29 - 14);  // column-major K x N
    std::vector<float> want(M * N);
    for (int n = 0; n < N; ++n)
        for (int m = 0; m < M; ++m) {
            float s = 0;
            for (int k = 0; k < K; ++k) s += A[m + k * M] * B[k + n * K];
            want[m + n * M] = s;
        }
    float *dA, *dB, *dD, *dDump;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, Dh.size() * 4));
    CK(cudaMalloc(&dDump, dump.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap ta, tb, ta32;
    {
        cuuint64_t gd[2] = {(cuuint64_t)M, (cuuint64_t)K}, gs[1] = {(cuuint64_t)M * 4};
        cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        CUresult r = enc(&ta32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode A32: %d\n", (int)r);
    }
    {
        cuuint64_t gd[2] = {(cuuint64_t)M, (cuuint64_t)K}, gs[1] = {(cuuint64_t)M * 4};
        cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode A: %d\n", (int)r);
    }
    {
        cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)N}, gs[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {32, 256}, es[2] = {1, 1};
        CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode B: %d\n", (int)r);
    }
    const int smem = A_BYTES + B_BYTES + 1024 + 64;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const uint32_t base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    struct Named {
        const char *name;
        Variant v;
    } vars[] = {
        {"mn lbo4096 sbo1024 kstep1024", {4096, 1024, 1024, base | (1u << 15), 2, 4096, 0}},
        {"mn lbo1024 sbo4096 kstep1024", {1024, 4096, 1024, base | (1u << 15), 2, 4096, 0}},
        {"mn lbo4096 sbo1024 kstep1024 (a_major bit16 instead)", {4096, 1024, 1024, base | (1u << 16), 2, 4096, 0}},
        {"K-major flag with MN data (expect wrong)", {16, 1024, 32, base, 2, 4096, 0}},
        {"mn32 lbo4096 sbo512 kstep1024", {4096, 512, 1024, base | (1u << 15), 1, 4096, 1}},
        {"mn32 lbo512 sbo4096 kstep1024", {512, 4096, 1024, base | (1u << 15), 1, 4096, 1}},
        {"mn32 lbo4096 sbo1024 kstep1024", {4096, 1024, 1024, base | (1u << 15), 1, 4096, 1}},
        {"mn32 lbo4096 sbo512 kstep1024 (std sw128 image)", {4096, 512, 1024, base | (1u << 15), 1, 4096, 0}},
    };
    for (auto &nv : vars) {
        CK(cudaMemset(dD, 0xFF, Dh.size() * 4));
        probe<<<1, 128, smem>>>(ta, ta32, tb, dDump, dD, nv.v);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%-55s: launch failed: %s\n", nv.name, cudaGetErrorString(e));
            return 1;
        }
        CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost));
        if (nv.v.use32) {
            long b32 = 0;   // 32-byte chunk c2 of row k holds m = 32g + 8*(c2 ^ (k & 3)) .. +7
            for (int g = 0; g < 4; ++g)
                for (int k = 0; k < 32; ++k)
                    for (int c2 = 0; c2 < 4; ++c2)
                        for (int e = 0; e < 8; ++e) b32 += dump[g * 1024 + k * 32 + c2 * 8 + e] != A[32 * g + 8 * (c2 ^ (k & 3)) + e + k * M];
            printf("   32B-atom smem image mismatches %ld; row1 = %g %g %g %g %g %g %g %g | %g (A[0..,1] = %g %g; A[8,1] = %g)\n", b32, dump[32], dump[33],
                   dump[34], dump[35], dump[36], dump[37], dump[38], dump[39], dump[40], A[M], A[M + 1], A[M + 8]);
        }
        long bad = 0, zeros = 0;
        for (size_t i = 0; i < Dh.size(); ++i) {
            bad += Dh[i] != want[i];
            zeros += Dh[i] == 0.0f;
        }
        printf("%-55s: mismatches %ld / %zu, zeros %ld;  D[0..3,0] = %g %g %g %g  want %g %g %g %g\n", nv.name, bad, Dh.size(),
               zeros, Dh[0], Dh[1], Dh[2], Dh[3], want[0], want[1], want[2], want[3]);
    }
    // shared-memory image check: A slab g, row k (128 B), 16-byte chunk c holds m = 32g + 4*(c ^ (k & 7)) .. +3
    long badA = 0, badB = 0;
    for (int g = 0; g < 4; ++g)
        for (int k = 0; k < 32; ++k)
            for (int c = 0; c < 8; ++c)
                for (int e = 0; e < 4; ++e) {
                    int m = 32 * g + 4 * (c ^ (k & 7)) + e;
                    badA += dump[g * 1024 + k * 32 + c * 4 + e] != A[m + k * M];
                }
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < 8; ++c)
            for (int e = 0; e < 4; ++e) {
                int k = 4 * (c ^ (n & 7)) + e;
                badB += dump[A_BYTES / 4 + n * 32 + c * 4 + e] != B[k + n * K];
            }
    printf("smem image: A mismatches %ld, B mismatches %ld;  A smem[0..7] = %g %g %g %g %g %g %g %g (A[0..7,0] = %g %g ...)\n", badA,
           badB, dump[0], dump[1], dump[2], dump[3], dump[4], dump[5], dump[6], dump[7], A[0], A[1]);
    return 0;
}
