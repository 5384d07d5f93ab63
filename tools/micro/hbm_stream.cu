// Micro-benchmark: what one direction of HBM delivers on its own.  Read-only and write-only streams over a 1 GiB buffer:
//   mode 0: TMA bulk loads global -> shared (ring of 4 stages, data discarded)      chunk KiB x CTAs per SM
//   mode 1: LDG.128 grid-stride reads, XOR-reduced (one store per thread at the end)  256 threads x CTAs per SM
//   mode 2: TMA bulk stores shared -> global of a constant chunk                     chunk KiB x CTAs per SM
//   mode 3: half of the CTAs read one buffer, the other half write another (independent streams)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_stream hbm_stream.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) bulk_read(const unsigned char *src, size_t bytes, uint32_t chunk, int stages) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[8];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t nchunks = bytes / chunk;
    const size_t mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](size_t i) {
        const int s = (int)(i % stages);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * chunk)),
                     "l"(src + (blockIdx.x + i * (size_t)gridDim.x) * chunk), "r"(chunk), "r"(smem_u32(&full[s]))
                     : "memory");
    };
    for (size_t i = 0; i < mine && i < (size_t)stages; ++i) issue(i);
    for (size_t i = 0; i < mine; ++i) {
        const int s = (int)(i % stages);
        const uint32_t parity = (uint32_t)((i / stages) & 1);
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&full[s])), "r"(parity) : "memory");
        if (i + stages < mine) issue(i + stages);
    }
}

__global__ void __launch_bounds__(256) ldg_read(const uint4 *src, size_t n16, uint4 *sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w;
        }
    }
    if (acc.x == 0x12345678u && acc.y == 1u) sink[threadIdx.x] = acc;   // never true for the test pattern: keeps the loads alive
}

__global__ void __launch_bounds__(32) bulk_write(unsigned char *dst, size_t bytes, uint32_t chunk) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (uint32_t i = threadIdx.x; i < chunk / 16; i += 32) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x != 0) return;
    for (size_t off = (size_t)blockIdx.x * chunk; off < bytes; off += (size_t)gridDim.x * chunk) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_u32(smem)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mode 3: even CTAs stream-read `src`, odd CTAs stream-write `dst` (independent streams, no data dependence): the mixed-traffic ceiling
__global__ void __launch_bounds__(32) bulk_mixed(const unsigned char *src, unsigned char *dst, size_t bytes, uint32_t chunk, int stages) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[8];
    const unsigned role = blockIdx.x & 1, id = blockIdx.x >> 1, n = gridDim.x >> 1;
    if (role == 1) {
        for (uint32_t i = threadIdx.x; i < chunk / 16; i += 32) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(1, 2, 3, 4);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (threadIdx.x != 0) return;
        for (size_t off = (size_t)id * chunk; off < bytes; off += (size_t)n * chunk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_u32(smem)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }
    if (threadIdx.x != 0) return;
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t nchunks = bytes / chunk;
    const size_t mine = nchunks > id ? (nchunks - id + n - 1) / n : 0;
    auto issue = [&](size_t i) {
        const int s = (int)(i % stages);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * chunk)),
                     "l"(src + (id + i * (size_t)n) * chunk), "r"(chunk), "r"(smem_u32(&full[s]))
                     : "memory");
    };
    for (size_t i = 0; i < mine && i < (size_t)stages; ++i) issue(i);
    for (size_t i = 0; i < mine; ++i) {
        const int s = (int)(i % stages);
        const uint32_t parity = (uint32_t)((i / stages) & 1);
        asm volatile("{\n.reg .pred p;\nW2: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(&full[s])), "r"(parity) : "memory");
        if (i + stages < mine) issue(i + stages);
    }
}

int main() {
    const size_t bytes = (size_t)1 << 30;
    unsigned char *buf[2];
    for (auto &b : buf) {
        cudaMalloc(&b, bytes);
        cudaMemset(b, 0x5a, bytes);
    }
    uint4 *sink;
    cudaMalloc(&sink, 4096);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](const char *name, int kb, int cps, auto launch) {
        for (int w = 0; w < 3; ++w) launch(buf[w & 1]);
        cudaDeviceSynchronize();
        const int iters = 10;
        cudaEventRecord(e0);
        for (int it = 0; it < iters; ++it) launch(buf[it & 1]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        printf("%s chunk_kb=%d ctas_per_sm=%d: %.1f us, %.0f GB/s%s\n", name, kb, cps, ms / iters * 1e3, bytes / (ms / iters) / 1e6, err ? " ERROR" : "");
    };
    for (int kb : {16, 32, 48}) {
        for (int cps : {1, 2, 3, 4, 8, 16, 32}) {
            const uint32_t chunk = kb << 10;
            const int stages = 4;
            cudaFuncSetAttribute(bulk_read, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * chunk);
            run("bulk_read ", kb, cps, [&](unsigned char *b) { bulk_read<<<sms * cps, 32, stages * chunk>>>(b, bytes, chunk, stages); });
        }
    }
    for (int cps : {2, 4, 8, 16, 32, 64})
        run("ldg_read  ", 0, cps, [&](unsigned char *b) { ldg_read<<<sms * cps, 256>>>((const uint4 *)b, bytes / 16, sink); });
    for (int kb : {32, 64}) {
        for (int cps : {8, 16, 32}) {
            const uint32_t chunk = kb << 10;
            cudaFuncSetAttribute(bulk_write, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
            run("bulk_write", kb, cps, [&](unsigned char *b) { bulk_write<<<sms * cps, 32, chunk>>>(b, bytes, chunk); });
        }
    }
    for (int kb : {16, 32}) {
        for (int cps : {2, 4, 8, 16, 32}) {
            const uint32_t chunk = kb << 10;
            const int stages = 4;
            cudaFuncSetAttribute(bulk_mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * chunk);
            // 1 GiB read + 1 GiB written per launch: the printed figure counts `bytes` once, so double it
            run("bulk_mixed (x2 = total traffic)", kb, cps, [&](unsigned char *b) { bulk_mixed<<<sms * cps, 32, stages * chunk>>>(b, buf[b == buf[0] ? 1 : 0], bytes, chunk, stages); });
        }
    }
    return 0;
}
