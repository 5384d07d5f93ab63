// Micro-benchmark: tcgen05.ld / tcgen05.st throughput (registers <-> TMEM) for one CTA of 128 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: ld only, 1: st only, 2: ld + add + st (RMW)
__global__ void __launch_bounds__(128, 1) tmem_kernel(float *out, long long *cycles, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    uint32_t r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((float)(threadIdx.x + j));
    // initialise 128 columns
    for (int c = 0; c < 128; c += 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(base + c),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                     "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 128; c += 16) {
            uint32_t v[16];
            if (MODE != 1) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(base + c));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = r[j] + it;
            }
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += __uint_as_float(v[j]);
            } else {
                if (MODE == 2) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + 1.0f);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(base + c),
                             "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                             "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
            }
        }
        if (MODE != 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    if (MODE != 0) {
        uint32_t v0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc = __uint_as_float(v0);
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 128 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 1000;
    const char *names[3] = {"ld only", "st only", "ld+add+st"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int grid : {1, 148}) {
            if (mode == 0) tmem_kernel<0><<<grid, 128>>>(out, cyc, iters);
            if (mode == 1) tmem_kernel<1><<<grid, 128>>>(out, cyc, iters);
            if (mode == 2) tmem_kernel<2><<<grid, 128>>>(out, cyc, iters);
            cudaError_t e = cudaDeviceSynchronize();
            long long h = 0;
            cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            float o = 0;
            cudaMemcpy(&o, out, sizeof(o), cudaMemcpyDeviceToHost);
            double bytes = (double)iters * 128 * 128 * 4;
            printf("%-10s grid %3d: %lld cycles for %d x 64 KiB -> %.1f B/cycle/SM  (%s, out0=%g)\n", names[mode], grid, h, iters,
                   bytes / (double)h, cudaGetErrorString(e), o);
        }
    }
    return 0;
}
