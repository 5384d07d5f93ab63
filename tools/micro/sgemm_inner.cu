// Micro-benchmark: the inner k loop of sgemm_tma_kernel in isolation (no TMA, no mbarriers, operands resident in shared
// memory), to find which ingredient costs FMA-pipe cycles.  8 warps per CTA, 1 CTA per SM, thread tile 8 x 8 as in the kernel.
//   V0: kernel loop as is (A: 2 LDS.128 per k, B: 8 LDS.128 per 4 k, FFMA2 with scalar-broadcast B, TILE16 totals)
//   V1: V0 without the per-tile FADD2 totals (running sum)
//   V2: V1 with B taken from registers (only the A loads remain)
//   V3: V1 with A taken from registers (only the B loads remain)
//   V4: no shared-memory loads at all (operands rotate in registers)
//   V5: 16 x 8 thread tile, running sum (A: 4 LDS.128 per k, B: 8 LDS.128 per 4 k; 10.7 FFMA2 per LDS.128)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sgemm_inner sgemm_inner.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int BM = 128, BN = 128, BK = 16;

__device__ __forceinline__ float4 lds128(const float *p) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return r;
}
__device__ __forceinline__ float2 lds64(const float *p) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return r;
}

template <int V>
__global__ void __launch_bounds__(256, 1) inner(const float *in, float *out, int tiles) {
    __shared__ __align__(16) float As_[BK * BM];
    __shared__ __align__(16) float Bs_[BN * BK];
    for (int i = threadIdx.x; i < BK * BM; i += 256) As_[i] = in[i];
    for (int i = threadIdx.x; i < BN * BK; i += 256) Bs_[i] = in[BK * BM + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int lx = (lane >> 1) & 7, ly = (lane & 1) | ((lane >> 4) << 1);
    const float *As = As_ + wm * 64 + 4 * lx;
    const float *Bs = Bs_ + (wn * 32 + ly) * BK;
    float2 acc[4][8], tot[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f), tot[i][j] = make_float2(-0.f, -0.f);
    float4 ra0 = *reinterpret_cast<const float4 *>(As), ra1 = *reinterpret_cast<const float4 *>(As + 32);
    float2 rb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rb[j] = *reinterpret_cast<const float2 *>(Bs + 4 * j * BK);
    for (int kt = 0; kt < tiles; ++kt) {
#pragma unroll
        for (int k2 = 0; k2 < BK; k2 += 2) {
            float2 b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (V == 2 || V == 4) b[j] = rb[j];
                else b[j] = lds64(Bs + (4 * j) * BK + k2);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = k2 + h;
                float4 a0, a1;
                if (V == 3 || V == 4) { a0 = ra0; a1 = ra1; }
                else {
                    a0 = lds128(As + kk * BM);
                    a1 = lds128(As + kk * BM + 32);
                }
                const float2 a[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float bj = h ? b[j].y : b[j].x;
                        const float2 c = (V == 0 && kk == 0) ? make_float2(0.f, 0.f) : acc[i][j];
                        acc[i][j] = __ffma2_rn(a[i], make_float2(bj, bj), c);
                    }
            }
            if (V == 2 || V == 4) {   // rotate so nothing is loop invariant
                float2 t0 = rb[0];
#pragma unroll
                for (int j = 0; j < 7; ++j) rb[j] = rb[j + 1];
                rb[7] = t0;
            }
            if (V == 3 || V == 4) { float4 t0 = ra0; ra0 = ra1; ra1 = t0; }
        }
        if (V == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) tot[i][j] = __fadd2_rn(tot[i][j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y + tot[i][j].x + tot[i][j].y;
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// V5: thread tile 16 (m) x 8 (n): acc[8][8] float2; CTA tile would be 256 x 128 with the same 8 warps
__global__ void __launch_bounds__(256, 1) inner16(const float *in, float *out, int tiles) {
    __shared__ __align__(16) float As_[BK * 256];
    __shared__ __align__(16) float Bs_[BN * BK];
    for (int i = threadIdx.x; i < BK * 256; i += 256) As_[i] = in[i % (BK * BM)];
    for (int i = threadIdx.x; i < BN * BK; i += 256) Bs_[i] = in[BK * BM + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int lx = (lane >> 1) & 7, ly = (lane & 1) | ((lane >> 4) << 1);
    const float *As = As_ + wm * 128 + 4 * lx;
    const float *Bs = Bs_ + (wn * 32 + ly) * BK;
    float2 acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int kt = 0; kt < tiles; ++kt) {
#pragma unroll
        for (int k2 = 0; k2 < BK; k2 += 2) {
            float2 b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = lds64(Bs + (4 * j) * BK + k2);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = k2 + h;
                float2 a[8];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 av = lds128(As + kk * 256 + 32 * g);
                    a[2 * g] = make_float2(av.x, av.y);
                    a[2 * g + 1] = make_float2(av.z, av.w);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float bj = h ? b[j].y : b[j].x;
                        acc[i][j] = __ffma2_rn(a[i], make_float2(bj, bj), acc[i][j]);
                    }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

// V6: no loads, no scalar-broadcast operand: acc = ffma2(a_pair, b_pair, acc) with genuinely two-valued b pairs
// V7: the same instruction count with the scalar-broadcast B form (b.x == b.y), for a like-for-like comparison
template <int BROADCAST>
__global__ void __launch_bounds__(256, 1) pairs(const float *in, float *out, int iters) {
    float2 a[4], b[8], acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = make_float2(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float x = in[threadIdx.x + 64 * j], y = in[threadIdx.x + 64 * j + 3];
        b[j] = BROADCAST ? make_float2(x, x) : make_float2(x, y);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = __ffma2_rn(a[i], b[j], acc[i][j]);
        float2 t = a[0];
#pragma unroll
        for (int i = 0; i < 3; ++i) a[i] = a[i + 1];
        a[3] = t;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

int main() {
    float *in, *out;
    cudaMalloc(&in, (BK * BM + BN * BK) * sizeof(float));
    cudaMemset(in, 0, (BK * BM + BN * BK) * sizeof(float));
    cudaMalloc(&out, 148 * 256 * sizeof(float));
    const int tiles = 40000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[5] = {"V0 kernel loop (TILE16)", "V1 running sum", "V2 A loads only", "V3 B loads only", "V4 no shared loads"};
    for (int v = 0; v < 5; ++v)
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) inner<0><<<148, 256>>>(in, out, tiles);
            if (v == 1) inner<1><<<148, 256>>>(in, out, tiles);
            if (v == 2) inner<2><<<148, 256>>>(in, out, tiles);
            if (v == 3) inner<3><<<148, 256>>>(in, out, tiles);
            if (v == 4) inner<4><<<148, 256>>>(in, out, tiles);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double flop = 2.0 * 64 * 256 * BK * (double)tiles * 148;
            if (rep == 1) printf("%-28s %.3f ms  %.2f TFLOP/s  (%.1f %% of 74.45) %s\n", names[v], ms, flop / ms * 1e-9, flop / ms * 1e-9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
        }
    for (int v = 0; v < 2; ++v)
        for (int rep = 0; rep < 2; ++rep) {
            const int iters = 400000;
            cudaEventRecord(e0);
            if (v == 0) pairs<0><<<148, 256>>>(in, out, iters);
            else pairs<1><<<148, 256>>>(in, out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double flop = 2.0 * 64 * 256 * (double)iters * 148;
            if (rep == 1) printf("%-28s %.3f ms  %.2f TFLOP/s  (%.1f %% of 74.45) %s\n", v == 0 ? "V6 pair x pair (no .F32)" : "V7 pair x broadcast (.F32)", ms, flop / ms * 1e-9,
                                 flop / ms * 1e-9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
        }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        inner16<<<148, 256>>>(in, out, tiles / 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flop = 2.0 * 128 * 256 * BK * (double)(tiles / 2) * 148;
        if (rep == 1) printf("%-28s %.3f ms  %.2f TFLOP/s  (%.1f %% of 74.45) %s\n", "V5 16x8 tile, running sum", ms, flop / ms * 1e-9, flop / ms * 1e-9 / 74.45 * 100, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
