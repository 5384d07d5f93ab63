// Micro-benchmark: FP32 FMA-pipe ceiling on sm_100a for the instruction shapes an 8x8 register-tiled SGEMM can use.
//   mode 0: 64 scalar FFMA per step (acc[i][j] += a[i] * b[j])
//   mode 1: 32 FFMA2 per step, b pre-duplicated outside the loop (pure pipe ceiling of the packed form)
//   mode 2: 32 FFMA2 per step, b duplicated inside the loop (8 extra 64-bit packs per step, as a GEMM would need)
//   mode 3: 32 FFMA2 + 2 FADD2-free variant with 16 independent chains only (latency exposure check)
// Reports TFLOP/s at 1, 2, 4 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) fma_kernel(const float *in, float *out, int iters) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = in[threadIdx.x + 32 * i];
        b[i] = in[threadIdx.x + 32 * (8 + i)];
    }
    if (MODE == 0) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            // rotate operands so the loop body cannot be hoisted
            float t = a[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) a[i] = a[i + 1];
            a[7] = t;
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += acc[i][j];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        float2 acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
        float2 a2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a2[i] = make_float2(a[2 * i], a[2 * i + 1]);
        float2 bd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bd[j] = make_float2(b[j], b[j]);
        for (int it = 0; it < iters; ++it) {
            if (MODE == 2) {
                // fresh scalars every step -> the pack {b,b} has to be rebuilt (what the GEMM inner loop pays)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float bj = a2[j & 3].x + b[j];   // cheap dependency on rotating data
                    b[j] = bj;
                    bd[j] = make_float2(bj, bj);
                }
            }
            constexpr int NJ = (MODE == 3) ? 2 : 8;
#pragma unroll
            for (int r = 0; r < 8 / NJ; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[i][j] = __ffma2_rn(a2[i], bd[j + (MODE == 3 ? 0 : 0)], acc[i][j]);
            float2 t = a2[0];
#pragma unroll
            for (int i = 0; i < 3; ++i) a2[i] = a2[i + 1];
            a2[3] = t;
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += acc[i][j].x + acc[i][j].y;
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 1024 * 16 * sizeof(float));
    cudaMemset(in, 0, 1024 * 16 * sizeof(float));
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[4] = {"FFMA x64", "FFMA2 x32 (b pre-dup)", "FFMA2 x32 (+8 packs/step)", "FFMA2 x32, 8 chains"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int threads : {128, 256, 512}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) fma_kernel<0><<<148, threads>>>(in, out, iters);
                if (mode == 1) fma_kernel<1><<<148, threads>>>(in, out, iters);
                if (mode == 2) fma_kernel<2><<<148, threads>>>(in, out, iters);
                if (mode == 3) fma_kernel<3><<<148, threads>>>(in, out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                double flop = 2.0 * 64 * (double)iters * threads * 148;
                if (rep == 1)
                    printf("%-28s %d warps/SMSP: %.3f ms  %.2f TFLOP/s (%s)\n", names[mode], threads / 128, ms,
                           flop / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
