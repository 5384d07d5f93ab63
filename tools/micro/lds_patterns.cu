// Micro-benchmark: shared-memory wavefronts per LDS for the lane->address patterns a register-tiled SGEMM can use.
// Run under `ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum`.
#include <cstdio>
#include <cuda_runtime.h>

template <int WIDTH>  // bytes per lane: 4, 8, 16
__global__ void lds_kernel(const int *lane_chunk, float *out, int iters) {
    __shared__ __align__(16) float smem[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) smem[i] = (float)i;
    __syncthreads();
    const int chunk = lane_chunk[threadIdx.x & 31];
    const char *base = reinterpret_cast<const char *>(smem) + chunk * WIDTH;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const char *p = base + (it & 7) * 512;
        if (WIDTH == 16) {
            float4 v = *reinterpret_cast<const float4 *>(p);
            acc += v.x + v.y + v.z + v.w;
        } else if (WIDTH == 8) {
            float2 v = *reinterpret_cast<const float2 *>(p);
            acc += v.x + v.y;
        } else {
            acc += *reinterpret_cast<const float *>(p);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    int *d_map;
    float *d_out;
    cudaMalloc(&d_map, 32 * sizeof(int));
    cudaMalloc(&d_out, 1024 * sizeof(float));
    struct Pat { const char *name; int (*f)(int); };
    Pat pats[] = {
        {"l%8   (8 distinct, interleaved: A-pattern)", [](int l) { return l % 8; }},
        {"l/4   (8 distinct, groups of 4)", [](int l) { return l / 4; }},
        {"l/8   (4 distinct, one per quarter: B-pattern)", [](int l) { return l / 8; }},
        {"l%4   (4 distinct, interleaved)", [](int l) { return l % 4; }},
        {"(l/2)%4 (4 distinct, pairs)", [](int l) { return (l / 2) % 4; }},
        {"l/16  (2 distinct, one per half)", [](int l) { return l / 16; }},
        {"l%2   (2 distinct, interleaved)", [](int l) { return l % 2; }},
        {"0     (1 distinct)", [](int l) { return 0; }},
        {"l%16  (16 distinct, interleaved)", [](int l) { return l % 16; }},
        {"l/2   (16 distinct, pairs)", [](int l) { return l / 2; }},
        {"l     (32 distinct)", [](int l) { return l; }},
        {"(l%8)*2 (8 distinct, stride 2 chunks)", [](int l) { return (l % 8) * 2; }},
        {"(l/8)*8 (4 distinct, 128B apart in units)", [](int l) { return (l / 8) * 8; }},
        // which lane bits may differ inside one wavefront?  (partner l^1 vs l^2 merging)
        {"(l>>1)&7 (8 distinct, bits 1-3)", [](int l) { return (l >> 1) & 7; }},
        {"bits 0,4 (4 distinct)", [](int l) { return (l & 1) | ((l >> 4) << 1); }},
        {"bits 0,3,4 (8 distinct)", [](int l) { return (l & 1) | ((l >> 3) << 1); }},
        {"bits 0,2,3,4 (16 distinct)", [](int l) { return (l & 1) | (((l >> 2) & 7) << 1); }},
        {"bits 2,4 (4 distinct)", [](int l) { return ((l >> 2) & 1) | ((l >> 4) << 1); }},
        {"bits 0,2 (4 distinct)", [](int l) { return (l & 1) | (((l >> 2) & 1) << 1); }},
        {"bits 1,2,3,4 (16 distinct) = l/2", [](int l) { return l >> 1; }},
        {"bits 2,3,4 (8 distinct) = l/4", [](int l) { return l >> 2; }},
        {"bits 0,1 + rows 64B apart: (l&3)*4", [](int l) { return (l & 3) * 4; }},
        {"bits 3,4 rows 64B apart: (l>>3)*4", [](int l) { return (l >> 3) * 4; }},
        {"bits 3,4 rows 64B apart swz64: ", [](int l) { int n = l >> 3; return n * 4 + ((n >> 1) & 3); }},
    };
    for (auto &p : pats) {
        int h[32];
        for (int l = 0; l < 32; ++l) h[l] = p.f(l);
        cudaMemcpy(d_map, h, sizeof(h), cudaMemcpyHostToDevice);
        // one warp, 1024 loads each: wavefronts / 1024 = wavefronts per instruction
        lds_kernel<16><<<1, 32>>>(d_map, d_out, 1024);
        lds_kernel<8><<<1, 32>>>(d_map, d_out, 1024);
        lds_kernel<4><<<1, 32>>>(d_map, d_out, 1024);
        cudaDeviceSynchronize();
        printf("pattern %s: launched w16,w8,w4\n", p.name);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
