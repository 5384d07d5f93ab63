// ka_device.cuh -- device-side restatement of KA's iteration space, written once for every kernel twin.
//
// KaCtx is the by-value image of CompilerMetadata{ndrange, iterspace} (reference src/compiler.jl:1-33).
// A KA launch is ALWAYS one-dimensional on the device: groups = prod(blocks) CTAs of prod(wgs) threads
// (reference src/pocl/backend.jl:133-143); the N-d indices are recovered per work-item by `expand`
// (reference src/nditeration.jl:74-82,115-126).  Ids are 1-based like KI.get_*_id (src/intrinsics.jl:19-103).
#pragma once
#include <cstdint>

namespace b200 {

constexpr int KA_MAX_NDIM = 4;

struct KaCtx {
    int ndim;
    int dynamic_check;
    long long ndrange[KA_MAX_NDIM];
    long long wgs[KA_MAX_NDIM];
    long long blocks[KA_MAX_NDIM];
};

#ifdef __CUDACC__
// @index(Group, Linear) / @index(Local, Linear) / @index(Global, Linear)   (src/KernelAbstractions.jl:491-501)
__device__ __forceinline__ long long ka_group_linear() { return (long long)blockIdx.x + 1; }
__device__ __forceinline__ long long ka_local_linear() { return (long long)threadIdx.x + 1; }
__device__ __forceinline__ long long ka_global_linear() {
    return (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
}

// CartesianIndices(dims)[lin], column-major, 1-based
__device__ __forceinline__ void ka_unravel(long long lin, int n, const long long *dims, long long *out) {
    long long r = lin - 1;
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d) {
        if (d < n) {
            const long long s = dims[d];
            const long long q = r / s;
            out[d] = r - q * s + 1;
            r = q;
        }
    }
}

// @index(Group|Local|Global, Cartesian)   (src/KernelAbstractions.jl:503-511)
__device__ __forceinline__ void ka_group_cartesian(const KaCtx &c, long long *I) {
    ka_unravel(ka_group_linear(), c.ndim, c.blocks, I);
}
__device__ __forceinline__ void ka_local_cartesian(const KaCtx &c, long long *I) {
    ka_unravel(ka_local_linear(), c.ndim, c.wgs, I);
}
__device__ __forceinline__ void ka_expand(const KaCtx &c, long long *I) {
    long long g[KA_MAX_NDIM], l[KA_MAX_NDIM];
    ka_group_cartesian(c, g);
    ka_local_cartesian(c, l);
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d)
        if (d < c.ndim) I[d] = (g[d] - 1) * c.wgs[d] + l[d];
}

// __validindex   (src/pocl/backend.jl:207-214)
__device__ __forceinline__ bool ka_valid(const KaCtx &c) {
    if (!c.dynamic_check) return true;
    long long I[KA_MAX_NDIM];
    ka_expand(c, I);
    bool ok = true;
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d)
        if (d < c.ndim) ok = ok && (I[d] >= 1) && (I[d] <= c.ndrange[d]);
    return ok;
}

// prod(@groupsize())
__device__ __forceinline__ long long ka_groupsize_prod(const KaCtx &c) {
    long long p = 1;
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d)
        if (d < c.ndim) p *= c.wgs[d];
    return p;
}
#endif

}  // namespace b200
