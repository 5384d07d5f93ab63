// ka_device.cuh -- device-side restatement of KA's iteration space, written once for every kernel twin and plug-in.
//
// KaCtx is the by-value image of CompilerMetadata{ndrange, iterspace} (reference src/compiler.jl:1-33).  The reference
// lowers a KA launch to a ONE-dimensional device launch of prod(blocks) groups x prod(wgs) items and recovers the N-d
// indices per work-item with `expand` (src/pocl/backend.jl:133-143, src/nditeration.jl:74-82,115-126) -- one integer
// division and one modulo per dimension per work-item, in Int64 (":116 this causes an exception branch and a div").
// On a B200 a 64-bit division is a ~40-instruction subroutine; at 2^26 work-items that alone costs more than the copy
// it indexes.  This header keeps the reference's index SEMANTICS (Appendix B of SURVEY.md: column-major unravel of the
// linear group / item ids, I_d = (g_d - 1) * wgs_d + l_d, __validindex = I in ndrange; 1-based ids) and strength-reduces
// the arithmetic, in the host's choice of three lowerings (ka_ctx_finish):
//
//   KA_MODE_GRID   the launch is issued as a native <=3-d grid of <=3-d blocks: group and item Cartesian ids ARE
//                  blockIdx / threadIdx -- no division at all.  CUDA linearises both x-fastest, i.e. column-major, so warps,
//                  group order and every linear id are exactly those of the 1-d launch.  Taken when ndim <= 3 and the
//                  grid limits allow (blocks_y, blocks_z <= 65535, wgs_z <= 64).
//   KA_MODE_MAGIC  1-d launch; every divisor (blocks_d, wgs_d) comes with a host-precomputed multiplier and shift
//                  (q = umulhi(n, mul) >> shr, exact for n < 2^31): one 32-bit multiply-high per dimension.
//   KA_MODE_DIV64  the reference's arithmetic verbatim (kept for A/B measurements: `registry.ka_mode = 2`).
//
// __validindex is strength-reduced the same way: the host marks the dimensions whose ndrange is not a multiple of the
// workgroup size (`ragged`); a launch without ragged dimensions has no invalid work-item and the test folds to `true`
// without expanding anything, and otherwise only the ragged dimensions are compared.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define KA_HD __host__ __device__
#else
#define KA_HD
#endif

namespace b200 {

constexpr int KA_MAX_NDIM = 4;
enum { KA_MODE_GRID = 0, KA_MODE_MAGIC = 1, KA_MODE_DIV64 = 2 };

struct KaCtx {
    int ndim;
    int dynamic_check;
    long long ndrange[KA_MAX_NDIM];
    long long wgs[KA_MAX_NDIM];
    long long blocks[KA_MAX_NDIM];
    // ---- filled by ka_ctx_finish ----
    int mode;                        // KA_MODE_*
    int ragged;                      // bit d: blocks[d] * wgs[d] != ndrange[d] (only those dimensions can hold invalid work-items)
    unsigned items;                  // prod(wgs)
    unsigned pad_;
    unsigned bmul[KA_MAX_NDIM], bshr[KA_MAX_NDIM];   // n / blocks[d]
    unsigned wmul[KA_MAX_NDIM], wshr[KA_MAX_NDIM];   // n / wgs[d]
};

// ---- host: magic numbers, lowering choice, launch geometry ---------------------------------------------------
// q = n / d for 0 <= n < 2^31, d >= 1:  d == 1 -> (mul, shr) = (0, 0) and the device takes n itself; otherwise
// p = 31 + ceil(log2 d), mul = ceil(2^p / d) (fits 32 bits), shr = p - 32, q = umulhi(n, mul) >> shr.
static inline void ka_magic(unsigned long long d, unsigned *mul, unsigned *shr) {
    if (d <= 1) {
        *mul = 0;
        *shr = 0;
        return;
    }
    unsigned lg = 0;
    while ((1ull << lg) < d) ++lg;
    const unsigned p = 31 + lg;
    *mul = (unsigned)(((1ull << p) + d - 1) / d);
    *shr = p - 32;
}

// Completes a KaCtx whose ndim / dynamic_check / ndrange / wgs / blocks are set.  force_mode < 0: choose.
static inline void ka_ctx_finish(KaCtx *c, int force_mode = -1) {
    unsigned long long items = 1, groups = 1;
    c->ragged = 0;
    for (int d = 0; d < KA_MAX_NDIM; ++d) {
        const bool in = d < c->ndim;
        if (!in) {
            c->ndrange[d] = 1;
            c->wgs[d] = 1;
            c->blocks[d] = 1;
        }
        items *= (unsigned long long)c->wgs[d];
        groups *= (unsigned long long)c->blocks[d];
        ka_magic((unsigned long long)c->blocks[d], &c->bmul[d], &c->bshr[d]);
        ka_magic((unsigned long long)c->wgs[d], &c->wmul[d], &c->wshr[d]);
        if (in && c->dynamic_check && c->blocks[d] * c->wgs[d] != c->ndrange[d]) c->ragged |= 1 << d;   // blocks = cld(ndrange, wgs)
    }
    c->items = (unsigned)items;
    c->pad_ = 0;
    const bool grid_ok = c->ndim <= 3 && c->blocks[1] <= 65535 && c->blocks[2] <= 65535 && c->wgs[2] <= 64 &&
                         c->blocks[0] <= 2147483647LL;
    const bool magic_ok = groups < (1ull << 31) && items <= 1024;
    int mode = grid_ok ? KA_MODE_GRID : (magic_ok ? KA_MODE_MAGIC : KA_MODE_DIV64);
    if (force_mode == KA_MODE_DIV64) mode = KA_MODE_DIV64;
    if (force_mode == KA_MODE_MAGIC && magic_ok) mode = KA_MODE_MAGIC;
    if (force_mode == KA_MODE_GRID && grid_ok) mode = KA_MODE_GRID;
    c->mode = mode;
}

#ifdef __CUDACC__
// launch geometry of a finished context: a native grid in KA_MODE_GRID, the reference's 1-d launch otherwise
static inline void ka_launch_dims(const KaCtx &c, dim3 *grid, dim3 *block) {
    if (c.mode == KA_MODE_GRID) {
        *grid = dim3((unsigned)c.blocks[0], (unsigned)c.blocks[1], (unsigned)c.blocks[2]);
        *block = dim3((unsigned)c.wgs[0], (unsigned)c.wgs[1], (unsigned)c.wgs[2]);
    } else {
        unsigned long long groups = 1;
        for (int d = 0; d < KA_MAX_NDIM; ++d) groups *= (unsigned long long)c.blocks[d];
        *grid = dim3((unsigned)groups);
        *block = dim3(c.items);
    }
}

// ---- device ---------------------------------------------------------------------------------------------------
// 0-based linear ids of the group and of the item inside its group -- the same numbers in every lowering
__device__ __forceinline__ unsigned ka_group0(const KaCtx &c) {
    return c.mode == KA_MODE_GRID ? blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) : blockIdx.x;
}
__device__ __forceinline__ unsigned ka_local0(const KaCtx &c) {
    return c.mode == KA_MODE_GRID ? threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z) : threadIdx.x;
}
// @index(Group, Linear) / @index(Local, Linear) / @index(Global, Linear)   (src/KernelAbstractions.jl:491-501)
__device__ __forceinline__ long long ka_group_linear(const KaCtx &c) { return (long long)ka_group0(c) + 1; }
__device__ __forceinline__ long long ka_local_linear(const KaCtx &c) { return (long long)ka_local0(c) + 1; }
__device__ __forceinline__ long long ka_global_linear(const KaCtx &c) {
    return (long long)((unsigned long long)ka_group0(c) * c.items + ka_local0(c)) + 1;
}

__device__ __forceinline__ unsigned ka_fastdiv(unsigned n, unsigned mul, unsigned shr) {
    return mul ? (__umulhi(n, mul) >> shr) : n;   // mul == 0 <=> divisor 1
}

// 0-based Cartesian ids of group and item (column-major unravel of the linear ids)
__device__ __forceinline__ void ka_group_cart0(const KaCtx &c, unsigned *g) {
    if (c.mode == KA_MODE_GRID) {
        g[0] = blockIdx.x;
        g[1] = blockIdx.y;
        g[2] = blockIdx.z;
        g[3] = 0;
    } else if (c.mode == KA_MODE_MAGIC) {
        unsigned r = blockIdx.x;
#pragma unroll
        for (int d = 0; d < KA_MAX_NDIM; ++d) {
            const unsigned q = ka_fastdiv(r, c.bmul[d], c.bshr[d]);
            g[d] = r - q * (unsigned)c.blocks[d];
            r = q;
        }
    } else {
        long long r = (long long)blockIdx.x;
#pragma unroll
        for (int d = 0; d < KA_MAX_NDIM; ++d) {
            const long long q = r / c.blocks[d];
            g[d] = (unsigned)(r - q * c.blocks[d]);
            r = q;
        }
    }
}
__device__ __forceinline__ void ka_local_cart0(const KaCtx &c, unsigned *l) {
    if (c.mode == KA_MODE_GRID) {
        l[0] = threadIdx.x;
        l[1] = threadIdx.y;
        l[2] = threadIdx.z;
        l[3] = 0;
    } else if (c.mode == KA_MODE_MAGIC) {
        unsigned r = threadIdx.x;
#pragma unroll
        for (int d = 0; d < KA_MAX_NDIM; ++d) {
            const unsigned q = ka_fastdiv(r, c.wmul[d], c.wshr[d]);
            l[d] = r - q * (unsigned)c.wgs[d];
            r = q;
        }
    } else {
        long long r = (long long)threadIdx.x;
#pragma unroll
        for (int d = 0; d < KA_MAX_NDIM; ++d) {
            const long long q = r / c.wgs[d];
            l[d] = (unsigned)(r - q * c.wgs[d]);
            r = q;
        }
    }
}

// @index(Group|Local|Global, Cartesian), 1-based   (src/KernelAbstractions.jl:503-511)
__device__ __forceinline__ void ka_group_cartesian(const KaCtx &c, long long *I) {
    unsigned g[KA_MAX_NDIM];
    ka_group_cart0(c, g);
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d) I[d] = (long long)g[d] + 1;
}
__device__ __forceinline__ void ka_local_cartesian(const KaCtx &c, long long *I) {
    unsigned l[KA_MAX_NDIM];
    ka_local_cart0(c, l);
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d) I[d] = (long long)l[d] + 1;
}
// expand (src/nditeration.jl:74-82): I_d = (g_d - 1) * wgs_d + l_d
__device__ __forceinline__ void ka_expand(const KaCtx &c, long long *I) {
    unsigned g[KA_MAX_NDIM], l[KA_MAX_NDIM];
    ka_group_cart0(c, g);
    ka_local_cart0(c, l);
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d) I[d] = (long long)((unsigned long long)g[d] * (unsigned)c.wgs[d] + l[d]) + 1;
}

// __validindex   (src/pocl/backend.jl:207-214): expand(...) in ndrange.  Only ragged dimensions can fail.
__device__ __forceinline__ bool ka_valid(const KaCtx &c) {
    if (c.ragged == 0) return true;
    unsigned g[KA_MAX_NDIM], l[KA_MAX_NDIM];
    ka_group_cart0(c, g);
    ka_local_cart0(c, l);
    bool ok = true;
#pragma unroll
    for (int d = 0; d < KA_MAX_NDIM; ++d)
        if ((c.ragged >> d) & 1)
            ok = ok && ((long long)((unsigned long long)g[d] * (unsigned)c.wgs[d] + l[d]) < c.ndrange[d]);
    return ok;
}

// prod(@groupsize())
__device__ __forceinline__ long long ka_groupsize_prod(const KaCtx &c) { return (long long)c.items; }
#endif

}  // namespace b200
