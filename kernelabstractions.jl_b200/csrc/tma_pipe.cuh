// tma_pipe.cuh -- mbarrier + TMA (cp.async.bulk.tensor) primitives shared by the TMA-staged kernels, and the host-side
// tensor-map encoder (driver entry point resolved through the runtime, so libcuda is not a link dependency).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b200 {
namespace pipe {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int get_encode_fn(EncodeTiledFn *fn) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess)
            return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        cached = (EncodeTiledFn)p;
    }
    *fn = cached;
    return B200_OK;
}

// 2-D tensor map over a column-major matrix (dim0 contiguous, dim1 strided by ld elements); out-of-bounds box
// elements are filled with zeros, which is exactly the reference's zero-padded edge tile.
static inline int make_tmap_2d(CUtensorMap *tm, CUtensorMapDataType dt, int elem_bytes, const void *ptr, int64_t dim0,
                               int64_t dim1, int64_t ld, int box0, int box1, CUtensorMapSwizzle swz,
                               CUtensorMapL2promotion l2) {
    EncodeTiledFn enc = nullptr;
    B200_TRY(get_encode_fn(&enc));
    cuuint64_t gdim[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, const_cast<void *>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return B200_OK;
}

}  // namespace pipe
}  // namespace b200
