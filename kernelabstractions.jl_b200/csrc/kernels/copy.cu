// copy.cu -- b200_copy / b200_fill: the device side of KA.copyto!(d<-d), copy_kernel(!), init_kernel.
//
// Reference semantics: A[I] = B[I] for I in 1:length(A)  (src/KernelAbstractions.jl:874-877,
// examples/memcopy.jl:4-7, examples/memcopy_static.jl:4-7) and arr[I] = f(T) (:869-872).
// The copy is dtype agnostic (bytes), bit-exact, and HBM bound: 2 bytes of traffic per byte copied.
//
// Two implementations, selected by the `copy.impl` tuning knob (default chosen from measurements,
// see DESIGN.md / profiles/):
//   impl 0  "ldg":  persistent grid of CTAs, each thread keeps UNROLL independent 128-bit
//                   ld.global.nc.L1::no_allocate in flight, then UNROLL st.global (streaming).
//   impl 1  "bulk": one elected thread per CTA drives a ring of cp.async.bulk (TMA, 1-D)
//                   global->shared loads completing on mbarriers and shared->global bulk stores
//                   (SASS: UBLKCP), no register staging at all.
#include "../common.cuh"
#include "../tune.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// impl 0: vectorised LDG/STG
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// n16 = number of 16-byte units.  Each CTA walks tiles of THREADS*UNROLL units, grid-strided.
template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) copy_vec16_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src,
                                                             size_t n16) {
    const size_t tile = (size_t)THREADS * UNROLL;
    const size_t ntiles = n16 / tile;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ld_stream(src + base + (size_t)u * THREADS);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st_stream(dst + base + (size_t)u * THREADS, v[u]);
    }
    // ragged remainder (< one tile): handled by the last CTA, one unit per thread per step
    if (blockIdx.x == gridDim.x - 1) {
        for (size_t i = ntiles * tile + threadIdx.x; i < n16; i += THREADS) dst[i] = src[i];
    }
}

// Generic element-width copy for pointers that are not mutually 16-byte aligned, and for byte tails.
template <typename T>
__global__ void copy_elem_kernel(T *__restrict__ dst, const T *__restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// impl 1: TMA 1-D bulk copies through shared memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// bytes16: total bytes, a multiple of 16; dst/src 16-byte aligned.  CTA c owns chunks c, c+grid, ...
template <int STAGES>
__global__ void __launch_bounds__(32) copy_bulk_kernel(unsigned char *__restrict__ dst,
                                                       const unsigned char *__restrict__ src, size_t bytes16,
                                                       uint32_t stage_bytes, int prefetch) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[STAGES];
    if (threadIdx.x != 0) return;
    // Programmatic dependent launch: let the next kernel of the stream be scheduled now (its CTAs fit next to ours), and do our own
    // set-up before waiting for the previous kernel's memory -- back-to-back copies then hide the launch latency (~2 us of 85).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const size_t nchunks = (bytes16 + stage_bytes - 1) / stage_bytes;
    const size_t mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto chunk_off = [&](size_t i) { return (blockIdx.x + i * (size_t)gridDim.x) * (size_t)stage_bytes; };
    auto chunk_len = [&](size_t i) {
        size_t off = chunk_off(i);
        size_t rem = bytes16 - off;
        return (uint32_t)(rem < stage_bytes ? rem : stage_bytes);
    };
    // While the previous kernel of the stream drains, pull this CTA's first ring of chunks into L2.  A prefetch has no architectural
    // effect -- L2 is the point of coherence, whatever the previous kernel still writes lands there too -- so it may run BEFORE
    // the dependency wait; the ring's first loads then hit L2 instead of paying the DRAM ramp (~1.5 us of a 12 us 32 MiB copy).
    if (prefetch)
        for (size_t i = 0; i < mine && i < (size_t)STAGES; ++i)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + chunk_off(i)), "r"(chunk_len(i)) : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");   // everything before us in the stream is complete and visible

    // prologue: fill the ring
    for (size_t i = 0; i < mine && i < (size_t)STAGES; ++i) {
        uint32_t len = chunk_len(i);
        mbar_expect_tx(&full[i], len);
        bulk_g2s(smem + i * stage_bytes, src + chunk_off(i), len, &full[i]);
    }
    for (size_t i = 0; i < mine; ++i) {
        const int s = (int)(i % STAGES);
        const uint32_t parity = (uint32_t)((i / STAGES) & 1);
        mbar_wait(&full[s], parity);
        bulk_s2g(dst + chunk_off(i), smem + (size_t)s * stage_bytes, chunk_len(i));
        bulk_commit();
        // refill the stage used by chunk i-1: its store (the previous group) must have finished
        // READING shared memory; allow the store just issued to stay in flight.
        if (i >= 1 && i - 1 + STAGES < mine) {
            bulk_wait_read<1>();
            const size_t j = i - 1 + STAGES;
            const int sj = (int)(j % STAGES);
            uint32_t len = chunk_len(j);
            mbar_expect_tx(&full[sj], len);
            bulk_g2s(smem + (size_t)sj * stage_bytes, src + chunk_off(j), len, &full[sj]);
        }
    }
    bulk_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------------------------
template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS) fill_vec16_kernel(uint4 *__restrict__ dst, size_t n16, uint4 pattern) {
    const size_t tile = (size_t)THREADS * UNROLL;
    const size_t ntiles = n16 / tile;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t base = t * tile + threadIdx.x;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st_stream(dst + base + (size_t)u * THREADS, pattern);
    }
    if (blockIdx.x == gridDim.x - 1)
        for (size_t i = ntiles * tile + threadIdx.x; i < n16; i += THREADS) dst[i] = pattern;
}

// TMA form: every CTA (one warp) writes the 16-byte pattern into a shared-memory chunk ONCE and then only issues
// cp.async.bulk shared -> global stores of that chunk: no LSU traffic, no registers, and -- the source never changes -- no
// wait between stores, so the store queue stays as deep as the hardware allows.  n16 = number of 16-byte words.
__global__ void __launch_bounds__(32) fill_bulk_kernel(unsigned char *__restrict__ dst, size_t n16, uint4 pattern, uint32_t chunk_bytes,
                                                       uint32_t span_chunks) {
    extern __shared__ __align__(128) unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (uint32_t i = threadIdx.x; i < chunk_bytes / 16; i += 32) reinterpret_cast<uint4 *>(smem)[i] = pattern;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes above -> visible to the bulk-copy engine
    __syncwarp();
    if (threadIdx.x != 0) return;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const size_t bytes = n16 * 16;
    if (span_chunks == 0) {   // chunks interleaved over the CTAs
        for (size_t off = (size_t)blockIdx.x * chunk_bytes; off < bytes; off += (size_t)gridDim.x * chunk_bytes) {
            const size_t rem = bytes - off;
            bulk_s2g(dst + off, smem, (uint32_t)(rem < chunk_bytes ? rem : chunk_bytes));
            bulk_commit();
        }
    } else {                  // one contiguous span of chunks per CTA
        size_t off = (size_t)blockIdx.x * span_chunks * chunk_bytes;
        const size_t end = off + (size_t)span_chunks * chunk_bytes < bytes ? off + (size_t)span_chunks * chunk_bytes : bytes;
        for (; off < end; off += chunk_bytes) {
            const size_t rem = end - off;
            bulk_s2g(dst + off, smem, (uint32_t)(rem < chunk_bytes ? rem : chunk_bytes));
            bulk_commit();
        }
    }
    bulk_wait_all<0>();
}

__global__ void fill_bytes_kernel(unsigned char *__restrict__ dst, size_t nbytes, uint4 pattern, int phase) {
    // byte-granular fill for unaligned heads/tails; `phase` = offset of dst[0] inside the 16-byte pattern
    const unsigned char *pat = reinterpret_cast<const unsigned char *>(&pattern);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = pat[(i + phase) & 15];
}

static int launch_copy_generic(unsigned char *d, const unsigned char *s, size_t bytes, cudaStream_t st) {
    // largest power-of-two width that divides both addresses and the byte count
    uintptr_t bits = (uintptr_t)d | (uintptr_t)s | (uintptr_t)bytes;
    int grid = sm_count() * 8;
    if ((bits & 7) == 0) {
        copy_elem_kernel<uint64_t><<<grid, 256, 0, st>>>((uint64_t *)d, (const uint64_t *)s, bytes / 8);
    } else if ((bits & 3) == 0) {
        copy_elem_kernel<uint32_t><<<grid, 256, 0, st>>>((uint32_t *)d, (const uint32_t *)s, bytes / 4);
    } else if ((bits & 1) == 0) {
        copy_elem_kernel<uint16_t><<<grid, 256, 0, st>>>((uint16_t *)d, (const uint16_t *)s, bytes / 2);
    } else {
        copy_elem_kernel<unsigned char><<<grid, 256, 0, st>>>(d, s, bytes);
    }
    B200_CHECK_LAUNCH("copy_elem_kernel");
    return B200_OK;
}

template <int UNROLL>
static int launch_vec16(unsigned char *d, const unsigned char *s, size_t n16, int ctas_per_sm, cudaStream_t st) {
    constexpr int THREADS = 256;
    size_t ntiles = n16 / ((size_t)THREADS * UNROLL);
    size_t grid = (size_t)sm_count() * ctas_per_sm;
    if (grid > ntiles) grid = ntiles > 0 ? ntiles : 1;
    copy_vec16_kernel<THREADS, UNROLL><<<(unsigned)grid, THREADS, 0, st>>>((uint4 *)d, (const uint4 *)s, n16);
    B200_CHECK_LAUNCH("copy_vec16_kernel");
    return B200_OK;
}

static int launch_bulk(unsigned char *d, const unsigned char *s, size_t bytes16, cudaStream_t st) {
    int stage_kb = tune_get("copy.bulk_stage_kb", 16);
    // ring depth: 4 stages at large sizes (r01 sweeps); short copies (<= 64 MiB, e.g. the 8-GPU slab of the 2^26 copy) take 8 --
    // the ring is also what a CTA prefetches into L2 before its dependency wait, and 8 x 16 KiB per CTA covers the DRAM ramp of
    // a 12 us launch: 12.4 -> 10.3 us for 32 MiB (profiles/r05_slab_probe.txt).  `copy.bulk_stages` pins a value.
    int stages = tune_get("copy.bulk_stages", 0);
    if (stages == 0) stages = bytes16 <= ((size_t)64 << 20) ? 8 : 4;
    int ctas_per_sm = tune_get("copy.bulk_ctas_per_sm", 1);
    const int prefetch = tune_get("copy.l2_prefetch", 1);
    uint32_t stage_bytes = (uint32_t)stage_kb * 1024u;
    size_t smem = (size_t)stage_bytes * stages;
    size_t nchunks = (bytes16 + stage_bytes - 1) / stage_bytes;
    size_t grid = (size_t)sm_count() * ctas_per_sm;
    if (grid > nchunks) grid = nchunks;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tune_get("copy.pdl", 1) != 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define B200_BULK_CASE(S)                                                                                       \
    case S:                                                                                                     \
        B200_CUDA(cudaFuncSetAttribute(copy_bulk_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                       (int)smem));                                                             \
        B200_CUDA(cudaLaunchKernelEx(&cfg, copy_bulk_kernel<S>, d, (const unsigned char *)s, bytes16, stage_bytes, prefetch)); \
        break;
    switch (stages) {
        B200_BULK_CASE(2)
        B200_BULK_CASE(3)
        B200_BULK_CASE(4)
        B200_BULK_CASE(5)
        B200_BULK_CASE(6)
        B200_BULK_CASE(8)
        default:
            return set_error(B200_ERR_INVALID_VALUE, "copy.bulk_stages must be 2,3,4,5,6 or 8");
    }
#undef B200_BULK_CASE
    B200_CHECK_LAUNCH("copy_bulk_kernel");
    return B200_OK;
}

int copy_bytes(void *dst, const void *src, size_t bytes, cudaStream_t st) {
    unsigned char *d = (unsigned char *)dst;
    const unsigned char *s = (const unsigned char *)src;
    if ((((uintptr_t)d | (uintptr_t)s) & 15) != 0) {
        // same misalignment on both sides: peel a head so the body is 16-byte aligned
        if ((((uintptr_t)d ^ (uintptr_t)s) & 15) == 0 && bytes >= 64) {
            size_t head = 16 - ((uintptr_t)d & 15);
            B200_TRY(launch_copy_generic(d, s, head, st));
            d += head;
            s += head;
            bytes -= head;
        } else {
            return launch_copy_generic(d, s, bytes, st);
        }
    }
    size_t body = bytes & ~(size_t)15;
    if (body) {
        int impl = tune_get("copy.impl", 1);  // measured: TMA bulk ring (16 KiB x 4) >= LDG/STG at every size >= 1 MiB
        if (impl == 1 && body >= (size_t)1 << 20) {
            B200_TRY(launch_bulk(d, s, body, st));
        } else {
            int unroll = tune_get("copy.unroll", 8);
            int cps = tune_get("copy.ctas_per_sm", 32);
            size_t n16 = body / 16;
            switch (unroll) {
                case 1: B200_TRY(launch_vec16<1>(d, s, n16, cps, st)); break;
                case 2: B200_TRY(launch_vec16<2>(d, s, n16, cps, st)); break;
                case 8: B200_TRY(launch_vec16<8>(d, s, n16, cps, st)); break;
                default: B200_TRY(launch_vec16<4>(d, s, n16, cps, st)); break;
            }
        }
    }
    if (bytes > body) B200_TRY(launch_copy_generic(d + body, s + body, bytes - body, st));
    return B200_OK;
}

int fill_elems(void *dst, size_t count, int elsize, const void *value, cudaStream_t st) {
    if (count == 0) return B200_OK;
    if (elsize != 1 && elsize != 2 && elsize != 4 && elsize != 8 && elsize != 16)
        return set_error(B200_ERR_UNSUPPORTED, "fill: element size %d not in {1,2,4,8,16}", elsize);
    uint4 pattern;
    unsigned char *pb = reinterpret_cast<unsigned char *>(&pattern);
    for (int i = 0; i < 16; ++i) pb[i] = ((const unsigned char *)value)[i % elsize];
    unsigned char *d = (unsigned char *)dst;
    size_t bytes = count * (size_t)elsize;
    // dst is at least elsize aligned, so the pattern phase at any 16-byte boundary is 0
    size_t head = ((uintptr_t)d & 15) ? 16 - ((uintptr_t)d & 15) : 0;
    if (head > bytes) head = bytes;
    if (head) {
        // pattern phase such that byte 0 of dst gets byte 0 of the element: phase = 0 and the aligned body
        // starts at a multiple of elsize because head is a multiple of elsize (16 and the address both are)
        fill_bytes_kernel<<<1, 32, 0, st>>>(d, head, pattern, 0);
        B200_CHECK_LAUNCH("fill_bytes_kernel");
    }
    size_t body = (bytes - head) & ~(size_t)15;
    if (body) {
        constexpr int THREADS = 256, UNROLL = 4;
        size_t n16 = body / 16;
        // >= 1 MiB: the bulk-store form.  32 KiB chunks interleaved over 32 CTAs per SM (more CTAs than fit at once, so the block
        // scheduler balances the tail and decorrelates the CTAs' address streams: 148 x 4 CTAs in lockstep measured 5.1-5.6 TB/s):
        // 7.2 TB/s at 2^26 and 2^28 Float32 against 6.0 for the 128-bit-store kernel (tools/fill_probe.py).
        if (tune_get("fill.impl", 1) == 1 && body >= ((size_t)1 << 20)) {
            const uint32_t chunk = (uint32_t)tune_get("fill.chunk_kb", 32) << 10;
            size_t grid = (size_t)sm_count() * tune_get("fill.ctas_per_sm", 32);
            if (grid > cdiv(body, (size_t)chunk)) grid = cdiv(body, (size_t)chunk);
            B200_CUDA(cudaFuncSetAttribute(fill_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid);
            cfg.blockDim = dim3(32);
            cfg.dynamicSmemBytes = chunk;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = tune_get("copy.pdl", 1) != 0;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            uint32_t span = 0;
            if (tune_get("fill.span", 0)) {
                span = (uint32_t)cdiv(cdiv(body, (size_t)chunk), grid);
                cfg.gridDim = dim3((unsigned)cdiv(cdiv(body, (size_t)chunk), span));
            }
            B200_CUDA(cudaLaunchKernelEx(&cfg, fill_bulk_kernel, d + head, n16, pattern, chunk, span));
            B200_CHECK_LAUNCH("fill_bulk_kernel");
        } else {
        size_t ntiles = n16 / ((size_t)THREADS * UNROLL);
        size_t grid = (size_t)sm_count() * 8;
        if (grid > ntiles) grid = ntiles > 0 ? ntiles : 1;
        fill_vec16_kernel<THREADS, UNROLL><<<(unsigned)grid, THREADS, 0, st>>>((uint4 *)(d + head), n16, pattern);
        B200_CHECK_LAUNCH("fill_vec16_kernel");
        }
    }
    size_t tail = bytes - head - body;
    if (tail) {
        fill_bytes_kernel<<<1, 32, 0, st>>>(d + head + body, tail, pattern, 0);
        B200_CHECK_LAUNCH("fill_bytes_kernel");
    }
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_copy(void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return B200_OK;
    if (!dst || !src) return set_error(B200_ERR_INVALID_VALUE, "b200_copy: NULL pointer");
    const char *d = (const char *)dst, *s = (const char *)src;
    if (d < s + bytes && s < d + bytes)  // Base.mightalias (reference src/pocl/backend.jl:45-47)
        return set_error(B200_ERR_ALIAS, "Arrays may not alias");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return copy_bytes(dst, src, bytes, st);
}

extern "C" b200_status b200_fill(void *dst, size_t count, int32_t elsize, const void *value) {
    if (count == 0) return B200_OK;
    if (!dst || !value) return set_error(B200_ERR_INVALID_VALUE, "b200_fill: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return fill_elems(dst, count, elsize, value, st);
}
