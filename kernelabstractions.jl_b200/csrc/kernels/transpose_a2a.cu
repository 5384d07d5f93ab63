// transpose_a2a.cu -- distributed transpose: the tile transpose and the all-to-all it implies, in ONE data kernel over NVLink
// peer memory (no NCCL, no staging buffer, no second pass over HBM).
//
// Setting ("all-to-all only if input were column-sharded"): a global N0 x N1 column-major matrix b is sharded
// by COLUMNS over G ranks (rank g holds b_g = b[:, c1(g)], N0 x N1/G, ld = N0) and so is the result a = b^T (rank h holds
// a_h = a[:, c0(h)], N1 x N0/G, ld = N1).  Element-wise  a_h[c1(g).start + i', j'] = b_g[c0(h).start + j', i'] : rank g owns,
// for every destination h, the (N0/G) x (N1/G) block b_g[c0(h), :] and must deliver its transpose into rows c1(g) of rank h's
// slab.  The reference has no multi-device code (examples/mpi.jl:24-39 is a host-staged ring); the NCCL formulation is an
// all-to-all of the blocks followed by G local transposes, i.e. every byte crosses HBM three times on the receiver.
//
// Here every rank runs the swizzled 64 x 64 tile transpose of transpose.cu once over its whole slab, destinations interleaved over
// consecutive CTAs: the 128-bit stores of a tile go straight into the destination's slab through its cudaIpc mapping (NVSwitch gives every
// pair full bandwidth), so the exchange IS the store phase of the transpose.  Completion is signalled by a second, tiny kernel
// (b200_sym_barrier): a system-scope fence, one flag store per peer, a spin on the own flags -- ordered behind the data kernel
// on the stream, so a peer that sees the flag sees the data.
//
// Bound: NVLink egress for G > 1 -- (G-1)/G of the slab leaves the GPU -- and HBM for the local part.
#include <cstring>

#include "../common.cuh"
#include "../tma_pipe.cuh"
#include "../tune.h"

namespace b200 {

constexpr int kSymMaxRanks = 8;
constexpr size_t kSymFlagBytes = 256;   // words [0..7]: arrival flags of b200_sym_barrier, [8..15]: "slab may be overwritten" flags
                                        // of b200_transpose_alltoall -- written by the peers, one 8-byte word per source rank;
                                        // words [16..]: LOCAL 32-bit counters of the fused completion (done per destination, all)

struct Sym {
    void *local = nullptr;      // user bytes, then kSymFlagBytes of flags
    size_t bytes = 0;
    int world = 0, rank = -1;
    void *peer[kSymMaxRanks] = {};
    unsigned long long epoch = 0;
    unsigned long long a2a_calls = 0;   // b200_transpose_alltoall calls so far (the value of the "ready" flags)
    bool ipc = true;                    // peers mapped with cudaIpcOpenMemHandle (b200_sym_connect) or plain pointers (..._ptrs)
    uint32_t *error_host = nullptr;     // pinned, mapped: set by a kernel whose wait for a peer timed out
    uint32_t *error_dev = nullptr;
};

// Every wait for a peer inside a kernel is BOUNDED (`a2a.timeout_ms`, default 10 s): a rank that never makes the matching call
// (failed validation, crashed process) must turn into an error on its peers' next call, not into a GPU that spins for ever.
struct A2ACtl {
    unsigned long long timeout_ns;      // 0 = wait for ever
    uint32_t *error;                    // host-visible flag
};
__device__ __forceinline__ unsigned long long a2a_now_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
// spin until *p >= want; false (and the error flag raised) when the budget runs out
__device__ __forceinline__ bool a2a_spin_until(const unsigned long long *p, unsigned long long want, const A2ACtl &ctl) {
    unsigned long long v;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= want) return true;
        if (ctl.timeout_ns && (++spins & 1023u) == 0) {
            const unsigned long long now = a2a_now_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > ctl.timeout_ns) {
                *reinterpret_cast<volatile uint32_t *>(ctl.error) = 1u;
                __threadfence_system();
                return false;
            }
        }
    }
}

struct A2APeers {
    void *dst[kSymMaxRanks];
};

template <typename T>
struct alignas(16) V16 {
    T e[16 / sizeof(T)];
};
template <typename T>
__device__ __forceinline__ V16<T> ldg_v(const T *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return *reinterpret_cast<V16<T> *>(&r);
}

// Write-after-read guard of the distributed transpose.  Rank h's slab may still be read by the kernels h enqueued after the
// PREVIOUS call; h's stream reaches its own call k only when they are done, and the first thing that call does is to raise
// "ready = k" in every peer's flag area (words 8 + h).  A CTA that is about to store into h's slab waits for that word -- in
// steady state it was raised long ago and the wait is one L2 read hidden behind the tile's loads.  No deadlock: a rank raises
// its flags without waiting for anybody.
__device__ __forceinline__ void a2a_post_ready(const A2APeers &flags, int world, int rank, unsigned long long call) {
    if (blockIdx.x < (unsigned)world && threadIdx.x == 0) {
        unsigned long long *theirs = reinterpret_cast<unsigned long long *>(flags.dst[blockIdx.x]) + 8 + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(call) : "memory");
    }
}
__device__ __forceinline__ void a2a_wait_ready(const A2APeers &flags, int rank, int h, unsigned long long call, const A2ACtl &ctl) {
    if (threadIdx.x == 0) a2a_spin_until(reinterpret_cast<const unsigned long long *>(flags.dst[rank]) + 8 + h, call, ctl);
}

// Rank g's data kernel.  src block for destination h: b_local + h * rows_blk (rows_blk = N0/G rows of every local column);
// its transpose (cols_blk x rows_blk) goes to peers.dst[h] + row_off (row_off = g * cols_blk) with leading dimension ld_dst = N1.
// Tile scheme and swizzle as transpose64_any_kernel (transpose.cu); rows_blk, cols_blk multiples of 64.
template <typename T>
__global__ void __launch_bounds__(256) transpose_a2a_kernel(A2APeers peers, const T *__restrict__ b_local, int64_t ld_src, int64_t ld_dst,
                                                            int64_t rows_blk, int64_t row_off, int tiles_i, int tiles_j, int first_dst, int order, int world,
                                                            A2APeers flags, unsigned long long call, A2ACtl ctl, unsigned long long epoch) {
    constexpr int V = 16 / sizeof(T), CH = 64 / V, PASSES = CH * CH / 256, ROWS_PER_PASS = 256 / CH;
    __shared__ V16<T> tile[64 * CH];
    __shared__ unsigned s_last;
    const int t = threadIdx.x;
    a2a_post_ready(flags, world, first_dst, call);
    // Destinations are INTERLEAVED over consecutive CTAs (a 1-d grid of world x tiles): the local tiles run at HBM speed, the
    // remote ones at NVLink speed, and a grid that did one destination after the other (blockIdx.y) finished the local block
    // first and only then started to use the links -- 0.481 ms instead of 0.40 ms for 16384^2 on two GPUs.
    const int h = (first_dst + (int)(blockIdx.x % (unsigned)world)) % world;
    const unsigned tile_id = blockIdx.x / (unsigned)world;
    // order 0: consecutive CTAs walk j (the source's contiguous dimension); order 1: they walk i (the DESTINATION's contiguous
    // dimension), so the CTAs in flight write neighbouring runs instead of runs one leading dimension apart
    const int tj = order ? tile_id / tiles_i : tile_id % tiles_j, ti = order ? tile_id % tiles_i : tile_id / tiles_j;
    // block-local coordinates: output tile rows i (0 .. cols_blk) = source columns, output tile columns j (0 .. rows_blk) = source rows
    const int64_t i0 = (int64_t)ti * 64, j0 = (int64_t)tj * 64;
    const T *src = b_local + (int64_t)h * rows_blk;                    // rows c0(h) of the local slab
    T *dst = reinterpret_cast<T *>(peers.dst[h]) + row_off;            // rows c1(g) of rank h's slab
    const int jq = t % CH;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
        const int iq = t / CH + (256 / CH) * p;
        V16<T> blk[V];
#pragma unroll
        for (int r = 0; r < V; ++r) blk[r] = ldg_v(src + (j0 + V * jq) + (i0 + V * iq + r) * ld_src);
#pragma unroll
        for (int c = 0; c < V; ++c) {
            V16<T> w;
#pragma unroll
            for (int r = 0; r < V; ++r) w.e[r] = blk[r].e[c];
            tile[(V * jq + c) * CH + (iq ^ jq)] = w;
        }
    }
    a2a_wait_ready(flags, first_dst, h, call, ctl);   // thread 0, before the barrier: nobody stores into h's slab earlier
    __syncthreads();
    const int ic = t % CH, jr = t / CH;
#pragma unroll
    for (int sidx = 0; sidx < 64 / ROWS_PER_PASS; ++sidx) {
        const int jl = jr + ROWS_PER_PASS * sidx;
        const V16<T> w = tile[jl * CH + (ic ^ ((jl / V) & (CH - 1)))];
        const uint4 r = *reinterpret_cast<const uint4 *>(&w);
        T *p = dst + (i0 + V * ic) + (j0 + jl) * ld_dst;
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
    }
    if (epoch == 0) return;   // completion is signalled by the separate barrier kernel (a2a.fused_signal = 0)
    // ---- completion, inside the data kernel (a2a.fused_signal, r05 A/B -- measured slower, see the launcher): ONE launch does
    // transpose, all-to-all and barrier.
    // Each CTA makes its tile visible system-wide (barrier, then ONE system-scope fence by thread 0: cumulative over the CTA's
    // stores) and counts itself done for its destination h on a LOCAL counter; the CTA that completes a destination raises this
    // rank's arrival flag at rank h (the same words b200_sym_barrier uses).  The fences of different CTAs overlap with each
    // other and with the data phase; only the last tile's fence is exposed, where the barrier kernel paid a launch and a fence
    // over everything.  The CTA that finishes last overall then waits for every peer's arrival flag, so that the kernel's
    // completion still means "my slab is complete" -- the contract of the two-launch form.
    __syncthreads();
    if (t == 0) {
        unsigned long long *local = reinterpret_cast<unsigned long long *>(flags.dst[first_dst]);   // first_dst == this rank
        __threadfence_system();
        const unsigned per_dst = gridDim.x / (unsigned)world;
        unsigned *done = reinterpret_cast<unsigned *>(local + 16);                                     // done[h], h < 8; [8] = all
        if (atomicAdd(done + h, 1u) == per_dst - 1) {
            done[h] = 0;
            unsigned long long *theirs = reinterpret_cast<unsigned long long *>(flags.dst[h]) + first_dst;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
        }
        __threadfence();
        s_last = atomicAdd(done + 8, 1u) == gridDim.x - 1;
        if (s_last) done[8] = 0;
    }
    __syncthreads();
    if (s_last && t < world)
        a2a_spin_until(reinterpret_cast<const unsigned long long *>(flags.dst[first_dst]) + t, epoch, ctl);
}

// The same data kernel with the store phase handed to TMA (opt-in, `a2a.tma_store`): the transposed tile is laid out densely and
// ONE cp.async.bulk.tensor.2d store per tile (64 x 64, or 256 x 16 for 1-KiB destination runs) moves it into the destination slab.
// Measured the same as the plain 128-bit stores (tools/a2a_time.py): linear copies to a peer reach 699 GB/s with either store
// type (tools/p2p_bw.py; cudaMemcpyPeer: 544), this kernel 681.  The dense layout costs bank conflicts on the transposing
// shared-memory writes, which do not matter: a tile needs ~7000 cycles of NVLink time per SM.
struct A2AMaps {
    CUtensorMap m[kSymMaxRanks];
};
// Tile = TI (along the destination's contiguous dimension i) x TJ: 256 x 16 gives 1 KiB contiguous runs in the destination slab
// (NVLink likes long runs: 64 x 64 tiles, i.e. 256-byte runs, stay at the 555 GB/s of plain stores), 64 x 64 is the fallback shape.
template <typename T, int TI, int TJ>
__global__ void __launch_bounds__(256) transpose_a2a_tma_kernel(const __grid_constant__ A2AMaps maps, const T *__restrict__ b_local,
                                                                int64_t ld_src, int64_t rows_blk, int row_off, int tiles_i, int tiles_j, int first_dst,
                                                                int order, int world, A2APeers flags, unsigned long long call, A2ACtl ctl) {
    constexpr int V = 16 / sizeof(T), CI = TI / V, CJ = TJ / V, PASSES = CI * CJ / 256;
    static_assert(CI * CJ % 256 == 0 && 256 % CJ == 0, "tile must be a whole number of passes of 256 threads");
    __shared__ __align__(128) V16<T> tile[TJ * CI];   // [j][chunk of i], dense: exactly the box the tensor map describes
    const int t = threadIdx.x;
    a2a_post_ready(flags, world, first_dst, call);
    const int h = (first_dst + (int)(blockIdx.x % (unsigned)world)) % world;   // destinations interleaved over consecutive CTAs
    const unsigned tile_id = blockIdx.x / (unsigned)world;
    const int tj = order ? tile_id / tiles_i : tile_id % tiles_j, ti = order ? tile_id % tiles_i : tile_id / tiles_j;
    const int64_t i0 = (int64_t)ti * TI, j0 = (int64_t)tj * TJ;
    const T *src = b_local + (int64_t)h * rows_blk;
    const int jq = t % CJ;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
        const int iq = t / CJ + (256 / CJ) * p;
        V16<T> blk[V];
#pragma unroll
        for (int r = 0; r < V; ++r) blk[r] = ldg_v(src + (j0 + V * jq) + (i0 + V * iq + r) * ld_src);
#pragma unroll
        for (int c = 0; c < V; ++c) {
            V16<T> w;
#pragma unroll
            for (int r = 0; r < V; ++r) w.e[r] = blk[r].e[c];
            tile[(V * jq + c) * CI + iq] = w;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk engine
    __syncthreads();
    a2a_wait_ready(flags, first_dst, h, call, ctl);
    if (t == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&maps.m[h]),
                     "r"(row_off + (int)i0), "r"((int)j0), "r"(pipe::smem_u32(tile))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the CTA may not retire (and free its smem) before the store
    }
}

// One CTA: make everything this GPU wrote before visible system-wide, raise `epoch` in every peer's flag word for this rank,
// wait until every peer has raised its word here.
__global__ void sym_barrier_kernel(A2APeers flags, int world, int rank, unsigned long long epoch, A2ACtl ctl) {
    __threadfence_system();
    const int r = threadIdx.x;
    if (r < world) {
        unsigned long long *theirs = reinterpret_cast<unsigned long long *>(flags.dst[r]) + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
        a2a_spin_until(reinterpret_cast<const unsigned long long *>(flags.dst[rank]) + r, epoch, ctl);
    }
}

static A2ACtl make_ctl(const Sym *h) {
    A2ACtl c;
    c.timeout_ns = (unsigned long long)tune_get("a2a.timeout_ms", 10000) * 1000000ull;
    c.error = h->error_dev;
    if (!c.error) c.timeout_ns = 0;
    return c;
}
static int check_sym_error(const Sym *h, const char *who) {
    if (h->error_host && *h->error_host != 0)
        return set_error(B200_ERR_CUDA, "%s: an earlier call on this symmetric buffer timed out waiting for a peer (a rank skipped the "
                         "collective or its launch failed); the ranks are out of step -- free the buffer and connect again", who);
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_sym_alloc(void **handle, size_t bytes) {
    if (!handle) return set_error(B200_ERR_INVALID_VALUE, "b200_sym_alloc: handle is NULL");
    B200_TRY(ensure_device());
    Sym *h = new Sym();
    h->bytes = (bytes + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&h->local, h->bytes + kSymFlagBytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        delete h;
        return set_error(B200_ERR_OOM, "b200_sym_alloc: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    B200_CUDA(cudaMemset((char *)h->local + h->bytes, 0, kSymFlagBytes));
    B200_CUDA(cudaHostAlloc((void **)&h->error_host, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    *h->error_host = 0;
    B200_CUDA(cudaHostGetDevicePointer((void **)&h->error_dev, h->error_host, 0));
    *handle = h;
    return B200_OK;
}

extern "C" b200_status b200_sym_export(void *handle, void *ipc64) {
    if (!handle || !ipc64) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    cudaIpcMemHandle_t ipc;
    B200_CUDA(cudaIpcGetMemHandle(&ipc, ((Sym *)handle)->local));
    memcpy(ipc64, &ipc, sizeof(ipc));
    return B200_OK;
}

extern "C" b200_status b200_sym_connect(void *handle, int32_t world, int32_t rank, const void *all_handles) {
    if (!handle || !all_handles) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (world < 1 || world > kSymMaxRanks || rank < 0 || rank >= world)
        return set_error(B200_ERR_INVALID_VALUE, "b200_sym_connect: bad rank %d / world %d (1..8 ranks of one NVLink domain)", rank, world);
    B200_TRY(ensure_device());
    Sym *h = (Sym *)handle;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            h->peer[r] = h->local;
            continue;
        }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, (const char *)all_handles + (size_t)r * B200_IPC_HANDLE_BYTES, sizeof(ipc));
        B200_CUDA(cudaIpcOpenMemHandle(&h->peer[r], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    h->world = world;
    h->rank = rank;
    return B200_OK;
}

// Single-process form of b200_sym_connect: the ranks are GPUs of ONE process ("per-GPU kernels launched
// concurrently from one process"), so the peers' buffers are handed over as pointers (b200_sym_ptr(peer_handle, -1) taken
// while that GPU was current) instead of IPC handles; peer access is enabled here.
extern "C" b200_status b200_sym_connect_ptrs(void *handle, int32_t world, int32_t rank, void *const *peer_bufs) {
    if (!handle || !peer_bufs) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (world < 1 || world > kSymMaxRanks || rank < 0 || rank >= world)
        return set_error(B200_ERR_INVALID_VALUE, "b200_sym_connect_ptrs: bad rank %d / world %d", rank, world);
    B200_TRY(ensure_device());
    Sym *h = (Sym *)handle;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            h->peer[r] = h->local;
            continue;
        }
        if (!peer_bufs[r]) return set_error(B200_ERR_INVALID_VALUE, "b200_sym_connect_ptrs: NULL buffer for rank %d", r);
        cudaPointerAttributes at;
        B200_CUDA(cudaPointerGetAttributes(&at, peer_bufs[r]));
        int cur = -1;
        B200_CUDA(cudaGetDevice(&cur));
        if (at.device != cur) {   // (a "peer" on the same GPU needs no enabling: used by the single-GPU tests)
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else B200_CUDA(e);
        }
        h->peer[r] = peer_bufs[r];
    }
    h->world = world;
    h->rank = rank;
    h->ipc = false;
    return B200_OK;
}

extern "C" b200_status b200_sym_ptr(void *handle, int32_t rank, void **ptr) {
    if (!handle || !ptr) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    Sym *h = (Sym *)handle;
    if (rank == -1) rank = h->rank < 0 ? 0 : h->rank;
    if (h->world < 1) {   // not connected yet: only the own buffer is known
        *ptr = h->local;
        return B200_OK;
    }
    if (rank < 0 || rank >= h->world) return set_error(B200_ERR_INVALID_VALUE, "b200_sym_ptr: rank %d out of range", rank);
    *ptr = h->peer[rank];
    return B200_OK;
}

extern "C" b200_status b200_sym_barrier(void *handle) {
    if (!handle) return set_error(B200_ERR_INVALID_VALUE, "b200_sym_barrier: handle is NULL");
    Sym *h = (Sym *)handle;
    if (h->world < 1) return set_error(B200_ERR_INVALID_VALUE, "b200_sym_barrier: b200_sym_connect has not been called");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    A2APeers flags;
    for (int r = 0; r < kSymMaxRanks; ++r) flags.dst[r] = r < h->world ? (char *)h->peer[r] + h->bytes : nullptr;
    B200_TRY(check_sym_error(h, "b200_sym_barrier"));
    sym_barrier_kernel<<<1, 32, 0, st>>>(flags, h->world, h->rank, ++h->epoch, make_ctl(h));
    B200_CHECK_LAUNCH("sym_barrier_kernel");
    return B200_OK;
}

extern "C" b200_status b200_sym_free(void *handle) {
    if (!handle) return B200_OK;
    Sym *h = (Sym *)handle;
    B200_CUDA(cudaDeviceSynchronize());
    for (int r = 0; r < h->world; ++r)
        if (h->ipc && r != h->rank && h->peer[r]) (void)cudaIpcCloseMemHandle(h->peer[r]);
    if (h->local) (void)cudaFree(h->local);
    if (h->error_host) (void)cudaFreeHost(h->error_host);
    delete h;
    return B200_OK;
}

extern "C" b200_status b200_transpose_alltoall(void *a_sym, const void *b_local, int64_t N0, int64_t N1, int32_t elsize) {
    if (!a_sym || !b_local) return set_error(B200_ERR_INVALID_VALUE, "b200_transpose_alltoall: NULL argument");
    Sym *h = (Sym *)a_sym;
    const int G = h->world;
    if (G < 1) return set_error(B200_ERR_INVALID_VALUE, "b200_transpose_alltoall: b200_sym_connect has not been called");
    if (elsize != 4 && elsize != 8) return set_error(B200_ERR_UNSUPPORTED, "b200_transpose_alltoall: element size %d not in {4,8}", elsize);
    if (N0 <= 0 || N1 <= 0 || N0 % (64 * G) != 0 || N1 % (64 * G) != 0)
        return set_error(B200_ERR_UNSUPPORTED, "b200_transpose_alltoall: N0 and N1 must be positive multiples of 64 x world (got %lld, %lld)",
                         (long long)N0, (long long)N1);
    if ((size_t)(N1 * (N0 / G)) * (size_t)elsize > h->bytes)
        return set_error(B200_ERR_INVALID_VALUE, "b200_transpose_alltoall: the symmetric output slab is smaller than N1 x N0/world elements");
    if (((uintptr_t)b_local & 15) != 0) return set_error(B200_ERR_UNSUPPORTED, "b200_transpose_alltoall: b must be 16-byte aligned");
    B200_TRY(check_sym_error(h, "b200_transpose_alltoall"));
    const A2ACtl ctl = make_ctl(h);
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    const int64_t rows_blk = N0 / G, cols_blk = N1 / G;   // a block of b_local: rows_blk x cols_blk; its transpose: cols_blk x rows_blk
    A2APeers peers, flags;
    for (int r = 0; r < kSymMaxRanks; ++r) {
        peers.dst[r] = r < G ? h->peer[r] : nullptr;
        flags.dst[r] = r < G ? (char *)h->peer[r] + h->bytes : nullptr;
    }
    // `a2a.debug_no_sync` (profiling only): this rank's data kernel runs alone -- it neither waits for the peers' "ready" flags
    // nor joins the trailing barrier -- so that ONE rank of a multi-rank transpose can be replayed under ncu (which serialises
    // kernels: two ranks spinning for each other would deadlock) while its stores still cross NVLink (tools/nvlink_proof.py).
    const bool no_sync = tune_get("a2a.debug_no_sync", 0) != 0;
    const unsigned long long call = no_sync ? 0ull : ++h->a2a_calls;
    const int tiles_i = (int)(cols_blk / 64), tiles_j = (int)(rows_blk / 64);
    if ((int64_t)tiles_i * tiles_j * G > INT32_MAX) return set_error(B200_ERR_UNSUPPORTED, "b200_transpose_alltoall: too many tiles");
    const dim3 grid((unsigned)(tiles_i * tiles_j * G));
    const int order = tune_get("a2a.order", 1);
    // a2a.tma_store = 1 selects the TMA-store variant; measured equal to the plain stores (0.3944 vs 0.3945 ms for 16384^2 on two
    // GPUs = 681 GB/s of NVLink egress, where a linear copy kernel reaches 699 and cudaMemcpyPeer 544), so the simpler kernel is the default
    if (tune_get("a2a.tma_store", 0) != 0 && N1 < (1LL << 31) && rows_blk < (1LL << 31)) {
        const bool wide = cols_blk % 256 == 0 && rows_blk % 16 == 0 && tune_get("a2a.wide_tiles", 1) != 0;
        const int TI = wide ? 256 : 64, TJ = wide ? 16 : 64;
        A2AMaps maps;
        memset(&maps, 0, sizeof(maps));
        for (int r = 0; r < G; ++r)   // rank r's slab: N1 (contiguous) x N0/G, box TI x TJ, no swizzle
            B200_TRY(pipe::make_tmap_2d(&maps.m[r], elsize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, elsize,
                                        h->peer[r], N1, rows_blk, N1, TI, TJ, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE));
        const int tj_n = (int)(rows_blk / TJ);
        if ((cols_blk / TI) * tj_n * G > INT32_MAX) return set_error(B200_ERR_UNSUPPORTED, "b200_transpose_alltoall: too many tiles");
        const dim3 g2((unsigned)((cols_blk / TI) * tj_n * G));
        const int roff = (int)(h->rank * cols_blk);
        if (elsize == 4 && wide)
            transpose_a2a_tma_kernel<uint32_t, 256, 16><<<g2, 256, 0, st>>>(maps, (const uint32_t *)b_local, N0, rows_blk, roff, (int)(cols_blk / TI), tj_n, h->rank, order, G, flags, call, ctl);
        else if (elsize == 4)
            transpose_a2a_tma_kernel<uint32_t, 64, 64><<<g2, 256, 0, st>>>(maps, (const uint32_t *)b_local, N0, rows_blk, roff, (int)(cols_blk / TI), tj_n, h->rank, order, G, flags, call, ctl);
        else if (wide)
            transpose_a2a_tma_kernel<uint64_t, 256, 16><<<g2, 256, 0, st>>>(maps, (const uint64_t *)b_local, N0, rows_blk, roff, (int)(cols_blk / TI), tj_n, h->rank, order, G, flags, call, ctl);
        else
            transpose_a2a_tma_kernel<uint64_t, 64, 64><<<g2, 256, 0, st>>>(maps, (const uint64_t *)b_local, N0, rows_blk, roff, (int)(cols_blk / TI), tj_n, h->rank, order, G, flags, call, ctl);
        B200_CHECK_LAUNCH("transpose_a2a_tma_kernel");
        return no_sync ? B200_OK : b200_sym_barrier(a_sym);
    }
    // a2a.fused_signal = 1: completion signalled inside the data kernel (one launch) instead of by the barrier kernel (two).
    // Measured on 2 GPUs, 16384^2 Float32 (r05): 0.632 ms against 0.394 ms -- a system-scope fence per CTA makes every CTA wait
    // for the NVLink acknowledgements of its own tile, and 8192 such waits cost far more than the one fence + one launch of the
    // barrier kernel.  Opt-in, parity-tested.
    const bool fused = !no_sync && tune_get("a2a.fused_signal", 0) != 0;
    const unsigned long long ep = fused ? ++h->epoch : 0ull;
    if (elsize == 4)
        transpose_a2a_kernel<uint32_t><<<grid, 256, 0, st>>>(peers, (const uint32_t *)b_local, N0, N1, rows_blk, (int64_t)h->rank * cols_blk,
                                                            tiles_i, tiles_j, h->rank, order, G, flags, call, ctl, ep);
    else
        transpose_a2a_kernel<uint64_t><<<grid, 256, 0, st>>>(peers, (const uint64_t *)b_local, N0, N1, rows_blk, (int64_t)h->rank * cols_blk,
                                                            tiles_i, tiles_j, h->rank, order, G, flags, call, ctl, ep);
    B200_CHECK_LAUNCH("transpose_a2a_kernel");
    return (no_sync || fused) ? B200_OK : b200_sym_barrier(a_sym);
}
