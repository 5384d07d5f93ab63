// histogram.cu -- b200_histogram_i32: bins[v-1] += #{t : input[t] == v} for v in 1..nbins.
//
// Reference: histogram_kernel! (examples/histogram.jl:18-58).  The reference algorithm gives every
// 256-item group a shared histogram and merges all `gs` bins of every group into global memory with one
// atomic per bin (:52-54): for 2^30 inputs that is 2^30 shared + 2^30 global atomics.  Counts are integer
// sums, so ANY regrouping of the additions gives the same bits; this file regroups them so the kernel is
// bound by the 4 bytes/element it must read from HBM:
//
//   impl 0 "bytes" (nbins <= 256): every THREAD owns a private column of 256 one-byte counters in shared
//       memory (256*T bytes per CTA).  The byte for (bin, thread) sits at bin*T + (warp/4)*128 + lane*4 + warp%4,
//       i.e. lane l always hits bank l: increments are plain LDS.U8/IADD/STS.U8 with zero bank conflicts and
//       no atomics, independent of the input distribution (all-equal inputs cost the same as uniform ones).
//       Counters are flushed (dp4a byte sums, rotated 16-byte chunks, conflict-free) into per-thread 32-bit
//       accumulators before any of them can reach 256; one red.global.add per (bin, part) per CTA at the end.
//   impl 1 "atomics": per-warp private 32-bit shared histograms updated with shared-memory atomics
//       (also the path for 256 < nbins <= 14336).
//   larger nbins: one red.global.add per in-range element.
// Values outside 1..nbins are ignored, exactly like the reference's range test (:45-46).
#include <cstring>
#include <vector>

#include "../common.cuh"
#include "../tma_pipe.cuh"
#include "../tune.h"

namespace b200 {

__device__ __forceinline__ int4 ldg_stream_i4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// CTA b counts the int4 units [slice_cut(b), slice_cut(b + 1)).  The cuts fall on 512-byte lines of the INPUT ADDRESS (32 units = one
// warp-wide LDG.128), so that no warp request straddles a line: with cuts at n4 * b / grid every warp of a CTA with an odd cut read
// 17 sectors per request and the shared sectors came from DRAM twice -- ncu: 142.8 M DRAM sectors for 134.2 M of input (+6.4 %),
// i.e. the kernel already moved 7.0 TB/s for 6.55 TB/s of input.
__device__ __forceinline__ int64_t slice_cut(const int4 *in4, int64_t n4, unsigned b) {
    if (b == 0) return 0;
    if (b >= gridDim.x) return n4;
    const int64_t off = (int64_t)(((uintptr_t)in4 >> 4) & 31);   // units between the previous 512-byte line and in4
    const int64_t u = ((n4 * (int64_t)b / gridDim.x + off) & ~(int64_t)31) - off;
    return u < 0 ? 0 : u;
}

// ---------------------------------------------------------------------------------------------
// impl 0: thread-private byte counters
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void bump1(unsigned char *cnt, uint32_t base, int v, uint32_t nbins) {
    uint32_t b = (uint32_t)v - 1u;
    if (b < nbins) {
        unsigned char *p = cnt + b * T + base;
        *p = (unsigned char)(*p + 1);
    }
}
template <int T>
__device__ __forceinline__ void bump2(unsigned char *cnt, uint32_t base, int v1, int v2, uint32_t nbins) {
    uint32_t b1 = (uint32_t)v1 - 1u, b2 = (uint32_t)v2 - 1u;
    bool ok1 = b1 < nbins, ok2 = b2 < nbins;
    unsigned char *p1 = cnt + (ok1 ? b1 : 0u) * T + base;
    unsigned char *p2 = cnt + (ok2 ? b2 : 0u) * T + base;
    uint32_t c1 = *p1, c2 = *p2;  // both loads issue before either store
    bool same = ok1 && ok2 && (b1 == b2);
    if (ok1) *p1 = (unsigned char)(c1 + 1 + (same ? 1 : 0));
    if (ok2 && !same) *p2 = (unsigned char)(c2 + 1);
}

template <int T, int ILP>
__device__ __forceinline__ void bump4(unsigned char *cnt, uint32_t base, const int4 &v, uint32_t nbins) {
    if (ILP == 2) {
        bump2<T>(cnt, base, v.x, v.y, nbins);
        bump2<T>(cnt, base, v.z, v.w, nbins);
    } else {
        bump1<T>(cnt, base, v.x, nbins);
        bump1<T>(cnt, base, v.y, nbins);
        bump1<T>(cnt, base, v.z, nbins);
        bump1<T>(cnt, base, v.w, nbins);
    }
}

// Sum this thread's slice (256 bytes of row `bin`) of the byte counters into acc and clear it.
template <int T>
__device__ __forceinline__ void flush_bytes(unsigned char *cnt, uint32_t &acc) {
    const int t = threadIdx.x, lane = t & 31;
    const int bin = t & 255, part = t >> 8;
    uint4 *row = reinterpret_cast<uint4 *>(cnt + bin * T + part * 256);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int c = (k + lane) & 15;
        uint4 x = row[c];
        s = __dp4a(x.x, 0x01010101u, s);
        s = __dp4a(x.y, 0x01010101u, s);
        s = __dp4a(x.z, 0x01010101u, s);
        s = __dp4a(x.w, 0x01010101u, s);
        row[c] = make_uint4(0, 0, 0, 0);
    }
    acc += s;
}

// in4: 16-byte aligned input viewed as int4 units; the CTA owns units [u0, u1).
// P = int4 loads in flight per thread per batch (double buffered); a flush happens every PERIOD batches
// with PERIOD * P * 4 <= 255 so no byte counter can wrap.
template <int T, int P, int ILP>
__global__ void __launch_bounds__(T, 1) hist256_bytes_kernel(int32_t *__restrict__ bins, uint32_t nbins,
                                                             const int4 *__restrict__ in4, int64_t n4,
                                                             const int32_t *__restrict__ tail, int ntail) {
    static_assert(T % 256 == 0, "flush mapping needs T multiple of 256");
    constexpr int PERIOD = 255 / (P * 4);
    extern __shared__ __align__(16) unsigned char cnt[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < 256 * T / 16; i += T) reinterpret_cast<uint4 *>(cnt)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t base = (uint32_t)((warp >> 2) * 128 + lane * 4 + (warp & 3));
    uint32_t acc = 0;

    const int64_t u0 = slice_cut(in4, n4, blockIdx.x), u1 = slice_cut(in4, n4, blockIdx.x + 1);
    // batch k covers units u0 + k*P*T + p*T + t, p = 0..P-1
    const int64_t nbatch = (u1 - u0 + (int64_t)P * T - 1) / ((int64_t)P * T);
    int4 cur[P], nxt[P];
    auto load_batch = [&](int64_t k, int4 *buf) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            int64_t u = u0 + (k * P + p) * (int64_t)T + t;
            buf[p] = u < u1 ? ldg_stream_i4(in4 + u) : make_int4(0, 0, 0, 0);  // 0 is outside 1..nbins
        }
    };
    if (nbatch > 0) load_batch(0, cur);
    int since_flush = 0;
    for (int64_t k = 0; k < nbatch; ++k) {
        if (k + 1 < nbatch) load_batch(k + 1, nxt);
#pragma unroll
        for (int p = 0; p < P; ++p) bump4<T, ILP>(cnt, base, cur[p], nbins);
#pragma unroll
        for (int p = 0; p < P; ++p) cur[p] = nxt[p];
        if (++since_flush == PERIOD) {
            __syncthreads();
            flush_bytes<T>(cnt, acc);
            __syncthreads();
            since_flush = 0;
        }
    }
    // scalar tail (n % 4 elements and/or an unaligned head), done by CTA 0
    if (blockIdx.x == 0 && t < ntail) bump1<T>(cnt, base, tail[t], nbins);
    __syncthreads();
    flush_bytes<T>(cnt, acc);
    const uint32_t bin = t & 255;
    if (acc != 0 && bin < nbins) atomicAdd(reinterpret_cast<unsigned int *>(bins) + bin, acc);
}

// ---------------------------------------------------------------------------------------------
// impl 1: shared-memory atomics on `copies` private 32-bit histograms per CTA
// ---------------------------------------------------------------------------------------------
//
// The counting phase is one NOT-inlined device function shared by the plain kernel and by the fused
// histogram + all-reduce kernel, so both run exactly the same SASS for the hot loop: inlined into the larger fused
// kernel, ptxas scheduled the four in-flight LDG.128 differently and the loop lost 13 % (745 us instead of 650 us for
// 2^30 elements; profiles/r02_hist_fused_notes.md).
// `tail`/`ntail`: the scalar leftovers, counted by CTA 0 -- the n % 4 elements behind the int4 body and, when the input is not
// 16-byte aligned, the (< 4) head elements in front of it (`head`/`nhead`).
//
// Which int4 units does a CTA count?  Two plans (template parameter IL), same total:
//  * slices (IL = false): one contiguous slice per CTA, cut on 512-byte lines of the address (slice_cut);
//  * interleaved (IL = true): the input is dealt out in UNITS of T int4 (16 KiB for T = 1024) round-robin over the CTAs,
//    unit u -> CTA u % grid, unit boundaries on 512-byte lines (the < 32 int4 in front of the first line are CTA 0's).
//    Every CTA then walks the whole address range at the same pace, sees the same mix of DRAM channels and owns
//    floor / ceil(units / grid) units; contiguous slices finished up to 6 us apart at 2^27 elements (b200_p2p_timeline,
//    profiles/r05_slab_probe.txt), and a kernel ends with its slowest CTA.
// The loop is issue-sensitive (16 shared atomics + their range tests per 4 loads keep the schedulers ~2/3 busy; a variant with
// 25 % more integer instructions per batch ran 20 % slower, r05), so the main loop carries NO bounds predicates: `nfb` batches
// of P loads per thread are complete by construction and ONE predicated batch follows for whatever is left.
template <int T, int P, bool IL, bool I64 = false>
__device__ __noinline__ void hist_count_phase(uint32_t *sh, uint32_t nbins, const int4 *__restrict__ in4, int64_t n4,
                                              const int32_t *__restrict__ tail, int ntail, int copies,
                                              const int32_t *__restrict__ head, int nhead, int npf,
                                              uint32_t *dyn_ctr = nullptr, int dyn_rounds = 0) {
    const int t = threadIdx.x, warp = t >> 5;
    // Programmatic dependent launch (see launch_pdl): the next kernel of the stream may be scheduled as SMs drain; our own set-up
    // runs before we wait for the previous kernel's memory.  Back-to-back steps on input shards hide ~2 us of the ~9 us a launch costs.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t nb1 = nbins + 1u;   // every private copy carries one extra counter: the sink of out-of-range values
    uint32_t *mine = sh + (uint32_t)(warp % copies) * nb1;
    // Values outside 1..nbins are ignored by the reference's range test (examples/histogram.jl:45-46).  Here they are CLAMPED into
    // the sink counter instead of branched around: (unsigned)(v - 1) >= nbins -> nbins.  One IADD + one unsigned MIN per element
    // replace compare + branch + reconvergence in a loop whose limit is instruction issue (see below); the sink is never read.
    auto bump = [&](int v) { atomicAdd(mine + min((uint32_t)v - 1u, nbins), 1u); };
    // Int64 values (I64: two per int4, `lo` / `hi` halves): the same clamp in 64 bits
    auto bump64 = [&](int lo, int hi) {
        const unsigned long long b = (((unsigned long long)(uint32_t)hi << 32) | (uint32_t)lo) - 1ull;
        atomicAdd(mine + (uint32_t)min(b, (unsigned long long)nbins), 1u);
    };
    auto bump4 = [&](const int4 &v) {
        if (I64) {
            bump64(v.x, v.y);
            bump64(v.z, v.w);
        } else {
            bump(v.x);
            bump(v.y);
            bump(v.z);
            bump(v.w);
        }
    };
    // Round r of this CTA loads int4 number  first + r * rstride + t  (t = thread); rounds [0, nr) are complete (all T
    // threads in range), then `extra` more int4 (< T) follow at round nr for the first `extra` threads.
    int64_t first, rstride, nr, extra, lead = 0;
    // Work-stealing tail (IL plan, `dyn_ctr` given; opt-in A/B, `hist.dyn_rounds`): the last `dyn_rounds` rounds of units are not
    // dealt out but CLAIMED, P units (one batch) at a time, from a global counter -- a CTA that runs ahead takes more of them.  It
    // does what it should -- the CTAs end 2.0 us apart instead of 3.6 -- and still LOSES: the claimed batches stream slower than
    // the static loop (two barriers per batch, no unrolling across batches): 2^27 elements, fused, one rank: 78.7 us without,
    // 80.0 / 80.8 / 83.7 us with 4 / 16 / 32 rounds (profiles/r05_hist_probe3.txt).  dyn_first = first unit of that region.
    int64_t dyn_first = 0, dyn_nu = 0, dyn_extra = 0;
    uint32_t dyn_chunks = 0;
    if (IL) {
        lead = min((int64_t)((32 - (((uintptr_t)in4 >> 4) & 31)) & 31), n4);                 // int4 in front of the first 512-byte line
        const int64_t m4 = n4 - lead;
        int64_t nu = m4 / T;                                                                  // whole units behind it
        first = lead + (int64_t)blockIdx.x * T;
        rstride = (int64_t)gridDim.x * T;
        if (dyn_ctr != nullptr) {
            int64_t s_rounds = nu / gridDim.x - dyn_rounds;                                   // rounds every CTA gets statically
            if (s_rounds < 0) s_rounds = 0;
            dyn_first = s_rounds * gridDim.x;
            dyn_nu = nu;
            dyn_extra = m4 - nu * T;
            dyn_chunks = (uint32_t)((nu - dyn_first + (dyn_extra > 0 ? 1 : 0) + P - 1) / P);
            nr = s_rounds;
            extra = 0;
        } else {
            nr = nu > blockIdx.x ? (nu - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
            extra = (int64_t)(nu % gridDim.x) == (int64_t)blockIdx.x ? m4 - nu * T : 0;       // the partial last unit has one owner
        }
    } else {
        const int64_t u0 = slice_cut(in4, n4, blockIdx.x), u1 = slice_cut(in4, n4, blockIdx.x + 1);
        first = u0;
        rstride = T;
        nr = (u1 - u0) / T;
        extra = (u1 - u0) - nr * T;
    }
    const int64_t nfb = nr / P;                      // complete batches
    const int left = (int)(nr - nfb * P);            // complete rounds behind them (< P)
    const int4 *ptr = in4 + first + t;
    const int64_t bstep = rstride * P;
    int4 cur[P], nxt[P];
    // CTAs that become resident while the previous kernel of the stream is still in its tail (the fused kernel's exchange leaves
    // HBM idle for ~5 us) pull their first rounds into L2 before the dependency wait.  A prefetch has no architectural effect --
    // L2 is the point of coherence, anything the previous kernel still writes lands there too -- and the first loads then come
    // from L2 instead of paying the DRAM ramp.  One thread per round: 16 KiB per instruction.
    if (t < npf && (int64_t)t < nr)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in4 + first + (int64_t)t * rstride), "r"(T * 16) : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");   // everything before us in the stream is complete and visible
    // the first batch of loads goes out BEFORE the private copies are cleared: the clear runs under the loads' latency
    if (nfb > 0) {
#pragma unroll
        for (int p = 0; p < P; ++p) cur[p] = ldg_stream_i4(ptr + p * rstride);
    }
    for (uint32_t i = t; i < nb1 * (uint32_t)copies; i += T) sh[i] = 0;
    __syncthreads();
    if (IL && blockIdx.x == 0 && t < lead) bump4(ldg_stream_i4(in4 + t));
    for (int64_t k = 0; k < nfb; ++k) {
        ptr += bstep;
        if (k + 1 < nfb) {
#pragma unroll
            for (int p = 0; p < P; ++p) nxt[p] = ldg_stream_i4(ptr + p * rstride);
        }
#pragma unroll
        for (int p = 0; p < P; ++p) bump4(cur[p]);
#pragma unroll
        for (int p = 0; p < P; ++p) cur[p] = nxt[p];
    }
    // the one predicated batch: `left` complete rounds, then the partial round
#pragma unroll
    for (int p = 0; p < P; ++p)
        cur[p] = (p < left || (p == left && t < extra)) ? ldg_stream_i4(ptr + p * rstride) : make_int4(0, 0, 0, 0);   // 0 is outside 1..nbins
#pragma unroll
    for (int p = 0; p < P; ++p) bump4(cur[p]);
    if (IL && dyn_ctr != nullptr) {
        // claimed batches: thread 0 keeps two claims ahead (the atomic's round trip runs under a batch of counting); the claims
        // travel through two spare words behind the private copies
        uint32_t *idx = sh + nb1 * (uint32_t)copies + 1;
        if (t == 0) {
            idx[0] = atomicAdd(dyn_ctr, 1u);
            idx[1] = atomicAdd(dyn_ctr, 1u);
        }
        __syncthreads();
        const int4 *dbase = in4 + lead + dyn_first * T + t;
        auto load_chunk = [&](uint32_t c, int4 *buf) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int64_t u = dyn_first + (int64_t)c * P + p;
                const bool valid = u < dyn_nu || (u == dyn_nu && t < dyn_extra);
                buf[p] = valid ? ldg_stream_i4(dbase + ((int64_t)c * P + p) * T) : make_int4(0, 0, 0, 0);
            }
        };
        uint32_t c = idx[0];
        if (c < dyn_chunks) load_chunk(c, cur);
        for (int k = 0; c < dyn_chunks; ++k) {
            const uint32_t cn = idx[(k + 1) & 1];
            if (cn < dyn_chunks) load_chunk(cn, nxt);
            uint32_t fresh = 0;
            if (t == 0) fresh = atomicAdd(dyn_ctr, 1u);
#pragma unroll
            for (int p = 0; p < P; ++p) bump4(cur[p]);
#pragma unroll
            for (int p = 0; p < P; ++p) cur[p] = nxt[p];
            __syncthreads();                 // every thread has read both claims
            if (t == 0) idx[k & 1] = fresh;
            __syncthreads();
            c = cn;
        }
    }
    if (I64) {
        if (blockIdx.x == 0 && t < ntail) {
            const unsigned long long v = reinterpret_cast<const unsigned long long *>(tail)[t];
            bump64((int)(uint32_t)v, (int)(uint32_t)(v >> 32));
        }
    } else {
        if (blockIdx.x == 0 && t < ntail) bump(tail[t]);
        if (blockIdx.x == 0 && t >= 32 && t < 32 + nhead) bump(head[t - 32]);
    }
    __syncthreads();
}

// Fold the private copies into copy 0 (sh[b], b < nbins); ends with the result visible to thread b only -- callers synchronise.
// With few bins (nbins <= T / 4, the 256-bin case) four threads share a column: thread (b, q) sums the copies of group q into
// the group's first copy (nobody else reads that column of those copies), then thread (b, 0) adds the four group sums: 8 + 4
// dependent shared loads instead of 32 on the critical path of every CTA.
template <int T>
__device__ __forceinline__ void fold_copies(uint32_t *sh, uint32_t nbins, int copies) {
    const int t = threadIdx.x;
    const uint32_t nb1 = nbins + 1u;
    if (nbins * 4u <= (uint32_t)T && copies >= 4) {
        const uint32_t b = (uint32_t)t % nbins, q = (uint32_t)t / nbins;
        const uint32_t per = ((uint32_t)copies + 3u) / 4u, c0 = q * per;
        if (q < 4u && c0 < (uint32_t)copies) {
            uint32_t s = 0;
            for (uint32_t c = c0; c < c0 + per && c < (uint32_t)copies; ++c) s += sh[c * nb1 + b];
            sh[c0 * nb1 + b] = s;
        }
        __syncthreads();
        if (q == 0) {
            uint32_t s = sh[b];
            for (uint32_t g = 1; g < 4u && g * per < (uint32_t)copies; ++g) s += sh[g * per * nb1 + b];
            sh[b] = s;
        }
    } else {
        for (uint32_t b = t; b < nbins; b += T) {
            uint32_t s = 0;
            for (int c = 0; c < copies; ++c) s += sh[(uint32_t)c * nb1 + b];
            sh[b] = s;   // only thread (b mod T) touches column b
        }
    }
}

template <int T, int P>
__global__ void __launch_bounds__(T) hist_atomic_kernel(int32_t *__restrict__ bins, uint32_t nbins,
                                                        const int4 *__restrict__ in4, int64_t n4,
                                                        const int32_t *__restrict__ tail, int ntail, int copies, int interleave) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem_raw);
    if (interleave & 1)   // `interleave`: bit 0 = unit plan, bits 8.. = rounds to prefetch into L2 before the dependency wait
        hist_count_phase<T, P, true>(sh, nbins, in4, n4, tail, ntail, copies, nullptr, 0, interleave >> 8);
    else
        hist_count_phase<T, P, false>(sh, nbins, in4, n4, tail, ntail, copies, nullptr, 0, interleave >> 8);
    fold_copies<T>(sh, nbins, copies);
    for (uint32_t b = threadIdx.x; b < nbins; b += T) {   // thread b reads back what it wrote
        const uint32_t s = sh[b];
        if (s) atomicAdd(reinterpret_cast<unsigned int *>(bins) + b, s);
    }
}

// ---------------------------------------------------------------------------------------------
// impl 2 (A/B only, `hist.impl = 2`): the same counting with the input FED BY TMA -- one elected thread keeps a ring of
// 16 KiB cp.async.bulk loads in flight (units dealt out round-robin like the interleaved plan), the 1024 threads read their
// int4 from shared memory (LDS.128) instead of global memory and count it with the same shared atomics.  Measured against the
// LDG feed in profiles/r05_hist_probe.txt; the LDG kernel stays the default unless this one wins.
// ---------------------------------------------------------------------------------------------
template <int T, int STAGES>
__global__ void __launch_bounds__(T) hist_tma_kernel(int32_t *__restrict__ bins, uint32_t nbins, const int4 *__restrict__ in4, int64_t n4,
                                                     const int32_t *__restrict__ tail, int ntail, int copies, uint32_t ring_off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem_raw);
    int4 *ring = reinterpret_cast<int4 *>(smem_raw + ring_off);                 // STAGES x T int4
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t nb1 = nbins + 1u;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (uint32_t i = t; i < nb1 * (uint32_t)copies; i += T) sh[i] = 0;
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) {
            pipe::mbar_init(&full[s], 1);
            pipe::mbar_init(&empty[s], T / 32);
        }
        pipe::mbar_fence_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    uint32_t *mine = sh + (uint32_t)(warp % copies) * nb1;
    auto bump = [&](int v) { atomicAdd(mine + min((uint32_t)v - 1u, nbins), 1u); };
    auto bump4 = [&](const int4 &v) {
        bump(v.x);
        bump(v.y);
        bump(v.z);
        bump(v.w);
    };
    const int64_t lead = min((int64_t)((32 - (((uintptr_t)in4 >> 4) & 31)) & 31), n4);
    if (blockIdx.x == 0 && t < lead) bump4(ldg_stream_i4(in4 + t));
    const int64_t m4 = n4 - lead, nu = m4 / T;
    const int64_t first = lead + (int64_t)blockIdx.x * T, rstride = (int64_t)gridDim.x * T;
    const int64_t nr = nu > blockIdx.x ? (nu - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t extra = (int64_t)(nu % gridDim.x) == (int64_t)blockIdx.x ? m4 - nu * T : 0;
    auto issue = [&](int64_t r) {                                               // thread 0
        const int s = (int)(r % STAGES);
        pipe::mbar_expect_tx(&full[s], T * 16);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         pipe::smem_u32(ring + (size_t)s * T)),
                     "l"(in4 + first + r * rstride), "r"(T * 16), "r"(pipe::smem_u32(&full[s]))
                     : "memory");
    };
    if (t == 0)
        for (int64_t r = 0; r < nr && r < STAGES; ++r) issue(r);
    for (int64_t r = 0; r < nr; ++r) {
        const int s = (int)(r % STAGES);
        const uint32_t parity = (uint32_t)((r / STAGES) & 1);
        pipe::mbar_wait(&full[s], parity);
        const int4 v = ring[(size_t)s * T + t];
        __syncwarp();
        if (lane == 0) pipe::mbar_arrive(&empty[s]);                            // this warp has read stage s
        bump4(v);
        if (t == 0 && r + STAGES < nr) {
            pipe::mbar_wait(&empty[s], parity);                                 // every warp has: the stage may be refilled
            issue(r + STAGES);
        }
    }
    if (t < extra) bump4(ldg_stream_i4(in4 + first + nr * rstride + t));        // the partial last unit (one owner)
    if (blockIdx.x == 0 && t < ntail) bump(tail[t]);
    __syncthreads();
    for (uint32_t b = t; b < nbins; b += T) {
        uint32_t sum = 0;
        for (int c = 0; c < copies; ++c) sum += sh[(uint32_t)c * nb1 + b];
        if (sum) atomicAdd(reinterpret_cast<unsigned int *>(bins) + b, sum);
    }
}

// ---------------------------------------------------------------------------------------------
// fused histogram + all-reduce over NVLink peer memory (b200_histogram_i32_allreduce)
// ---------------------------------------------------------------------------------------------
// The CTAs merge their counts into a device-local partial; the LAST CTA to finish (ticket counter) then stores that
// partial into a slot of EVERY rank's symmetric buffer through NVLink peer memory (cudaIpc mappings) and sums the slots
// the other ranks stored into its own buffer.  Each slot word is an 8-byte {count, tag} pair written with ONE store, so
// data and "it has arrived" travel together and no fence, flag or atomic crosses NVLink (the LL idea of NCCL): a
// system-scope fence alone costs ~7 us on this box (b200_p2p_debug timeline), more than a whole NCCL all-reduce of 1 KiB.
// tag = step + 1; slots are double-buffered by step parity, which is race-free because a rank can start step e+2 only
// after every rank has finished step e+1, hence has long read its step-e slots.  Integer adds: bit-identical to
// b200_histogram_i32 + b200_allreduce_sum_i32 in any order.  One launch per step, no NCCL kernel.
//
// What is on the critical path after the slowest CTA of the slowest rank has left the counting loop (r05): fold of the private
// copies -> red.global into the partial -> fence + ticket -> [last CTA] partial -> slot stores over NVLink -> poll -> bins.
// The poll used to walk the 8 sources of a bin one after the other, every step a dependent ~0.7 us system-scope load: ~5 us
// for 8 ranks.  Now the (bin, source) pairs are spread over all threads of the CTA -- 8 consecutive lanes hold the 8 sources of
// one bin, every thread has all its slot loads in flight at once, and the sources are summed with three shuffles -- so the
// poll costs one round trip (plus re-polls of slots that have not landed); the value the sum is added to (`bins[b] +=`) is
// fetched under the poll, and the fold of the private copies is four-way parallel.  Timeline of a step: DESIGN.md section 6.
struct HistPeer {
    int world;
    int rank;
    uint32_t epoch;                   // step counter, identical on all ranks
    uint32_t nbins_max;               // slot stride
    uint32_t *counter;                // local ticket counter (zero between launches)
    uint32_t *partial;                // local per-bin partial (zero between launches)
    unsigned long long *stamps;       // %globaltimer stamps of the last launch (b200_p2p_debug / b200_p2p_timeline)
    unsigned long long *slots[8];     // rank r's slot array: [2 halves][8 source ranks][nbins_max] x {count, tag}
    uint32_t *error_flag;             // pinned host word (zero-copy): set when the poll below gives up
    unsigned long long timeout_ns;    // 0 = wait for ever
    unsigned long long *cta_slots;    // local [kLLMaxCtas][kLLMaxBins] x {count, tag}: the intra-GPU merge (nullptr: ticket + partial)
};
constexpr uint32_t kLLMaxBins = 1024, kLLMaxCtas = 160;
__device__ __forceinline__ void st_slot_gpu(unsigned long long *p, uint32_t count, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(count), "r"(tag) : "memory");
}
__device__ __forceinline__ void ld_slot_gpu(const unsigned long long *p, uint32_t &count, uint32_t &tag) {
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(count), "=r"(tag) : "l"(p) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
__device__ __forceinline__ void st_slot(unsigned long long *p, uint32_t count, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(count), "r"(tag) : "memory");
}
__device__ __forceinline__ void ld_slot(const unsigned long long *p, uint32_t &count, uint32_t &tag) {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(count), "=r"(tag) : "l"(p) : "memory");
}

template <int T>
__device__ __noinline__ void fused_exchange(uint32_t *sh, uint32_t nbins, int copies, int32_t *__restrict__ bins,
                                            const HistPeer &peer) {
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        const unsigned long long now = globaltimer_ns();
        atomicMax(peer.stamps + 0, now);                                                 // slowest CTA leaves the counting loop
        atomicMin(peer.stamps + 4, now);                                                 // fastest CTA leaves the counting loop
    }
    fold_copies<T>(sh, nbins, copies);   // copy 0 <- sum of the private copies
    const uint32_t nb1 = nbins + 1u;
    uint32_t &ticket = sh[nb1 * (uint32_t)copies];   // one spare word behind the private copies (dynamic smem)
    __syncthreads();
    const uint32_t tag = peer.epoch + 1u;
    if (peer.cta_slots != nullptr && nbins <= kLLMaxBins && gridDim.x <= kLLMaxCtas) {
        // ---- intra-GPU merge, LL style (r05 A/B, opt-in).  Ticket + red.global + fence + counter cost the last CTA three dependent
        // L2 round trips (~3 us, b200_p2p_timeline) after the slowest CTA left the counting loop.  Here every CTA stores its folded
        // counts as {count, tag} words into its own row of a local slot array -- one 8-byte store carries data and "it is there",
        // no fence, no atomic -- and CTA 0, the reducer, polls the rows and adds them up.  Measured slower (see the launcher).
        if (blockIdx.x != 0) {
            unsigned long long *row = peer.cta_slots + (size_t)blockIdx.x * kLLMaxBins;
            for (uint32_t b = t; b < nbins; b += T) st_slot_gpu(row + b, sh[b], tag);
            return;
        }
        // thread (part, b): rows part + 1, part + 1 + parts, ... of bin b; `parts` threads share a bin
        uint32_t parts = (uint32_t)T / nbins;
        if (parts > 16u) parts = 16u;
        if (parts < 1u) parts = 1u;                       // nbins <= kLLMaxBins == T
        const uint32_t part = (uint32_t)t / nbins, b = (uint32_t)t % nbins;
        uint32_t acc = 0;
        if (part < parts) {
            for (uint32_t c0 = 1u + part; c0 < gridDim.x; c0 += 8u * parts) {            // 8 rows in flight per thread
                uint32_t v[8], g[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t c = c0 + (uint32_t)k * parts;
                    v[k] = 0;
                    g[k] = tag;
                    if (c < gridDim.x) ld_slot_gpu(peer.cta_slots + (size_t)c * kLLMaxBins + b, v[k], g[k]);
                }
                while (true) {
                    bool pending = false;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (g[k] != tag) {
                            ld_slot_gpu(peer.cta_slots + (size_t)(c0 + (uint32_t)k * parts) * kLLMaxBins + b, v[k], g[k]);
                            pending = pending || g[k] != tag;
                        }
                    if (!pending) break;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += v[k];
            }
        }
        // combine the parts through shared memory (rows 1 .. parts of the private-copy area are free after the fold)
        uint32_t *psum = sh + (nbins + 1u);
        if (part < parts && part > 0) psum[(part - 1u) * nbins + b] = acc;
        __syncthreads();
        if (part == 0) {
            for (uint32_t q = 1; q < parts; ++q) acc += psum[(q - 1u) * nbins + b];
            sh[b] += acc;                                                                 // this GPU's total for bin b
        }
        __syncthreads();
        if (t == 0) {
            peer.stamps[1] = globaltimer_ns();                                           // every CTA's counts merged
            peer.counter[1] = 0;                                                         // the work-stealing claims of this launch
        }
    } else {
        // ---- publish + ticket (any nbins up to the shared-memory limit) ----
        // Every CTA adds its folded counts to the device-local partial (red.global), fences, and only THEN draws its ticket: the
        // CTA that draws the last one knows that every count -- its own included -- is in the partial.  Critical path after the
        // slowest CTA leaves the counting loop (b200_p2p_timeline, 2^27 elements): fold 0.2 us, reds + fence ~1.0, ticket ~0.7,
        // partial read-back 0.7.  (Drawing the ticket FIRST, so that the last CTA could keep its counts in shared memory, was
        // tried in r05 and is slower: the last CTA then waits for the publish chains of the CTAs just ahead of it, 4.0 us in all.)
        if (warp == 0) {
            for (uint32_t b = t; b < nbins; b += 32) {
                const uint32_t s = sh[b];
                if (s) atomicAdd(peer.partial + b, s);
            }
            __threadfence();                               // every lane: its reds are performed before the ticket below is visible
            __syncwarp();
            // (no fence behind the ticket: the last CTA reads the partial with ld.global.cg -- L2, where the reds were performed --
            // and those loads depend on the ticket's value through the branch below)
            if (t == 0) ticket = atomicAdd(peer.counter, 1u);
        }
        __syncthreads();
        if (ticket != gridDim.x - 1) return;   // only the last CTA of this GPU goes on
        if (t == 0) {
            peer.stamps[1] = globaltimer_ns();                                           // local counting merged (last ticket drawn)
            peer.counter[0] = 0;                                                         // ready for the next launch
            peer.counter[1] = 0;                                                         // (the work-stealing claims too)
            peer.stamps[6] = peer.stamps[1];
        }
        for (uint32_t b = t; b < nbins; b += T) {
            sh[b] = __ldcg(peer.partial + b);
            peer.partial[b] = 0;
        }
        if (t == 0) peer.stamps[7] = globaltimer_ns();                                   // partial read back
        // (each thread reads back only the sh[b] it wrote)
    }
    const size_t half = (size_t)(peer.epoch & 1u) * 8u * peer.nbins_max;
    for (uint32_t b = t; b < nbins; b += T) {
        const uint32_t v = sh[b];
        for (int r = 0; r < peer.world; ++r) st_slot(peer.slots[r] + half + (size_t)peer.rank * peer.nbins_max + b, v, tag);
    }
    if (t == 0) peer.stamps[2] = globaltimer_ns();                                       // slots sent
    // Poll: pair index q = b * 8 + src (8 consecutive lanes = the 8 possible sources of bin b).  A thread owns pairs
    // q = t, t + T, ... ; all of its loads are issued before any is checked.
    const unsigned long long *mine = peer.slots[peer.rank] + half;
    const unsigned long long t_start = globaltimer_ns();
    const int src = t & 7;
    const bool src_on = src < peer.world;
    bool gave_up = false;
    for (uint32_t b0 = 0; b0 < nbins; b0 += 4 * (T / 8)) {       // 4 bins per thread and pass
        uint32_t v[4], g[4];
        int32_t old[4];
        const unsigned long long *p[4];
        bool want[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t b = b0 + (uint32_t)k * (T / 8) + (uint32_t)(t >> 3);
            want[k] = src_on && b < nbins;
            p[k] = mine + (size_t)src * peer.nbins_max + (want[k] ? b : 0);
            v[k] = 0;
            g[k] = tag;
            // the value the sum is added to (`bins[b] +=`) is fetched NOW, under the slot poll, not after it
            old[k] = (src == 0 && b < nbins) ? bins[b] : 0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (want[k]) ld_slot(p[k], v[k], g[k]);
        uint32_t spins = 0;
        while (true) {
            bool pending = false;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (want[k] && g[k] != tag) {
                    ld_slot(p[k], v[k], g[k]);
                    pending = pending || g[k] != tag;
                }
            if (!pending) break;
            if (peer.timeout_ns && (++spins & 1023u) == 0 && globaltimer_ns() - t_start > peer.timeout_ns) {
                gave_up = true;                      // a peer never arrived (failed launch / validation on that rank)
                break;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t sum = (want[k] && g[k] == tag) ? v[k] : 0u;
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            const uint32_t b = b0 + (uint32_t)k * (T / 8) + (uint32_t)(t >> 3);
            if (src == 0 && b < nbins) bins[b] = old[k] + (int32_t)sum;                  // accumulate like the reference's +=
        }
    }
    if (gave_up) {
        *reinterpret_cast<volatile uint32_t *>(peer.error_flag) = tag;                   // bins are incomplete: the host reports it
        __threadfence_system();
    }
    __syncthreads();
    if (t == 0) peer.stamps[3] = globaltimer_ns();                                       // every rank's slots summed
}

template <int T, int P>
__global__ void __launch_bounds__(T) hist_allreduce_kernel(int32_t *__restrict__ bins, uint32_t nbins,
                                                           const int4 *__restrict__ in4, int64_t n4,
                                                           const int32_t *__restrict__ tail, int ntail, int copies,
                                                           const int32_t *__restrict__ head, int nhead, int interleave,
                                                           const __grid_constant__ HistPeer peer) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem_raw);
    if (blockIdx.x == 0 && threadIdx.x == 0) peer.stamps[5] = globaltimer_ns();          // first CTA resident
    // `interleave`: bit 0 = unit plan, bits 1-7 = rounds of the work-stealing tail (0: none), bits 8.. = rounds to prefetch
    if (interleave & 1)
        hist_count_phase<T, P, true>(sh, nbins, in4, n4, tail, ntail, copies, head, nhead, interleave >> 8,
                                     ((interleave >> 1) & 127) ? peer.counter + 1 : nullptr, (interleave >> 1) & 127);
    else
        hist_count_phase<T, P, false>(sh, nbins, in4, n4, tail, ntail, copies, head, nhead, interleave >> 8);
    fused_exchange<T>(sh, nbins, copies, bins, peer);
}

// nbins too large for shared memory: one global reduction per in-range element.
__global__ void hist_global_kernel(int32_t *__restrict__ bins, uint64_t nbins, const int32_t *__restrict__ in,
                                   int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = in[i];
        const uint64_t b = (uint64_t)((uint32_t)v - 1u);
        if (v >= 1 && b < nbins) atomicAdd(reinterpret_cast<unsigned int *>(bins) + b, 1u);
    }
}

// Launch with the programmatic-stream-serialization attribute (`hist.pdl` = 0 turns it off): the kernel may START before the previous
// kernel of the stream has finished IF that kernel executed griddepcontrol.launch_dependents (only this library's copy / fill / histogram
// kernels do); it touches global memory only after its own griddepcontrol.wait, which returns when the previous grid is complete.
template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tune_get("hist.pdl", 1) != 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
    return B200_OK;
}

template <int T, int P, int ILP>
static int launch_bytes(int32_t *bins, uint32_t nbins, const int4 *in4, int64_t n4, const int32_t *tail, int ntail,
                        int grid, cudaStream_t st) {
    size_t smem = (size_t)256 * T;
    B200_CUDA(cudaFuncSetAttribute(hist256_bytes_kernel<T, P, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    hist256_bytes_kernel<T, P, ILP><<<grid, T, smem, st>>>(bins, nbins, in4, n4, tail, ntail);
    B200_CHECK_LAUNCH("hist256_bytes_kernel");
    return B200_OK;
}

// Unit plan of the counting phase (`hist.interleave`: 0 slices, 1 interleaved, -1 = by size).  Measured on one B200
// (profiles/r05_hist_probe.txt): 2^30 elements 587.7 us sliced / 593.1 interleaved, 2^29 297.0 / 298.5, 2^28 151.2 / 151.6,
// 2^27 79.0 / 77.8 -- contiguous slices stream a little faster, interleaved units end closer together (4 us against 6 us between
// the first and the last CTA), which is what counts once a launch is short.
static int pick_interleave(int64_t n4) {
    const int forced = tune_get("hist.interleave", -1);
    const int npf = tune_get("hist.l2_prefetch_rounds", 16) << 8;   // rounds (16 KiB each) a CTA pulls into L2 before the dependency wait
    if (forced == 0 || forced == 1) return forced | npf;
    return (n4 * 16 < ((int64_t)768 << 20) ? 1 : 0) | npf;
}

template <int T, int P>
static int launch_atomic(int32_t *bins, uint32_t nbins, const int4 *in4, int64_t n4, const int32_t *tail, int ntail,
                         int grid, int copies, cudaStream_t st) {
    size_t smem = (size_t)(nbins + 1) * copies * 4;
    if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(hist_atomic_kernel<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    B200_TRY(launch_pdl(hist_atomic_kernel<T, P>, grid, T, smem, st, bins, nbins, in4, n4, tail, ntail, copies, pick_interleave(n4)));
    B200_CHECK_LAUNCH("hist_atomic_kernel");
    return B200_OK;
}

int histogram_launch(int32_t *bins, int64_t nbins, const int32_t *input, int64_t n, cudaStream_t st) {
    if (n == 0 || nbins == 0) return B200_OK;
    const int sms = sm_count();
    if (nbins > 14336) {
        int grid = sms * 8;
        hist_global_kernel<<<grid, 256, 0, st>>>(bins, (uint64_t)nbins, input, n);
        B200_CHECK_LAUNCH("hist_global_kernel");
        return B200_OK;
    }
    // split into [unaligned head | 16-byte aligned int4 body | tail]; head+tail (< 8 elements) go through `tail`
    // by launching the scalar pieces as a separate tiny range handled by CTA 0.
    int64_t head = 0;
    uintptr_t addr = (uintptr_t)input;
    if (addr & 15) head = (int64_t)((16 - (addr & 15)) / 4);
    if (head > n) head = n;
    const int32_t *body = input + head;
    int64_t n4 = (n - head) / 4;
    int64_t tail_n = (n - head) - n4 * 4;
    const int4 *in4 = reinterpret_cast<const int4 *>(body);
    const int32_t *tailp = body + n4 * 4;
    int impl = tune_get("hist.impl", 1);  // measured: 32-bit shared atomics beat the byte counters (profiles/)
    if (nbins > 256) impl = 1;
    if (head > 0) {
        // rare: unaligned input.  Count the (< 4) head elements with the global kernel first.
        hist_global_kernel<<<1, 32, 0, st>>>(bins, (uint64_t)nbins, input, head);
        B200_CHECK_LAUNCH("hist_global_kernel");
    }
    int grid = sms * tune_get("hist.ctas_per_sm", 1);
    if (impl == 2) {   // A/B: TMA-fed counting (see hist_tma_kernel)
        constexpr int T = 1024;
        const int stages = tune_get("hist.tma_stages", 4);
        int copies = (int)((48 * 1024) / ((nbins + 1) * 4));
        if (copies > T / 32) copies = T / 32;
        if (copies < 1) copies = 1;
        grid = sms;
        if ((int64_t)grid > (n4 + T - 1) / T) grid = (int)((n4 + T - 1) / T);
        if (grid < 1) grid = 1;
        const uint32_t ring_off = (uint32_t)((((size_t)(nbins + 1) * copies * 4) + 127) / 128 * 128);
        const size_t smem = ring_off + (size_t)stages * T * 16;
#define B200_HT(S)                                                                                                                 \
    case S:                                                                                                                        \
        B200_CUDA(cudaFuncSetAttribute(hist_tma_kernel<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        B200_TRY(launch_pdl(hist_tma_kernel<T, S>, grid, T, smem, st, bins, (uint32_t)nbins, in4, n4, tailp, (int)tail_n, copies, ring_off)); \
        break;
        switch (stages) {
            B200_HT(2)
            B200_HT(4)
            B200_HT(6)
            B200_HT(8)
            default: return set_error(B200_ERR_INVALID_VALUE, "hist.tma_stages must be 2, 4, 6 or 8");
        }
#undef B200_HT
        B200_CHECK_LAUNCH("hist_tma_kernel");
        return B200_OK;
    }
    if (impl == 0) {
        int T = tune_get("hist.threads", 512), P = tune_get("hist.batch", 8), ILP = tune_get("hist.ilp", 2);
        int64_t min_units = (int64_t)T;  // do not launch more CTAs than there is work
        if ((int64_t)grid > (n4 + min_units - 1) / min_units) grid = (int)((n4 + min_units - 1) / min_units);
        if (grid < 1) grid = 1;
#define B200_HB(TT, PP, II)                                                                                      \
    if (T == TT && P == PP && ILP == II)                                                                         \
        return launch_bytes<TT, PP, II>(bins, (uint32_t)nbins, in4, n4, tailp, (int)tail_n, grid, st);
        B200_HB(512, 8, 2)
        B200_HB(512, 8, 1)
        B200_HB(512, 4, 2)
        B200_HB(512, 4, 1)
        B200_HB(768, 4, 2)
        B200_HB(768, 4, 1)
        B200_HB(768, 8, 2)
        B200_HB(256, 8, 2)
        B200_HB(256, 8, 1)
#undef B200_HB
        return set_error(B200_ERR_INVALID_VALUE, "hist: unsupported (threads=%d, batch=%d, ilp=%d)", T, P, ILP);
    }
    {
        int T = tune_get("hist.atomic_threads", 1024), P = tune_get("hist.atomic_batch", 4);
        int warps = T / 32;
        int copies = (int)((48 * 1024) / ((nbins + 1) * 4));
        if (copies > warps) copies = warps;
        if (copies < 1) copies = 1;
        int cps = tune_get("hist.atomic_ctas_per_sm", 1);
        grid = sms * cps;
        if ((int64_t)grid > (n4 + T - 1) / T) grid = (int)((n4 + T - 1) / T);
        if (grid < 1) grid = 1;
#define B200_HA(TT, PP)                                                                                          \
    if (T == TT && P == PP)                                                                                      \
        return launch_atomic<TT, PP>(bins, (uint32_t)nbins, in4, n4, tailp, (int)tail_n, grid, copies, st);
        B200_HA(512, 4)
        B200_HA(512, 8)
        B200_HA(256, 4)
        B200_HA(256, 8)
        B200_HA(1024, 4)
        B200_HA(1024, 2)
        B200_HA(1024, 8)
        B200_HA(768, 4)
#undef B200_HA
        return set_error(B200_ERR_INVALID_VALUE, "hist: unsupported atomic config (threads=%d, batch=%d)", T, P);
    }
}

// ---------------------------------------------------------------------------------------------
// Int64 values and bins (Julia's default integer; histogram_kernel! is generic in T, examples/histogram.jl:18): the same
// design -- per-warp private 32-bit shared counters (a CTA sees < 2^32 elements), two values per 128-bit load, one
// 64-bit red.global.add per non-empty bin per CTA.  8 bytes/element read: HBM-bound at half the atomics per byte.
// ---------------------------------------------------------------------------------------------
template <int T, int P>
__global__ void __launch_bounds__(T) hist64_atomic_kernel(long long *__restrict__ bins, uint32_t nbins,
                                                          const int4 *__restrict__ in4, int64_t n4,
                                                          const long long *__restrict__ tail, int ntail, int copies, int plan) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *sh = reinterpret_cast<uint32_t *>(smem_raw);
    // the counting phase of the Int32 kernel with 64-bit values (two per 128-bit load): branch-free clamp into the sink counter,
    // unpredicated main loop, unit plan by size, L2 prefetch under the dependency wait (`plan` as in hist_atomic_kernel)
    if (plan & 1)
        hist_count_phase<T, P, true, true>(sh, nbins, in4, n4, reinterpret_cast<const int32_t *>(tail), ntail, copies, nullptr, 0, plan >> 8);
    else
        hist_count_phase<T, P, false, true>(sh, nbins, in4, n4, reinterpret_cast<const int32_t *>(tail), ntail, copies, nullptr, 0, plan >> 8);
    fold_copies<T>(sh, nbins, copies);
    for (uint32_t b = threadIdx.x; b < nbins; b += T) {
        const uint32_t sum = sh[b];
        if (sum) atomicAdd(reinterpret_cast<unsigned long long *>(bins) + b, (unsigned long long)sum);
    }
}
__global__ void hist64_global_kernel(long long *__restrict__ bins, uint64_t nbins, const long long *__restrict__ in, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = (uint64_t)in[i] - 1ull;
        if (b < nbins) atomicAdd(reinterpret_cast<unsigned long long *>(bins) + b, 1ull);
    }
}

int histogram64_launch(long long *bins, int64_t nbins, const long long *input, int64_t n, cudaStream_t st) {
    if (n == 0 || nbins == 0) return B200_OK;
    const int sms = sm_count();
    if (nbins > 14336) {
        hist64_global_kernel<<<sms * 8, 256, 0, st>>>(bins, (uint64_t)nbins, input, n);
        B200_CHECK_LAUNCH("hist64_global_kernel");
        return B200_OK;
    }
    // [8-byte aligned but not 16-byte aligned head element | int4 body of value pairs | odd tail element]
    int64_t head = ((uintptr_t)input & 15) ? 1 : 0;
    if (head > n) head = n;
    if (head) {
        hist64_global_kernel<<<1, 32, 0, st>>>(bins, (uint64_t)nbins, input, head);
        B200_CHECK_LAUNCH("hist64_global_kernel");
    }
    const long long *body = input + head;
    const int64_t n4 = (n - head) / 2, tail_n = (n - head) - 2 * n4;
    constexpr int T = 1024, P = 4;
    int copies = (int)((48 * 1024) / ((nbins + 1) * 4));
    if (copies > T / 32) copies = T / 32;
    if (copies < 1) copies = 1;
    int grid = sms;
    if ((int64_t)grid > (n4 + T - 1) / T) grid = (int)((n4 + T - 1) / T);
    if (grid < 1) grid = 1;
    const size_t smem = (size_t)(nbins + 1) * copies * 4;
    if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(hist64_atomic_kernel<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_TRY(launch_pdl(hist64_atomic_kernel<T, P>, grid, T, smem, st, bins, (uint32_t)nbins, reinterpret_cast<const int4 *>(body), n4,
                        body + 2 * n4, (int)tail_n, copies, pick_interleave(n4)));
    B200_CHECK_LAUNCH("hist64_atomic_kernel");
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// symmetric peer buffers (cudaIpc) for the fused histogram + all-reduce
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kP2PMaxBins = 14336;
constexpr size_t kP2PCounterOff = 0 /* [0] ticket, [1] work-stealing claims */, kP2PStampsOff = 8 /* 8 x u64 */, kP2PPartialOff = 128;
constexpr size_t kP2PSlotsOff = (kP2PPartialOff + (size_t)kP2PMaxBins * 4 + 255) / 256 * 256;
constexpr size_t kP2PCtaSlotsOff = kP2PSlotsOff + 2 * 8 * (size_t)kP2PMaxBins * 8;
constexpr size_t kP2PBytes = kP2PCtaSlotsOff + (size_t)kLLMaxCtas * kLLMaxBins * 8;

struct P2P {
    void *local = nullptr;
    int world = 0, rank = -1;
    void *peer[8] = {};
    uint32_t epoch = 0;
    uint32_t *error_host = nullptr;   // pinned, mapped: the kernel's "a peer never arrived" flag
    uint32_t *error_dev = nullptr;
    bool ipc = true;                  // peers mapped through cudaIpc (b200_p2p_connect) or plain pointers (b200_p2p_connect_ptrs)
};

int histogram_allreduce_launch(P2P *h, int32_t *bins, int64_t nbins, const int32_t *input, int64_t n, cudaStream_t st) {
    if (h->world < 1) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32_allreduce: b200_p2p_connect has not been called");
    if (nbins > kP2PMaxBins || nbins < 1)
        return set_error(B200_ERR_UNSUPPORTED, "b200_histogram_i32_allreduce: nbins must be 1..%u", kP2PMaxBins);
    if (((uintptr_t)input & 3) != 0)
        return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32_allreduce: input must be 4-byte aligned");
    if (h->error_host && *h->error_host != 0) {
        const uint32_t step = *h->error_host;
        return set_error(B200_ERR_CUDA, "b200_histogram_i32_allreduce: step %u timed out waiting for a peer's counts (a rank skipped the "
                         "collective or its launch failed); the communicator is out of step -- free it and connect again", step);
    }
    HistPeer peer;
    peer.world = h->world;
    peer.rank = h->rank;
    peer.epoch = h->epoch;           // advanced only after a successful launch: a rank whose launch fails stays in step
    peer.nbins_max = kP2PMaxBins;
    peer.counter = (uint32_t *)((char *)h->local + kP2PCounterOff);
    peer.stamps = (unsigned long long *)((char *)h->local + kP2PStampsOff);
    peer.partial = (uint32_t *)((char *)h->local + kP2PPartialOff);
    peer.error_flag = h->error_dev;
    // measured (profiles/r05_hist_probe2.txt, 2^27 elements, one rank): ticket + partial 82.2 us, per-CTA slot rows polled by CTA 0
    // 84.9 us -- 147 x 256 slots are five dependent rounds of loads for one CTA; opt-in (`hist.fused_ll = 1`)
    peer.cta_slots = tune_get("hist.fused_ll", 0) != 0 ? (unsigned long long *)((char *)h->local + kP2PCtaSlotsOff) : nullptr;
    peer.timeout_ns = (unsigned long long)tune_get("hist.fused_timeout_ms", 10000) * 1000000ull;
    for (int r = 0; r < 8; ++r)
        peer.slots[r] = r < h->world ? (unsigned long long *)((char *)h->peer[r] + kP2PSlotsOff) : nullptr;
    // [unaligned head (< 4 elements) | 16-byte aligned int4 body | tail (< 4 elements)]: head and tail are counted by CTA 0
    int64_t head = ((uintptr_t)input & 15) ? (int64_t)((16 - ((uintptr_t)input & 15)) / 4) : 0;
    if (head > n) head = n;
    const int32_t *body = input + head;
    const int64_t n4 = (n - head) / 4;
    const int tail_n = (int)((n - head) - n4 * 4);
    constexpr int T = 1024, P = 4;
    int copies = (int)((48 * 1024 - 16) / ((nbins + 1) * 4));
    if (copies > T / 32) copies = T / 32;
    if (copies < 1) copies = 1;
    int grid = sm_count();
    if ((int64_t)grid > (n4 + T - 1) / T) grid = (int)((n4 + T - 1) / T);
    if (grid < 1) grid = 1;   // n == 0 still takes part in the exchange
    const size_t smem = (size_t)(nbins + 1) * copies * 4 + 16;   // + the ticket word of fused_exchange
    if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(hist_allreduce_kernel<T, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_TRY(launch_pdl(hist_allreduce_kernel<T, P>, grid, T, smem, st, bins, (uint32_t)nbins, reinterpret_cast<const int4 *>(body), n4,
                        body + n4 * 4, tail_n, copies, input, (int)head,
                        pick_interleave(n4) | ((tune_get("hist.dyn_rounds", 0) & 127) << 1), peer));
    B200_CHECK_LAUNCH("hist_allreduce_kernel");
    h->epoch++;
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_p2p_alloc(void **handle) {
    if (!handle) return set_error(B200_ERR_INVALID_VALUE, "handle is NULL");
    B200_TRY(ensure_device());
    P2P *h = new P2P();
    cudaError_t e = cudaMalloc(&h->local, kP2PBytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        delete h;
        return set_error(B200_ERR_OOM, "b200_p2p_alloc: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    B200_CUDA(cudaMemset(h->local, 0, kP2PBytes));
    B200_CUDA(cudaMemset((char *)h->local + kP2PStampsOff + 4 * 8, 0xff, 8));   // stamp 4 is an atomicMin
    B200_CUDA(cudaHostAlloc((void **)&h->error_host, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    *h->error_host = 0;
    B200_CUDA(cudaHostGetDevicePointer((void **)&h->error_dev, h->error_host, 0));
    *handle = h;
    return B200_OK;
}

extern "C" b200_status b200_p2p_export(void *handle, void *ipc64) {
    if (!handle || !ipc64) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200_IPC_HANDLE_BYTES, "ipc handle size");
    cudaIpcMemHandle_t ipc;
    B200_CUDA(cudaIpcGetMemHandle(&ipc, ((P2P *)handle)->local));
    memcpy(ipc64, &ipc, sizeof(ipc));
    return B200_OK;
}

extern "C" b200_status b200_p2p_connect(void *handle, int32_t world, int32_t rank, const void *all_handles) {
    if (!handle || !all_handles) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (world < 1 || world > 8 || rank < 0 || rank >= world)
        return set_error(B200_ERR_INVALID_VALUE, "b200_p2p_connect: bad rank %d / world %d (1..8 ranks of one NVLink domain)", rank, world);
    B200_TRY(ensure_device());
    P2P *h = (P2P *)handle;
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
    h->epoch = 0;
    return B200_OK;
}

// Single-process form of b200_p2p_connect (the ranks are GPUs of one process): peers as pointers to their b200_p2p buffers.
extern "C" b200_status b200_p2p_connect_ptrs(void *handle, int32_t world, int32_t rank, void *const *peer_handles) {
    if (!handle || !peer_handles) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (world < 1 || world > 8 || rank < 0 || rank >= world)
        return set_error(B200_ERR_INVALID_VALUE, "b200_p2p_connect_ptrs: bad rank %d / world %d", rank, world);
    B200_TRY(ensure_device());
    P2P *h = (P2P *)handle;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            h->peer[r] = h->local;
            continue;
        }
        if (!peer_handles[r]) return set_error(B200_ERR_INVALID_VALUE, "b200_p2p_connect_ptrs: NULL handle for rank %d", r);
        void *buf = ((P2P *)peer_handles[r])->local;
        cudaPointerAttributes at;
        B200_CUDA(cudaPointerGetAttributes(&at, buf));
        int cur = -1;
        B200_CUDA(cudaGetDevice(&cur));
        if (at.device != cur) {   // (a "peer" on the same GPU needs no enabling: used by the single-GPU tests)
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else B200_CUDA(e);
        }
        h->peer[r] = buf;
    }
    h->world = world;
    h->rank = rank;
    h->epoch = 0;
    h->ipc = false;
    return B200_OK;
}

// Diagnostic: four %globaltimer stamps (ns) of the last fused launch: slowest CTA left the counting loop, local counting
// merged (ticket), slots sent to all ranks, all ranks' slots summed.  Synchronises the calling thread's stream.
extern "C" b200_status b200_p2p_debug(void *handle, uint64_t *stamps4) {
    uint64_t *stamps6 = stamps4;
    if (!handle || !stamps6) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    B200_CUDA(cudaStreamSynchronize(st));
    B200_CUDA(cudaMemcpy(stamps6, (char *)((P2P *)handle)->local + kP2PStampsOff, 32, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemset((char *)((P2P *)handle)->local + kP2PStampsOff, 0, 8));   // stamp 0 is an atomicMax
    B200_CUDA(cudaMemset((char *)((P2P *)handle)->local + kP2PStampsOff + 4 * 8, 0xff, 8));   // stamp 4 is an atomicMin
    return B200_OK;
}

// Eight stamps: [0..3] as b200_p2p_debug, [4] fastest CTA left the counting loop, [5] CTA 0 became resident, [6] every other CTA
// has published its counts (last CTA), [7] the partial has been read back.
extern "C" b200_status b200_p2p_timeline(void *handle, uint64_t *stamps6) {
    if (!handle || !stamps6) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    B200_CUDA(cudaStreamSynchronize(st));
    B200_CUDA(cudaMemcpy(stamps6, (char *)((P2P *)handle)->local + kP2PStampsOff, 64, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemset((char *)((P2P *)handle)->local + kP2PStampsOff, 0, 8));
    B200_CUDA(cudaMemset((char *)((P2P *)handle)->local + kP2PStampsOff + 4 * 8, 0xff, 8));
    return B200_OK;
}

// Profiling aid (tools/nvlink_proof.py): store `counts` as source rank `src`'s contribution to the NEXT fused step into this
// rank's own slot array, exactly as that peer's kernel would -- so that ONE rank of a multi-rank histogram can be replayed under
// ncu (which serialises kernels: ranks polling each other would deadlock) while its slot stores still cross NVLink.
extern "C" b200_status b200_p2p_debug_prefill(void *handle, int32_t src, const uint32_t *counts_host, int64_t nbins) {
    if (!handle || !counts_host) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    P2P *h = (P2P *)handle;
    if (src < 0 || src >= h->world || nbins < 1 || nbins > (int64_t)kP2PMaxBins) return set_error(B200_ERR_INVALID_VALUE, "b200_p2p_debug_prefill: bad src / nbins");
    B200_TRY(ensure_device());
    std::vector<unsigned long long> words((size_t)nbins);
    const uint32_t tag = h->epoch + 1u;
    for (int64_t b = 0; b < nbins; ++b) words[(size_t)b] = (unsigned long long)counts_host[b] | ((unsigned long long)tag << 32);
    const size_t half = (size_t)(h->epoch & 1u) * 8u * kP2PMaxBins;
    B200_CUDA(cudaMemcpy((char *)h->local + kP2PSlotsOff + (half + (size_t)src * kP2PMaxBins) * 8, words.data(), (size_t)nbins * 8,
                         cudaMemcpyHostToDevice));
    return B200_OK;
}

extern "C" b200_status b200_p2p_free(void *handle) {
    if (!handle) return B200_OK;
    P2P *h = (P2P *)handle;
    B200_CUDA(cudaDeviceSynchronize());
    for (int r = 0; r < h->world; ++r)
        if (h->ipc && r != h->rank && h->peer[r]) (void)cudaIpcCloseMemHandle(h->peer[r]);
    if (h->local) (void)cudaFree(h->local);
    if (h->error_host) (void)cudaFreeHost(h->error_host);
    delete h;
    return B200_OK;
}

extern "C" b200_status b200_histogram_i32_allreduce(void *handle, int32_t *bins, int64_t nbins, const int32_t *input,
                                                    int64_t n) {
    if (!handle) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32_allreduce: handle is NULL");
    if (nbins < 0 || n < 0) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32_allreduce: negative size");
    if (!bins || (n > 0 && !input)) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32_allreduce: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return histogram_allreduce_launch((P2P *)handle, bins, nbins, input, n, st);
}

extern "C" b200_status b200_histogram_i32(int32_t *bins, int64_t nbins, const int32_t *input, int64_t n) {
    if (nbins < 0 || n < 0) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32: negative size");
    if (n == 0 || nbins == 0) return B200_OK;
    if (!bins || !input) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32: NULL pointer");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return histogram_launch(bins, nbins, input, n, st);
}

extern "C" b200_status b200_histogram_i64(int64_t *bins, int64_t nbins, const int64_t *input, int64_t n) {
    if (nbins < 0 || n < 0) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i64: negative size");
    if (n == 0 || nbins == 0) return B200_OK;
    if (!bins || !input) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i64: NULL pointer");
    if (((uintptr_t)input & 7) != 0) return set_error(B200_ERR_UNSUPPORTED, "b200_histogram_i64: input must be 8-byte aligned");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    return histogram64_launch((long long *)bins, nbins, (const long long *)input, n, st);
}
