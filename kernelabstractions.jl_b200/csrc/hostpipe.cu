// hostpipe.cu -- host-buffer ("out-of-core") forms of the north-star kernels: the arrays live in HOST memory and the
// library streams slabs of them through the GPU, overlapping the host->device copy of slab s+1, the kernel of slab s
// and the device->host copy of slab s-1 on three streams.  PCIe is full duplex, so the end-to-end time of a transpose
// approaches max(h2d, d2h) instead of h2d + kernel + d2h.
//
// Role in the reference: what a user of the CPU() backend gets by calling naive_transpose!(a, b) / histogram!(...)
// on plain Arrays (examples/naive_transpose.jl:9-21, examples/histogram.jl:61-72): host data in, host data out.
// Pattern for the overlap: examples/mpi.jl:24-39 (stream-ordered copyto! + cooperative wait), pagelock!
// (src/KernelAbstractions.jl:166).  Host buffers should be page-locked (b200_host_alloc / b200_host_register);
// pageable memory still works but the copies then serialise inside the driver.
//
// Ordering: the pipeline starts after everything already enqueued on the calling thread's stream and that stream
// waits for the pipeline's last copy, so KA.synchronize(backend) covers it like any other stream-ordered call.
#include "common.cuh"
#include "tune.h"

namespace b200 {

int transpose_launch(void *a, const void *b, int64_t m, int64_t n, int64_t lda, int64_t ldb, int elsize, cudaStream_t st);
int histogram_launch(int32_t *bins, int64_t nbins, const int32_t *input, int64_t n, cudaStream_t st);
int sgemm_launch(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N, int64_t lda, int64_t ldb, int64_t ldc, int mode,
                 cudaStream_t st);

namespace {

constexpr int NBUF = 3;

struct HostPipe {
    bool ready = false;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t start = nullptr, in_done[NBUF] = {}, k_done[NBUF] = {}, out_done[NBUF] = {};
    void *ws = nullptr;
    size_t ws_bytes = 0;
};

thread_local HostPipe g_pipe[kMaxDevices];

int get_pipe(HostPipe **out, size_t ws_bytes) {
    B200_TRY(ensure_device());
    HostPipe &p = g_pipe[tls().device];
    if (!p.ready) {
        B200_CUDA(cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&p.s_k, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming));
        for (int i = 0; i < NBUF; ++i) {
            B200_CUDA(cudaEventCreateWithFlags(&p.in_done[i], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&p.k_done[i], cudaEventDisableTiming));
            B200_CUDA(cudaEventCreateWithFlags(&p.out_done[i], cudaEventDisableTiming));
        }
        p.ready = true;
    }
    if (p.ws_bytes < ws_bytes) {
        // the previous workspace may still be in use by an earlier pipelined call on these streams
        if (p.ws) {
            B200_CUDA(cudaStreamSynchronize(p.s_in));
            B200_CUDA(cudaStreamSynchronize(p.s_k));
            B200_CUDA(cudaStreamSynchronize(p.s_out));
            B200_CUDA(cudaFree(p.ws));
            p.ws = nullptr;
            p.ws_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&p.ws, ws_bytes);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return set_error(B200_ERR_OOM, "host pipeline: cannot allocate %zu bytes of device workspace", ws_bytes);
        }
        p.ws_bytes = ws_bytes;
    }
    *out = &p;
    return B200_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace
}  // namespace b200

using namespace b200;

// a (m x n, lda) = transpose of b (n x m, ldb), both in host memory.  Slab s covers columns [j0, j1) of a, i.e. rows
// [j0, j1) of b: a strided (pitch ldb) host->device copy into a dense w x m slab, one transpose, and a contiguous
// device->host copy of the dense m x w result.
extern "C" b200_status b200_transpose_host(void *a_host, const void *b_host, int64_t m, int64_t n, int64_t lda,
                                           int64_t ldb, int32_t elsize) {
    if (m < 0 || n < 0 || lda < m || ldb < n)
        return set_error(B200_ERR_INVALID_VALUE, "b200_transpose_host: bad shape m=%lld n=%lld lda=%lld ldb=%lld",
                         (long long)m, (long long)n, (long long)lda, (long long)ldb);
    if (elsize != 1 && elsize != 2 && elsize != 4 && elsize != 8 && elsize != 16)
        return set_error(B200_ERR_UNSUPPORTED, "b200_transpose_host: element size %d not in {1,2,4,8,16}", elsize);
    if (m == 0 || n == 0) return B200_OK;
    if (!a_host || !b_host) return set_error(B200_ERR_INVALID_VALUE, "b200_transpose_host: NULL pointer");
    cudaStream_t user;
    B200_TRY(current_stream(&user));

    // slab width: ~hostpipe.slab_mib MiB of output per slab, a multiple of 64 columns (fast transpose path)
    const int64_t target = (int64_t)tune_get("hostpipe.slab_mib", 32) << 20;
    int64_t w = target / (m * elsize);
    w = w / 64 * 64;
    if (w < 64) w = 64;
    if (w > n) w = n;
    const int64_t nslabs = cdiv(n, w);
    const size_t slab_bytes = align_up((size_t)w * (size_t)m * (size_t)elsize, 256);

    HostPipe *p = nullptr;
    B200_TRY(get_pipe(&p, 2 * NBUF * slab_bytes));
    char *in_buf[NBUF], *out_buf[NBUF];
    for (int i = 0; i < NBUF; ++i) {
        in_buf[i] = (char *)p->ws + (size_t)i * slab_bytes;
        out_buf[i] = (char *)p->ws + (size_t)(NBUF + i) * slab_bytes;
    }
    B200_CUDA(cudaEventRecord(p->start, user));
    B200_CUDA(cudaStreamWaitEvent(p->s_in, p->start, 0));
    B200_CUDA(cudaStreamWaitEvent(p->s_k, p->start, 0));
    B200_CUDA(cudaStreamWaitEvent(p->s_out, p->start, 0));

    for (int64_t s = 0; s < nslabs; ++s) {
        const int i = (int)(s % NBUF);
        const int64_t j0 = s * w, wj = (j0 + w <= n) ? w : n - j0;
        // h2d: rows [j0, j0+wj) of b -> dense wj x m slab (ld = wj)
        if (s >= NBUF) B200_CUDA(cudaStreamWaitEvent(p->s_in, p->k_done[i], 0));       // input buffer free again
        B200_CUDA(cudaMemcpy2DAsync(in_buf[i], (size_t)wj * elsize, (const char *)b_host + (size_t)j0 * elsize,
                                    (size_t)ldb * elsize, (size_t)wj * elsize, (size_t)m, cudaMemcpyHostToDevice, p->s_in));
        B200_CUDA(cudaEventRecord(p->in_done[i], p->s_in));
        // kernel: a_slab (m x wj, ld m) = transpose(b_slab)
        B200_CUDA(cudaStreamWaitEvent(p->s_k, p->in_done[i], 0));
        if (s >= NBUF) B200_CUDA(cudaStreamWaitEvent(p->s_k, p->out_done[i], 0));     // output buffer free again
        B200_TRY(transpose_launch(out_buf[i], in_buf[i], m, wj, m, wj, elsize, p->s_k));
        B200_CUDA(cudaEventRecord(p->k_done[i], p->s_k));
        // d2h: dense m x wj slab -> columns [j0, j0+wj) of a
        B200_CUDA(cudaStreamWaitEvent(p->s_out, p->k_done[i], 0));
        char *dst = (char *)a_host + (size_t)j0 * (size_t)lda * elsize;
        if (lda == m)
            B200_CUDA(cudaMemcpyAsync(dst, out_buf[i], (size_t)m * wj * elsize, cudaMemcpyDeviceToHost, p->s_out));
        else
            B200_CUDA(cudaMemcpy2DAsync(dst, (size_t)lda * elsize, out_buf[i], (size_t)m * elsize, (size_t)m * elsize,
                                        (size_t)wj, cudaMemcpyDeviceToHost, p->s_out));
        B200_CUDA(cudaEventRecord(p->out_done[i], p->s_out));
    }
    // the caller's stream resumes after the last device->host copy (s_out is in order, so the last event covers all)
    B200_CUDA(cudaStreamWaitEvent(user, p->out_done[(nslabs - 1) % NBUF], 0));
    return B200_OK;
}

// bins_host[v-1] += #{t : input_host[t] == v}: the input is streamed through two device slabs while the previous
// slab is being counted; the (tiny) bins make one round trip.
extern "C" b200_status b200_histogram_i32_host(int32_t *bins_host, int64_t nbins, const int32_t *input_host,
                                               int64_t n) {
    if (nbins < 0 || n < 0) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32_host: negative size");
    if (nbins == 0 || n == 0) return B200_OK;
    if (!bins_host || !input_host) return set_error(B200_ERR_INVALID_VALUE, "b200_histogram_i32_host: NULL pointer");
    cudaStream_t user;
    B200_TRY(current_stream(&user));
    const int64_t slab_elems = ((int64_t)tune_get("hostpipe.slab_mib", 32) << 20) / 4;
    const int64_t nslabs = cdiv(n, slab_elems);
    const size_t slab_bytes = (size_t)slab_elems * 4;
    const size_t bins_bytes = align_up((size_t)nbins * 4, 256);
    HostPipe *p = nullptr;
    B200_TRY(get_pipe(&p, NBUF * slab_bytes + bins_bytes));
    int32_t *d_bins = (int32_t *)((char *)p->ws + NBUF * slab_bytes);

    B200_CUDA(cudaEventRecord(p->start, user));
    B200_CUDA(cudaStreamWaitEvent(p->s_in, p->start, 0));
    B200_CUDA(cudaStreamWaitEvent(p->s_k, p->start, 0));
    B200_CUDA(cudaMemcpyAsync(d_bins, bins_host, (size_t)nbins * 4, cudaMemcpyHostToDevice, p->s_k));
    for (int64_t s = 0; s < nslabs; ++s) {
        const int i = (int)(s % NBUF);
        const int64_t t0 = s * slab_elems, cnt = (t0 + slab_elems <= n) ? slab_elems : n - t0;
        int32_t *d_in = (int32_t *)((char *)p->ws + (size_t)i * slab_bytes);
        if (s >= NBUF) B200_CUDA(cudaStreamWaitEvent(p->s_in, p->k_done[i], 0));
        B200_CUDA(cudaMemcpyAsync(d_in, input_host + t0, (size_t)cnt * 4, cudaMemcpyHostToDevice, p->s_in));
        B200_CUDA(cudaEventRecord(p->in_done[i], p->s_in));
        B200_CUDA(cudaStreamWaitEvent(p->s_k, p->in_done[i], 0));
        B200_TRY(histogram_launch(d_bins, nbins, d_in, cnt, p->s_k));
        B200_CUDA(cudaEventRecord(p->k_done[i], p->s_k));
    }
    B200_CUDA(cudaMemcpyAsync(bins_host, d_bins, (size_t)nbins * 4, cudaMemcpyDeviceToHost, p->s_k));
    B200_CUDA(cudaEventRecord(p->out_done[0], p->s_k));
    B200_CUDA(cudaStreamWaitEvent(user, p->out_done[0], 0));
    return B200_OK;
}

// C (M x N, ldc) = A (M x K, lda) * B (K x N, ldb), all in host memory, FP32, summation order `mode` of b200_matmul_f32
// (coalesced_matmul_kernel!, examples/performant_matmul.jl:79-92, on plain host Arrays).  B goes to the device once; A is streamed
// as row panels (pitched h2d into a dense panel), each panel's product is computed while the next panel arrives and the
// previous C panel leaves (pitched d2h).  k is never split, so every bit equals the one-shot device product.
extern "C" b200_status b200_matmul_f32_host(float *C_host, const float *A_host, const float *B_host, int64_t M, int64_t K, int64_t N,
                                            int64_t lda, int64_t ldb, int64_t ldc, int32_t mode) {
    if (M < 0 || N < 0 || K < 0 || lda < M || ldb < K || ldc < M) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_host: bad shape");
    if (M == 0 || N == 0) return B200_OK;
    if (!C_host || (K > 0 && (!A_host || !B_host))) return set_error(B200_ERR_INVALID_VALUE, "b200_matmul_f32_host: NULL pointer");
    cudaStream_t user;
    B200_TRY(current_stream(&user));
    // panel height: a multiple of 128 rows (the GEMM's tile), sized so that a panel's A and C pieces are ~hostpipe.slab_mib MiB
    const int64_t target = (int64_t)tune_get("hostpipe.slab_mib", 32) << 20;
    int64_t mp = target / (4 * (K > N ? K : N));
    mp = mp / 128 * 128;
    if (mp < 128) mp = 128;
    if (mp > M) mp = M;
    const int64_t npanels = cdiv(M, mp);
    const int64_t kpad = (K + 3) / 4 * 4;                                  // device B with ld % 4 == 0 keeps the TMA engine eligible
    const size_t b_bytes = align_up((size_t)kpad * (size_t)N * 4, 256);
    const size_t a_bytes = align_up((size_t)mp * (size_t)(K ? K : 1) * 4, 256), c_bytes = align_up((size_t)mp * (size_t)N * 4, 256);
    HostPipe *p = nullptr;
    B200_TRY(get_pipe(&p, b_bytes + NBUF * (a_bytes + c_bytes)));
    float *dB = (float *)p->ws;
    char *abuf = (char *)p->ws + b_bytes, *cbuf = abuf + NBUF * a_bytes;

    B200_CUDA(cudaEventRecord(p->start, user));
    B200_CUDA(cudaStreamWaitEvent(p->s_in, p->start, 0));
    B200_CUDA(cudaStreamWaitEvent(p->s_k, p->start, 0));
    B200_CUDA(cudaStreamWaitEvent(p->s_out, p->start, 0));
    if (K > 0) {
        if (kpad != K) B200_CUDA(cudaMemsetAsync(dB, 0, (size_t)kpad * (size_t)N * 4, p->s_in));   // padding rows never enter a product
        B200_CUDA(cudaMemcpy2DAsync(dB, (size_t)kpad * 4, B_host, (size_t)ldb * 4, (size_t)K * 4, (size_t)N, cudaMemcpyHostToDevice, p->s_in));
    }
    for (int64_t s = 0; s < npanels; ++s) {
        const int i = (int)(s % NBUF);
        const int64_t m0 = s * mp, mh = (m0 + mp <= M) ? mp : M - m0;
        float *dA = (float *)(abuf + (size_t)i * a_bytes), *dC = (float *)(cbuf + (size_t)i * c_bytes);
        if (s >= NBUF) B200_CUDA(cudaStreamWaitEvent(p->s_in, p->k_done[i], 0));
        if (K > 0)   // rows [m0, m0+mh) of A: K columns of mh floats, host pitch lda -> dense panel (ld = mp keeps 16-byte aligned columns)
            B200_CUDA(cudaMemcpy2DAsync(dA, (size_t)mp * 4, A_host + m0, (size_t)lda * 4, (size_t)mh * 4, (size_t)K, cudaMemcpyHostToDevice, p->s_in));
        B200_CUDA(cudaEventRecord(p->in_done[i], p->s_in));
        B200_CUDA(cudaStreamWaitEvent(p->s_k, p->in_done[i], 0));
        if (s >= NBUF) B200_CUDA(cudaStreamWaitEvent(p->s_k, p->out_done[i], 0));
        B200_TRY(sgemm_launch(dC, dA, dB, mh, K, N, mp, kpad, mp, mode, p->s_k));
        B200_CUDA(cudaEventRecord(p->k_done[i], p->s_k));
        B200_CUDA(cudaStreamWaitEvent(p->s_out, p->k_done[i], 0));
        B200_CUDA(cudaMemcpy2DAsync(C_host + m0, (size_t)ldc * 4, dC, (size_t)mp * 4, (size_t)mh * 4, (size_t)N, cudaMemcpyDeviceToHost, p->s_out));
        B200_CUDA(cudaEventRecord(p->out_done[i], p->s_out));
    }
    B200_CUDA(cudaStreamWaitEvent(user, p->out_done[(npanels - 1) % NBUF], 0));
    return B200_OK;
}
