/*
 * b200ka.h -- C ABI of libb200ka.so, the native runtime behind `B200Backend <: KA.GPU`.
 *
 * This header is the drop-in boundary for the KernelAbstractions.jl hot path.  It plays the
 * role that the 23 `@ccall libpocl.POcl*` symbols play for the in-tree POCL backend
 * (reference: src/pocl/nanoOpenCL.jl:492-682): plain `extern "C"`, raw pointers and sizes,
 * every function returns an int32 status (0 == OK, negative == error) exactly like the
 * `cl_int` returned by the OpenCL entry points that `@checked` wraps
 * (src/pocl/nanoOpenCL.jl:22-52, CLError at :392-408).
 *
 * Conventions
 *   - arrays are column-major, dims[0] is the fastest dimension (Julia layout,
 *     reference src/pocl/device/array.jl:167-189);
 *   - all indices crossing this ABI are 0-based (the host layers translate Julia's 1-based ids);
 *   - all work is enqueued on the calling thread's current stream (one non-blocking stream per
 *     (host thread, device), see b200_get_stream); nothing blocks unless documented;
 *   - the library never frees or retains caller memory.
 *
 * There is NO CPU fallback anywhere in this library: without a CUDA device every compute
 * entry point returns B200_ERR_NO_DEVICE.
 */
#ifndef B200KA_H
#define B200KA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200KA_VERSION 100 /* 0.1.0 */
#define B200_MAX_NDIM 4

/* ---- status codes -------------------------------------------------------------------------
 * Mapping to the exception types the reference testsuite asserts:
 *   B200_ERR_PARTITION       -> ErrorException (src/KernelAbstractions.jl:768,773,780; test/test.jl:29,41)
 *   B200_ERR_LAUNCH_DIMS     -> ArgumentError  (src/intrinsics.jl:182-188; test/intrinsics.jl:79-80)
 *   B200_ERR_BAD_DEVICE      -> ArgumentError  (src/KernelAbstractions.jl:686-691; test/devices.jl:10-11)
 *   B200_ERR_LENGTH_MISMATCH -> error("Arrays must match in length") (src/pocl/backend.jl:42-44)
 *   B200_ERR_ALIAS           -> error("Arrays may not alias")        (src/pocl/backend.jl:45-47)
 *   B200_ERR_CUDA            -> error("Kernel execution failed")     (src/pocl/nanoOpenCL.jl:1423-1432)
 */
typedef int32_t b200_status;
enum {
    B200_OK = 0,
    B200_NOT_READY = 1,            /* b200_stream_query: work still in flight */
    B200_ERR_CUDA = -1,            /* a CUDA runtime call failed; see b200_last_error() */
    B200_ERR_INVALID_VALUE = -2,
    B200_ERR_UNKNOWN_KERNEL = -3,  /* b200_launch: no precompiled kernel of that name/arity */
    B200_ERR_BAD_ARGS = -4,        /* kernel exists but argument kinds/eltypes/shapes do not fit */
    B200_ERR_LAUNCH_DIMS = -5,     /* more than 3 launch dims, or workgroup > 1024 items */
    B200_ERR_LENGTH_MISMATCH = -6,
    B200_ERR_ALIAS = -7,
    B200_ERR_UNSUPPORTED = -8,
    B200_ERR_NO_DEVICE = -9,
    B200_ERR_OOM = -10,
    B200_ERR_PARTITION = -11,
    B200_ERR_BAD_DEVICE = -12,
    B200_ERR_NCCL = -13
};

/* Thread-local, human readable description of the last non-OK status on this thread. */
const char *b200_last_error(void);
const char *b200_status_name(int32_t status);
int32_t b200_version(void);

/* ---- device management (KA.ndevices/device/device!, src/KernelAbstractions.jl:670-691;
 *      KI.multiprocessor_count / KI.max_work_group_size, src/pocl/backend.jl:174-179) -------- */
b200_status b200_init(void);
b200_status b200_device_count(int32_t *count);
b200_status b200_set_device(int32_t device);       /* 0-based; out of range -> B200_ERR_BAD_DEVICE */
b200_status b200_get_device(int32_t *device);
b200_status b200_sm_count(int32_t *count);
b200_status b200_max_threads_per_block(int32_t *count);
b200_status b200_mem_info(size_t *free_bytes, size_t *total_bytes);
b200_status b200_device_name(char *buf, size_t buflen);

/* ---- memory (KA.allocate / unsafe_free! / pagelock!, src/pocl/backend.jl:25,57;
 *      src/KernelAbstractions.jl:166,193-195,574-583) ---------------------------------------- */
b200_status b200_malloc(void **ptr, size_t bytes, int32_t unified); /* bytes==0 -> *ptr = NULL, OK */
b200_status b200_free(void *ptr);
b200_status b200_host_alloc(void **ptr, size_t bytes);   /* pinned host memory */
b200_status b200_host_free(void *ptr);
b200_status b200_host_register(void *ptr, size_t bytes); /* KA.pagelock! */
b200_status b200_host_unregister(void *ptr);

/* KA.copyto! legs that cross the host/device boundary (src/pocl/backend.jl:52 -> Base.copyto!).
 * Asynchronous and stream ordered: the caller keeps `src` alive until b200_synchronize
 * (contract: src/KernelAbstractions.jl:118-151). */
b200_status b200_memcpy_h2d_async(void *dst, const void *src, size_t bytes);
b200_status b200_memcpy_d2h_async(void *dst, const void *src, size_t bytes);
b200_status b200_memcpy_d2d_async(void *dst, const void *src, size_t bytes);
/* pitched device-to-device copy of `height` rows of `width_bytes` (column-major sub-matrix views: one row = one column) */
b200_status b200_memcpy_2d_d2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t height);
b200_status b200_memset_async(void *dst, int32_t byte, size_t bytes);

/* ---- streams / synchronisation (KA.synchronize, KA.priority!; src/pocl/backend.jl:67,
 *      src/KernelAbstractions.jl:176,658-663; docs/src/implementations.md:3-11) --------------- */
b200_status b200_stream_query(void);  /* B200_OK when idle, B200_NOT_READY otherwise: yield-loop for Julia */
b200_status b200_synchronize(void);   /* blocking wait on the calling thread's stream */
b200_status b200_device_synchronize(void);
b200_status b200_set_priority(int32_t priority); /* -1 :low, 0 :normal, 1 :high */
b200_status b200_get_stream(void **cuda_stream);
b200_status b200_set_stream(void *cuda_stream);  /* adopt an external cudaStream_t (NULL = back to own) */
/* Explicit stream handles for hosts whose unit of execution is a TASK, not an OS thread.  The reference keeps device and
 * queue in task_local_storage (src/pocl/pocl.jl:12-41) and Julia tasks migrate between OS threads at every yield point
 * (KA.synchronize itself yields), so the Julia layer owns one stream per (task, device, priority) and calls b200_bind
 * before every other entry point: set the calling OS thread's device and adopt the stream in one call (NULL = the
 * thread's own stream).  b200_stream_create makes a non-blocking stream on the calling thread's current device. */
b200_status b200_stream_create(void **cuda_stream, int32_t priority);
b200_status b200_stream_destroy(void *cuda_stream);
b200_status b200_bind(int32_t device, void *cuda_stream);

/* CUDA events on the calling thread's stream (device-side timing for bench.py). */
b200_status b200_event_create(void **event);
b200_status b200_event_record(void *event);
b200_status b200_event_synchronize(void *event);
b200_status b200_event_elapsed_ms(void *start, void *stop, float *ms);
b200_status b200_event_destroy(void *event);

/* CUDA graphs for launch-bound sequences (the reference waits after EVERY launch, src/pocl/backend.jl:144-145; here even the
 * launch overhead of a recorded sequence collapses into one submission).  b200_graph_begin starts recording on the calling
 * thread's stream; every stream-ordered call made by that thread until b200_graph_end is recorded instead of executed.
 * Calls that wait for the stream (b200_synchronize, b200_reduce, synchronous copies ...) invalidate the recording.
 * b200_graph_launch replays on the calling thread's stream (same argument pointers and values as recorded). */
b200_status b200_graph_begin(void);
b200_status b200_graph_end(void **graph);
b200_status b200_graph_launch(void *graph);
b200_status b200_graph_destroy(void *graph);

/* ---- generic kernel launch ----------------------------------------------------------------
 * Replaces  clSetKernelArg* + clEnqueueNDRangeKernel  (src/pocl/nanoOpenCL.jl:1161-1283) for
 *   (obj::KA.Kernel{B})(args...; ndrange, workgroupsize)      src/pocl/backend.jl:117-147  (kind 0)
 *   (obj::KI.Kernel{B})(args...; numworkgroups, workgroupsize) src/pocl/backend.jl:156-168  (kind 1)
 * There is no JIT: `name` selects one of the precompiled sm_100a kernels (b200_kernel_name).
 * b200_launch_desc is the flattened CompilerMetadata/NDRange (src/compiler.jl:1-33,
 * src/nditeration.jl:49-60) that the host obtains from the unchanged KA.partition. */
enum { B200_LAUNCH_KA = 0, B200_LAUNCH_KI = 1 };

typedef struct b200_launch_desc {
    int32_t kind;                    /* B200_LAUNCH_KA or B200_LAUNCH_KI */
    int32_t ndim;                    /* KA: N of the ndrange (1..4); KI: number of given dims (1..3) */
    int64_t ndrange[B200_MAX_NDIM];  /* KA only: size(ndrange) */
    int64_t wgs[B200_MAX_NDIM];      /* work-items per group, padded with 1 */
    int64_t blocks[B200_MAX_NDIM];   /* groups per dimension, padded with 1 */
    int32_t dynamic_check;           /* KA: __validindex does the `I in ndrange` test */
    int32_t flags;                   /* bit0 static ndrange, bit1 static workgroupsize (informational) */
} b200_launch_desc;

enum { /* b200_arg.kind */
    B200_ARG_ARRAY = 0,   /* device array: ptr, eltype, elsize, ndim, dims */
    B200_ARG_SCALAR = 1,  /* by-value bits in .scalar (i64) / .fscalar (f64), eltype says which */
    B200_ARG_VAL = 2      /* Julia Val{N}() / type argument: value in .scalar */
};

enum { /* b200_arg.eltype */
    B200_T_OPAQUE = 0,
    B200_T_I8 = 1, B200_T_U8 = 2, B200_T_I16 = 3, B200_T_U16 = 4,
    B200_T_I32 = 5, B200_T_U32 = 6, B200_T_I64 = 7, B200_T_U64 = 8,
    B200_T_F16 = 9, B200_T_BF16 = 10, B200_T_F32 = 11, B200_T_F64 = 12, B200_T_BOOL = 13
};

typedef struct b200_arg {
    int32_t kind;
    int32_t eltype;
    int32_t elsize;                 /* bytes per element */
    int32_t ndim;
    int64_t dims[B200_MAX_NDIM];
    void *ptr;
    int64_t scalar;
    double fscalar;
} b200_arg;

b200_status b200_launch(const char *name, const b200_launch_desc *desc, const b200_arg *args, int32_t nargs);
/* Run-time registration of precompiled kernels that live in the caller's own shared library -- the no-JIT counterpart of
 * KI.kernel_function's compilation cache (reference src/pocl/backend.jl:151-154, src/pocl/compiler/execution.jl:190-217).
 * `fn` receives the launch exactly as b200_launch got it, plus the calling thread's cudaStream_t (as void*) on which it must
 * enqueue its kernel(s); it returns B200_OK or an error status (B200_ERR_BAD_ARGS for arguments it does not accept).
 * include/b200ka_plugin.cuh holds the device-side index helpers (@index, __validindex) and the grid computation.
 * Built-in names cannot be overridden.  Thread safe. */
typedef b200_status (*b200_kernel_handler)(const b200_launch_desc *desc, const b200_arg *args, int32_t nargs,
                                            void *cuda_stream, void *user);
b200_status b200_register_kernel(const char *name, int32_t kind, b200_kernel_handler fn, void *user);
b200_status b200_unregister_kernel(const char *name);
b200_status b200_kernel_count(int32_t *count);
const char *b200_kernel_name(int32_t index); /* NULL when out of range */
/* KI.kernel_max_work_group_size (src/pocl/backend.jl:170-173; test/intrinsics.jl:95-96) */
b200_status b200_kernel_max_work_group_size(const char *name, int64_t max_work_items, int64_t *out);
/* number of device kernels this thread has launched since the last reset (bench.py gpu_launches) */
b200_status b200_launch_counter(int64_t *count, int32_t reset);

/* Tuning knobs (kernel variant / grid sizing sweeps; defaults are the measured best). Not part of the
 * reference-facing surface. */
b200_status b200_tune_set(const char *key, int32_t value);
b200_status b200_tune_get(const char *key, int32_t dflt, int32_t *value);

/* ---- typed fast paths (what b200_launch routes the north-star kernels to) ------------------- */

/* copy_kernel / copy_kernel! : A[I] = B[I]  (src/KernelAbstractions.jl:874-877,
 * examples/memcopy.jl:4-7).  Byte-exact, dtype agnostic. Overlap -> B200_ERR_ALIAS. */
b200_status b200_copy(void *dst, const void *src, size_t bytes);
/* init_kernel: arr[I] = f(T) (src/KernelAbstractions.jl:869-872); value = elsize bytes, elsize in {1,2,4,8,16}. */
b200_status b200_fill(void *dst, size_t count, int32_t elsize, const void *value);

/* naive_transpose_kernel!: a[i,j] = b[j,i]  (examples/naive_transpose.jl:4-7).
 * a is m x n (leading dim lda >= m), b is n x m (ldb >= n), column-major, elsize in {1,2,4,8,16} (4: the tuned tile kernel). */
b200_status b200_transpose(void *a, const void *b, int64_t m, int64_t n, int64_t lda, int64_t ldb,
                           int32_t elsize);

/* histogram_kernel!: bins[v-1] += #{t : input[t] == v}, v in 1..nbins (values outside are ignored)
 * (examples/histogram.jl:18-58).  Accumulates into `bins` like the reference's atomic +=. */
b200_status b200_histogram_i32(int32_t *bins, int64_t nbins, const int32_t *input, int64_t n);
/* the same for Int64 values and bins (Julia's default integer; the reference kernel is generic in the element type) */
b200_status b200_histogram_i64(int64_t *bins, int64_t nbins, const int64_t *input, int64_t n);

/* saxpy_kernel! / saxpy_kernel: Z[I] = a * X[I] + Y[I], product and sum rounded separately like the reference's
 * unfused Julia expression (benchmark/benchmarks.jl:29-32; examples/numa_aware.jl:10-13 is the in-place form Z == Y).
 * n elements of Float32 (is_f64 == 0) or Float64; Z may alias Y but not X (-> B200_ERR_ALIAS). */
b200_status b200_saxpy(void *Z, double a, const void *X, const void *Y, size_t n, int32_t is_f64);

/* matmul_kernel! / coalesced_matmul_kernel!: C(MxN) = A(MxK) * B(KxN), column-major
 * (examples/matmul.jl:5-15, examples/performant_matmul.jl:15-77).
 * mode 0: FP32 SIMT, summation order of the reference tiled kernel (16-wide k partials, k ascending,
 *         fused multiply-add) -> bit-identical to the oracle;
 * mode 1: FP32 SIMT, single running accumulator per output (k ascending). */
enum { B200_MATMUL_TILE16 = 0, B200_MATMUL_RUNNING = 1,
       B200_MATMUL_UNFUSED = 2 /* matmul_kernel! (examples/matmul.jl:5-15): tmp += a*b, product and sum rounded separately */ };
b200_status b200_matmul_f32(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N,
                            int64_t lda, int64_t ldb, int64_t ldc, int32_t mode);

/* Opt-in tensor-core paths (tcgen05 + TMEM + TMA).  No reference implementation exists for these
 * ("parity unpinned"): defined as the FP32 oracle on BF16/TF32-rounded inputs, rtol 1e-2.
 * Abf: M x K bf16 with K contiguous (row-major, ld = K); Bbf: N x K bf16 with K contiguous
 * (i.e. the column-major K x N matrix B as stored); C: M x N float32 column-major, ldc >= M. */
b200_status b200_matmul_bf16(float *C, const uint16_t *Abf, const uint16_t *Bbf, int64_t M, int64_t K,
                             int64_t N, int64_t ldc);
/* TF32 tensor-core path: C = A*B straight from the FP32 column-major matrices (no conversion pass, no workspace);
 * the tensor core uses the top 19 bits of every operand word, FP32 accumulate.  lda, ldb multiples of 4, A and B
 * 16-byte aligned; any M, N, K > 0 (edge tiles are zero-filled by TMA and stores are predicated). */
b200_status b200_matmul_tf32(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N,
                             int64_t lda, int64_t ldb, int64_t ldc);
/* "3xTF32": FP32-grade C = A*B on the tensor cores.  Every operand is split into its TF32 head and the TF32 head of the
 * remainder; A_lo*B_hi + A_hi*B_lo + A_hi*B_hi are accumulated per k step, K in chunks whose TMEM sums are added to C in FP32
 * (the tensor core's own accumulation truncates).  Same operand rules as b200_matmul_tf32; workspace: (lda*K + ldb*N) floats,
 * 16-byte aligned.  Not the reference's summation order: parity is "rtol 1e-5 against Float64" (measured 2e-6), not bit-exact.
 * Finite operands only: an Inf operand makes its row / column of C non-finite (Inf, or NaN where it meets a zero remainder). */
b200_status b200_matmul_f32_tf32x3(float *C, const float *A, const float *B, int64_t M, int64_t K, int64_t N,
                                   int64_t lda, int64_t ldb, int64_t ldc, void *workspace);
/* float32 column-major (rows x cols, ld) -> bf16; transpose!=0 writes the cols x rows... see DESIGN.md:
 * out is K-contiguous storage for the GEMM above: for A (M x K col-major) use transpose=1,
 * for B (K x N col-major) use transpose=0. Round-to-nearest-even. */
b200_status b200_convert_f32_bf16(uint16_t *out, const float *in, int64_t rows, int64_t cols, int64_t ld,
                                  int32_t transpose);
/* C = A*B with FP32 storage, BF16 tensor-core arithmetic: converts then calls b200_matmul_bf16.
 * `workspace` must hold (M*K + K*N) * 2 bytes. */
b200_status b200_matmul_f32_via_bf16(float *C, const float *A, const float *B, int64_t M, int64_t K,
                                     int64_t N, int64_t lda, int64_t ldb, int64_t ldc, void *workspace);

/* The same with B converted once and kept resident (Bbf = b200_convert_f32_bf16(B, K, N, ldb, transpose 0)): the sharded form --
 * B is replicated at set-up, only the rank's A row panel is converted per call.  `workspace` must hold M*K * 2 bytes.
 * Same bits as b200_matmul_f32_via_bf16. */
b200_status b200_matmul_f32_bf16_resident_b(float *C, const float *A, const uint16_t *Bbf, int64_t M, int64_t K,
                                            int64_t N, int64_t lda, int64_t ldc, void *workspace);

/* ---- device-array conveniences: what Base / LinearAlgebra give a CPU() user for free on plain Arrays ----------------
 * (SURVEY.md 8(f)-4).  These return a VALUE, so they wait for the calling thread's stream.  Any element-aligned pointer
 * (unaligned views take a scalar prologue; two operands vectorise when they are equally misaligned).
 *
 * maximum / minimum / sum over n elements (examples/histogram.jl:10,88-90 sizes the histogram with `maximum(input)`).
 * op: 0 = sum, 1 = maximum, 2 = minimum.  eltype: B200_T_I32/U32/I64/U64/F32/F64.  *result_host receives a value of the
 * element type for maximum/minimum (NaN propagates, max(-0.0, 0.0) == 0.0 as in Julia) and an 8-byte Int64 / UInt64 /
 * Float64 for sum.  n == 0: sum is 0, maximum/minimum are B200_ERR_INVALID_VALUE (Julia throws ArgumentError). */
b200_status b200_reduce(int32_t op, int32_t eltype, const void *x, size_t n, void *result_host);
/* `A == B` (examples/memcopy.jl:23, naive_transpose.jl:32, test/copyto.jl:17-19): number of positions with !(a == b);
 * F32/F64 compare numerically (NaN != NaN, -0.0 == 0.0), every other eltype compares elsize bytes per element. */
b200_status b200_count_mismatch(const void *a, const void *b, size_t n, int32_t eltype, int32_t elsize, int64_t *count_host);
/* isapprox(A, B) (examples/matmul.jl:36, performant_matmul.jl:92): out3_host = { sum (x-y)^2, sum x^2, sum y^2 } in Float64;
 * the host layer applies  sqrt(out[0]) <= max(atol, rtol * max(sqrt(out[1]), sqrt(out[2]))). */
b200_status b200_norms(const void *x, const void *y, size_t n, int32_t is_f64, double *out3_host);
/* Deterministic input streams of SURVEY.md 8(d) generated on the device, bit-identical to the oracle's generators:
 * uniform Float32 in [0,1) with 24 random bits, and Int32 values uniform in 1..nbins (examples/histogram.jl:76). */
b200_status b200_rand_uniform_f32(float *out, size_t n, uint64_t seed);
b200_status b200_rand_bins_i32(int32_t *out, size_t n, uint64_t seed, int32_t nbins);

/* ---- host-buffer ("out-of-core") forms ---------------------------------------------------------
 * Same semantics as b200_transpose / b200_histogram_i32, but every array lives in HOST memory: what a user of
 * the reference's CPU() backend calls on plain Arrays (examples/naive_transpose.jl:9-21, examples/histogram.jl:61-72).
 * The library streams slabs through the GPU with the h2d copy, the kernel and the d2h copy of consecutive slabs
 * overlapped on three streams (pattern: examples/mpi.jl:24-39; pagelock!: src/KernelAbstractions.jl:166).
 * Stream ordered like every other call: results are valid after b200_synchronize().  Host buffers should be
 * page-locked (b200_host_alloc / b200_host_register). */
b200_status b200_transpose_host(void *a_host, const void *b_host, int64_t m, int64_t n, int64_t lda, int64_t ldb,
                                int32_t elsize);
b200_status b200_histogram_i32_host(int32_t *bins_host, int64_t nbins, const int32_t *input_host, int64_t n);
/* C = A * B on HOST arrays (FP32, column-major, `mode` as b200_matmul_f32): B is uploaded once, A row panels stream in and
 * C row panels stream out while the panels in between are multiplied.  Same bits as the device-resident product. */
b200_status b200_matmul_f32_host(float *C_host, const float *A_host, const float *B_host, int64_t M, int64_t K, int64_t N,
                                 int64_t lda, int64_t ldb, int64_t ldc, int32_t mode);

/* ---- multi-GPU plumbing (pattern: examples/mpi.jl:24-39; NCCL over NVLink) ------------------ */
#define B200_UNIQUE_ID_BYTES 128
b200_status b200_comm_unique_id(void *id128);
b200_status b200_comm_init_rank(void **comm, int32_t nranks, int32_t rank, const void *id128);
b200_status b200_comm_destroy(void *comm);
b200_status b200_allreduce_sum_i32(void *comm, int32_t *buf, int64_t count);
b200_status b200_broadcast(void *comm, void *buf, size_t bytes, int32_t root);
/* exchange! (examples/mpi.jl:24-39): send to dst_rank, receive from src_rank, device buffers, stream ordered. */
b200_status b200_sendrecv(void *comm, const void *sendbuf, int32_t dst_rank, void *recvbuf, int32_t src_rank,
                          size_t bytes);


/* ---- symmetric device buffers + distributed transpose (fused tile transpose + all-to-all over NVLink peer memory) -----------
 * b200_sym_alloc: `bytes` of device memory on the calling rank; export the 64-byte IPC handle, all-gather the handles with the
 * launcher's own transport, b200_sym_connect maps every peer's buffer (<= 8 ranks of one NVLink domain, one process per GPU).
 * b200_sym_ptr(handle, r): rank r's buffer as addressable from this rank (r == -1: the own buffer).
 * b200_sym_barrier: stream-ordered barrier across the ranks of the buffer (system fence, one flag store per peer, spin) --
 * everything this rank wrote to any peer before it is visible to a peer that has passed it.  Every rank must call it the
 * same number of times.
 * b200_transpose_alltoall: a = b^T for a GLOBAL N0 x N1 column-major matrix sharded by columns: the caller holds
 * b_local = b[:, rank-th block of N1/world columns] (N0 x N1/world, ld N0); on return of the stream the symmetric buffer of
 * every rank h holds a[:, h-th block of N0/world columns] (N1 x N0/world, ld N1).  One data kernel per rank whose stores go
 * straight into the destination slabs, then b200_sym_barrier.  N0, N1 multiples of 64 x world; elsize 4 or 8.
 * Back-to-back calls are safe without a barrier in between: every rank announces at the start of its call that its slab
 * may be overwritten (its stream has finished whatever read the previous result) and stores into a slab wait for that.
 * Every wait for a peer inside these kernels is bounded (`a2a.timeout_ms` tuning knob, default 10 s): a rank that never makes
 * the matching call makes its peers' NEXT call on the buffer fail with B200_ERR_CUDA instead of hanging their GPUs.
 * (the all-to-all the reference-style formulation would need, examples/mpi.jl:24-39 being its only exchange.) */
b200_status b200_sym_alloc(void **handle, size_t bytes);
b200_status b200_sym_export(void *handle, void *ipc64);
b200_status b200_sym_connect(void *handle, int32_t world, int32_t rank, const void *all_handles /* world x 64 bytes */);
/* single-process form: the ranks are GPUs of one process, peers given as pointers (b200_sym_ptr(peer, -1)); enables peer access */
b200_status b200_sym_connect_ptrs(void *handle, int32_t world, int32_t rank, void *const *peer_bufs);
b200_status b200_sym_ptr(void *handle, int32_t rank, void **ptr);
b200_status b200_sym_barrier(void *handle);
b200_status b200_sym_free(void *handle);
b200_status b200_transpose_alltoall(void *a_sym, const void *b_local, int64_t N0, int64_t N1, int32_t elsize);

/* ---- fused compute + collective over NVLink peer memory ------------------------------------------------------
 * One rank (process) per GPU of one NVLink/NVSwitch domain.  Each rank allocates a small symmetric buffer, the
 * launcher exchanges the 64-byte IPC handles (all-gather), every rank maps its peers' buffers, and from then on
 *   b200_histogram_i32_allreduce == b200_histogram_i32 on the rank's input slab + all-reduce(sum) of the bins
 * in ONE kernel: the last CTA stores the rank's partial bins as {count, tag} words straight into every peer's slot
 * array over NVLink and sums the slots the peers stored into its own (csrc/kernels/histogram.cu).  bins[b] += total count over ALL ranks, on
 * every rank -- bit-identical to b200_histogram_i32 + b200_allreduce_sum_i32 (integer adds).  Collective: every
 * rank must call it the same number of times.  nbins <= 14336, input 16-byte aligned, at most 8 ranks. */
#define B200_IPC_HANDLE_BYTES 64
b200_status b200_p2p_alloc(void **handle);
b200_status b200_p2p_export(void *handle, void *ipc64);
b200_status b200_p2p_connect(void *handle, int32_t world, int32_t rank, const void *all_handles /* world x 64 bytes */);
/* single-process form of b200_p2p_connect: peers given as their b200_p2p handles (created while their GPU was current) */
b200_status b200_p2p_connect_ptrs(void *handle, int32_t world, int32_t rank, void *const *peer_handles);
/* profiling aid: store source rank `src`'s contribution to the NEXT fused step into this rank's slots (tools/nvlink_proof.py) */
b200_status b200_p2p_debug_prefill(void *handle, int32_t src, const uint32_t *counts_host, int64_t nbins);
b200_status b200_p2p_free(void *handle);
/* diagnostic: four %globaltimer stamps (ns) of the last fused launch (slowest CTA left the counting loop, local counts
 * merged, slots sent, all ranks' slots summed) */
b200_status b200_p2p_debug(void *handle, uint64_t *stamps4);
/* the same plus [4] fastest CTA left the counting loop, [5] first CTA resident, [6] every other CTA has published its counts,
 * [7] the merged partial has been read back (eight stamps) */
b200_status b200_p2p_timeline(void *handle, uint64_t *stamps8);
b200_status b200_histogram_i32_allreduce(void *handle, int32_t *bins, int64_t nbins, const int32_t *input, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* B200KA_H */
