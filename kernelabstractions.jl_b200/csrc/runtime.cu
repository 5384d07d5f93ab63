// runtime.cu -- device / memory / stream / event entry points of the C ABI (include/b200ka.h).
//
// Role in the reference: what src/pocl/nanoOpenCL.jl obtains from libpocl (context, queue, SVM
// allocation, enqueue + wait: :723-1140,1228-1283,1423-1432) and what src/pocl/backend.jl:25-69
// builds on top of it.  Here the "queue" is a per-thread CUDA stream and nothing waits unless asked:
// KA's contract for a GPU backend is asynchronous, stream-ordered execution
// (reference src/KernelAbstractions.jl:118-151,176; docs/src/implementations.md:3-11).
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace b200 {

static thread_local ThreadState g_tls;
ThreadState &tls() { return g_tls; }

int set_error(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_tls.err, sizeof(g_tls.err), fmt, ap);
    va_end(ap);
    return status;
}

static int g_device_count = -1;
static int g_sm_count[kMaxDevices];
static std::once_flag g_once;

static void probe_devices() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
    }
    if (n > kMaxDevices) n = kMaxDevices;
    for (int d = 0; d < n; ++d) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d);
        g_sm_count[d] = sms;
    }
    g_device_count = n;
}

int ensure_device() {
    std::call_once(g_once, probe_devices);
    if (g_device_count <= 0)
        return set_error(B200_ERR_NO_DEVICE, "no CUDA device available (libb200ka has no CPU fallback)");
    ThreadState &t = tls();
    if (!t.inited) {
        // A fresh host thread starts on whatever device the CUDA runtime reports for it (device 0 unless
        // something else in the process, e.g. torch.cuda.set_device, already bound it).
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
        t.device = cur < g_device_count ? cur : 0;
        t.inited = true;
    }
    B200_CUDA(cudaSetDevice(t.device));
    return B200_OK;
}

int current_stream(cudaStream_t *s) {
    B200_TRY(ensure_device());
    ThreadState &t = tls();
    if (t.use_external) {
        *s = t.external;
        return B200_OK;
    }
    cudaStream_t &slot = t.own[t.device][t.priority + 1];
    if (!slot) {
        int lo = 0, hi = 0;  // numerically lower == higher priority
        B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        int prio = t.priority > 0 ? hi : (t.priority < 0 ? lo : (lo + hi) / 2);
        B200_CUDA(cudaStreamCreateWithPriority(&slot, cudaStreamNonBlocking, prio));
    }
    *s = slot;
    return B200_OK;
}

int sm_count() {
    if (ensure_device() != B200_OK) return 0;
    return g_sm_count[tls().device];
}

}  // namespace b200

using namespace b200;

extern "C" {

const char *b200_last_error(void) { return tls().err; }

const char *b200_status_name(int32_t s) {
    switch (s) {
        case B200_OK: return "B200_OK";
        case B200_NOT_READY: return "B200_NOT_READY";
        case B200_ERR_CUDA: return "B200_ERR_CUDA";
        case B200_ERR_INVALID_VALUE: return "B200_ERR_INVALID_VALUE";
        case B200_ERR_UNKNOWN_KERNEL: return "B200_ERR_UNKNOWN_KERNEL";
        case B200_ERR_BAD_ARGS: return "B200_ERR_BAD_ARGS";
        case B200_ERR_LAUNCH_DIMS: return "B200_ERR_LAUNCH_DIMS";
        case B200_ERR_LENGTH_MISMATCH: return "B200_ERR_LENGTH_MISMATCH";
        case B200_ERR_ALIAS: return "B200_ERR_ALIAS";
        case B200_ERR_UNSUPPORTED: return "B200_ERR_UNSUPPORTED";
        case B200_ERR_NO_DEVICE: return "B200_ERR_NO_DEVICE";
        case B200_ERR_OOM: return "B200_ERR_OOM";
        case B200_ERR_PARTITION: return "B200_ERR_PARTITION";
        case B200_ERR_BAD_DEVICE: return "B200_ERR_BAD_DEVICE";
        case B200_ERR_NCCL: return "B200_ERR_NCCL";
        default: return "B200_ERR_?";
    }
}

int32_t b200_version(void) { return B200KA_VERSION; }

b200_status b200_init(void) {
    B200_TRY(ensure_device());
    cudaStream_t s;
    return current_stream(&s);
}

b200_status b200_device_count(int32_t *count) {
    if (!count) return set_error(B200_ERR_INVALID_VALUE, "count is NULL");
    std::call_once(g_once, probe_devices);
    *count = g_device_count < 0 ? 0 : g_device_count;
    return B200_OK;
}

b200_status b200_set_device(int32_t device) {
    std::call_once(g_once, probe_devices);
    if (g_device_count <= 0) return set_error(B200_ERR_NO_DEVICE, "no CUDA device available");
    if (device < 0 || device >= g_device_count)
        return set_error(B200_ERR_BAD_DEVICE, "device %d out of range 0..%d", device, g_device_count - 1);
    ThreadState &t = tls();
    t.device = device;
    t.inited = true;
    B200_CUDA(cudaSetDevice(device));
    return B200_OK;
}

b200_status b200_get_device(int32_t *device) {
    if (!device) return set_error(B200_ERR_INVALID_VALUE, "device is NULL");
    B200_TRY(ensure_device());
    *device = tls().device;
    return B200_OK;
}

b200_status b200_sm_count(int32_t *count) {
    if (!count) return set_error(B200_ERR_INVALID_VALUE, "count is NULL");
    B200_TRY(ensure_device());
    *count = g_sm_count[tls().device];
    return B200_OK;
}

b200_status b200_max_threads_per_block(int32_t *count) {
    if (!count) return set_error(B200_ERR_INVALID_VALUE, "count is NULL");
    B200_TRY(ensure_device());
    int v = 0;
    B200_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxThreadsPerBlock, tls().device));
    *count = v;
    return B200_OK;
}

b200_status b200_mem_info(size_t *free_bytes, size_t *total_bytes) {
    B200_TRY(ensure_device());
    size_t f = 0, t = 0;
    B200_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return B200_OK;
}

b200_status b200_device_name(char *buf, size_t buflen) {
    if (!buf || buflen == 0) return set_error(B200_ERR_INVALID_VALUE, "buf is NULL");
    B200_TRY(ensure_device());
    cudaDeviceProp p;
    B200_CUDA(cudaGetDeviceProperties(&p, tls().device));
    snprintf(buf, buflen, "%s", p.name);
    return B200_OK;
}

// ---- memory -------------------------------------------------------------------------------
b200_status b200_malloc(void **ptr, size_t bytes, int32_t unified) {
    if (!ptr) return set_error(B200_ERR_INVALID_VALUE, "ptr is NULL");
    *ptr = nullptr;
    B200_TRY(ensure_device());
    if (bytes == 0) return B200_OK;  // zero-size arrays are legal (reference test/test.jl:290-298)
    cudaError_t e = unified ? cudaMallocManaged(ptr, bytes) : cudaMalloc(ptr, bytes);
    if (e == cudaErrorMemoryAllocation) {
        (void)cudaGetLastError();
        return set_error(B200_ERR_OOM, "out of device memory allocating %zu bytes", bytes);
    }
    B200_CUDA(e);
    return B200_OK;
}

b200_status b200_free(void *ptr) {
    if (!ptr) return B200_OK;
    B200_TRY(ensure_device());
    B200_CUDA(cudaFree(ptr));
    return B200_OK;
}

b200_status b200_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return set_error(B200_ERR_INVALID_VALUE, "ptr is NULL");
    *ptr = nullptr;
    B200_TRY(ensure_device());
    if (bytes == 0) return B200_OK;
    B200_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return B200_OK;
}

b200_status b200_host_free(void *ptr) {
    if (!ptr) return B200_OK;
    B200_CUDA(cudaFreeHost(ptr));
    return B200_OK;
}

b200_status b200_host_register(void *ptr, size_t bytes) {
    if (!ptr || bytes == 0) return B200_OK;
    B200_TRY(ensure_device());
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        (void)cudaGetLastError();
        return B200_OK;
    }
    B200_CUDA(e);
    return B200_OK;
}

b200_status b200_host_unregister(void *ptr) {
    if (!ptr) return B200_OK;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        (void)cudaGetLastError();
        return B200_OK;
    }
    B200_CUDA(e);
    return B200_OK;
}

static int memcpy_async(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
    if (bytes == 0) return B200_OK;
    if (!dst || !src) return set_error(B200_ERR_INVALID_VALUE, "NULL pointer in memcpy of %zu bytes", bytes);
    cudaStream_t s;
    B200_TRY(current_stream(&s));
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, s));
    return B200_OK;
}

// strided (pitched) device-to-device copy: `height` rows of `width_bytes`, row pitches in bytes -- sub-matrix views
b200_status b200_memcpy_2d_d2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t height) {
    if (width_bytes == 0 || height == 0) return B200_OK;
    if (!dst || !src) return set_error(B200_ERR_INVALID_VALUE, "b200_memcpy_2d_d2d_async: NULL pointer");
    if (dpitch < width_bytes || spitch < width_bytes) return set_error(B200_ERR_INVALID_VALUE, "b200_memcpy_2d_d2d_async: pitch < width");
    cudaStream_t s;
    B200_TRY(current_stream(&s));
    B200_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, cudaMemcpyDeviceToDevice, s));
    return B200_OK;
}

b200_status b200_memcpy_h2d_async(void *dst, const void *src, size_t bytes) {
    return memcpy_async(dst, src, bytes, cudaMemcpyHostToDevice);
}
b200_status b200_memcpy_d2h_async(void *dst, const void *src, size_t bytes) {
    return memcpy_async(dst, src, bytes, cudaMemcpyDeviceToHost);
}
b200_status b200_memcpy_d2d_async(void *dst, const void *src, size_t bytes) {
    return memcpy_async(dst, src, bytes, cudaMemcpyDeviceToDevice);
}

b200_status b200_memset_async(void *dst, int32_t byte, size_t bytes) {
    if (bytes == 0) return B200_OK;
    if (!dst) return set_error(B200_ERR_INVALID_VALUE, "NULL pointer in memset");
    cudaStream_t s;
    B200_TRY(current_stream(&s));
    B200_CUDA(cudaMemsetAsync(dst, byte, bytes, s));
    return B200_OK;
}

// ---- streams ------------------------------------------------------------------------------
b200_status b200_stream_query(void) {
    cudaStream_t s;
    B200_TRY(current_stream(&s));
    cudaError_t e = cudaStreamQuery(s);
    if (e == cudaSuccess) return B200_OK;
    if (e == cudaErrorNotReady) {
        (void)cudaGetLastError();
        return B200_NOT_READY;
    }
    B200_CUDA(e);
    return B200_OK;
}

b200_status b200_synchronize(void) {
    cudaStream_t s;
    B200_TRY(current_stream(&s));
    B200_CUDA(cudaStreamSynchronize(s));
    return B200_OK;
}

b200_status b200_device_synchronize(void) {
    B200_TRY(ensure_device());
    B200_CUDA(cudaDeviceSynchronize());
    return B200_OK;
}

b200_status b200_set_priority(int32_t priority) {
    if (priority < -1 || priority > 1)
        return set_error(B200_ERR_INVALID_VALUE, "priority must be -1 (:low), 0 (:normal) or 1 (:high)");
    tls().priority = priority;
    return B200_OK;
}

b200_status b200_get_stream(void **cuda_stream) {
    if (!cuda_stream) return set_error(B200_ERR_INVALID_VALUE, "cuda_stream is NULL");
    cudaStream_t s;
    B200_TRY(current_stream(&s));
    *cuda_stream = (void *)s;
    return B200_OK;
}

b200_status b200_set_stream(void *cuda_stream) {
    ThreadState &t = tls();
    t.external = (cudaStream_t)cuda_stream;
    t.use_external = cuda_stream != nullptr;
    return B200_OK;
}

// Explicit stream handles for hosts whose unit of execution is not the OS thread (Julia tasks migrate between threads at
// every yield point; the reference keeps its queue in task_local_storage, src/pocl/pocl.jl:12-41): the host owns the
// handles, keeps (device, stream) per task and binds them to the OS thread it is on before each call.
b200_status b200_stream_create(void **cuda_stream, int32_t priority) {
    if (!cuda_stream) return set_error(B200_ERR_INVALID_VALUE, "cuda_stream is NULL");
    *cuda_stream = nullptr;
    if (priority < -1 || priority > 1)
        return set_error(B200_ERR_INVALID_VALUE, "priority must be -1 (:low), 0 (:normal) or 1 (:high)");
    B200_TRY(ensure_device());
    int lo = 0, hi = 0;  // numerically lower == higher priority
    B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int prio = priority > 0 ? hi : (priority < 0 ? lo : (lo + hi) / 2);
    cudaStream_t s = nullptr;
    B200_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio));
    *cuda_stream = (void *)s;
    return B200_OK;
}

b200_status b200_stream_destroy(void *cuda_stream) {
    if (!cuda_stream) return B200_OK;
    ThreadState &t = tls();
    if (t.use_external && t.external == (cudaStream_t)cuda_stream) {
        t.external = nullptr;
        t.use_external = false;
    }
    B200_CUDA(cudaStreamDestroy((cudaStream_t)cuda_stream));   // returns at once; resources go when the stream drains
    return B200_OK;
}

b200_status b200_bind(int32_t device, void *cuda_stream) {
    B200_TRY(b200_set_device(device));
    return b200_set_stream(cuda_stream);
}

// ---- events -------------------------------------------------------------------------------
b200_status b200_event_create(void **event) {
    if (!event) return set_error(B200_ERR_INVALID_VALUE, "event is NULL");
    B200_TRY(ensure_device());
    cudaEvent_t e;
    B200_CUDA(cudaEventCreate(&e));
    *event = (void *)e;
    return B200_OK;
}

b200_status b200_event_record(void *event) {
    cudaStream_t s;
    B200_TRY(current_stream(&s));
    B200_CUDA(cudaEventRecord((cudaEvent_t)event, s));
    return B200_OK;
}

b200_status b200_event_synchronize(void *event) {
    B200_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return B200_OK;
}

b200_status b200_event_elapsed_ms(void *start, void *stop, float *ms) {
    if (!ms) return set_error(B200_ERR_INVALID_VALUE, "ms is NULL");
    B200_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return B200_OK;
}

b200_status b200_event_destroy(void *event) {
    if (!event) return B200_OK;
    B200_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return B200_OK;
}

// ---- CUDA graphs: record a launch-bound sequence of stream-ordered calls once, replay it with one launch -----------------
// Every b200 call that only ENQUEUES work (kernel launches, async copies, fills, events) may be recorded; calls that wait for
// the stream (b200_synchronize, value-returning reductions, synchronous copies) fail while a capture is open.
namespace {
struct GraphRec {           // what a b200 graph handle points to
    cudaGraphExec_t exec;
    int64_t launches;       // device kernels recorded (added to the thread's launch counter on every replay)
};
thread_local int64_t g_capture_base = 0;
}  // namespace

b200_status b200_graph_begin(void) {
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    B200_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    g_capture_base = tls().launches;
    return B200_OK;
}

b200_status b200_graph_end(void **graph) {
    if (!graph) return set_error(B200_ERR_INVALID_VALUE, "b200_graph_end: graph is NULL");
    *graph = nullptr;
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &g);
    const int64_t recorded = tls().launches - g_capture_base;
    tls().launches = g_capture_base;      // recorded launches did not run
    if (e != cudaSuccess || !g) {
        cudaGetLastError();
        return set_error(B200_ERR_CUDA, "b200_graph_end: capture failed or was invalidated by a synchronising call: %s", cudaGetErrorString(e));
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return set_error(B200_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    *graph = (void *)new GraphRec{exec, recorded};
    return B200_OK;
}

b200_status b200_graph_launch(void *graph) {
    if (!graph) return set_error(B200_ERR_INVALID_VALUE, "b200_graph_launch: NULL graph");
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    GraphRec *r = (GraphRec *)graph;
    B200_CUDA(cudaGraphLaunch(r->exec, st));
    count_launch((int)r->launches);
    return B200_OK;
}

b200_status b200_graph_destroy(void *graph) {
    if (!graph) return B200_OK;
    GraphRec *r = (GraphRec *)graph;
    cudaError_t e = cudaGraphExecDestroy(r->exec);
    delete r;
    B200_CUDA(e);
    return B200_OK;
}

b200_status b200_launch_counter(int64_t *count, int32_t reset) {
    if (count) *count = tls().launches;
    if (reset) tls().launches = 0;
    return B200_OK;
}

}  // extern "C"
