// common.cuh -- shared plumbing of libb200ka.so: status/error handling, per-thread runtime state.
//
// Per-thread state mirrors the reference's task-local platform/device/context/queue
// (reference src/pocl/pocl.jl:12-41): every host thread has a current device and, per device,
// one non-blocking stream on which everything it enqueues is ordered.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/b200ka.h"

namespace b200 {

constexpr int kMaxDevices = 16;

struct ThreadState {
    bool inited = false;
    int device = 0;
    int priority = 0;                              // -1 low, 0 normal, 1 high
    cudaStream_t own[kMaxDevices][3] = {};         // lazily created, indexed by priority + 1
    cudaStream_t external = nullptr;               // adopted via b200_set_stream
    bool use_external = false;
    int64_t launches = 0;
    char err[512] = {0};
};

ThreadState &tls();
int set_error(int status, const char *fmt, ...);   // records message, returns status
int ensure_device();                               // B200_OK or B200_ERR_NO_DEVICE; binds the thread's device
int current_stream(cudaStream_t *s);               // the calling thread's stream on its current device
int sm_count();                                    // cached multiprocessor count of the current device
inline void count_launch(int n = 1) { tls().launches += n; }

#define B200_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::b200::set_error(B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);                  \
    } while (0)

#define B200_CHECK_LAUNCH(name)                                                                    \
    do {                                                                                           \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return ::b200::set_error(B200_ERR_CUDA, "launch of %s failed: %s", name,               \
                                     cudaGetErrorString(_e));                                      \
        ::b200::count_launch();                                                                    \
    } while (0)

#define B200_TRY(expr)                                                                             \
    do {                                                                                           \
        int _s = (expr);                                                                           \
        if (_s != B200_OK) return _s;                                                              \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace b200
