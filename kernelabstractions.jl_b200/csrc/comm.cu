// comm.cu -- multi-GPU plumbing: NCCL over NVLink 5 / NVSwitch, one rank per GPU.
//
// The reference has no collective call sites in src/; its only communication code is the host-staged ring
// exchange of examples/mpi.jl:24-39 (Isend/Irecv + stream-ordered KA.copyto! + cooperative synchronize).
// The north-star sharding needs exactly one exchange: the 256 x Int32 partial histograms are summed with an
// all-reduce (integer adds: order independent, bit-exact).  NCCL is loaded lazily with dlopen so the library
// has no link-time dependency on it (single-GPU users never touch it); inside a torch process the already
// loaded libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace b200 {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclUint8 = 1, ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
#define B200_SYM(field, sym) *(void **)(&g_nccl.field) = dlsym(g_nccl.handle, sym)
    B200_SYM(GetUniqueId, "ncclGetUniqueId");
    B200_SYM(CommInitRank, "ncclCommInitRank");
    B200_SYM(CommDestroy, "ncclCommDestroy");
    B200_SYM(AllReduce, "ncclAllReduce");
    B200_SYM(Broadcast, "ncclBroadcast");
    B200_SYM(GetErrorString, "ncclGetErrorString");
    B200_SYM(Send, "ncclSend");
    B200_SYM(Recv, "ncclRecv");
    B200_SYM(GroupStart, "ncclGroupStart");
    B200_SYM(GroupEnd, "ncclGroupEnd");
#undef B200_SYM
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce &&
                g_nccl.Broadcast && g_nccl.GetErrorString && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart &&
                g_nccl.GroupEnd;
}

static int need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) return set_error(B200_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
    return B200_OK;
}

#define B200_NCCL(call)                                                                                       \
    do {                                                                                                      \
        ncclResult_t _r = (call);                                                                             \
        if (_r != 0) return set_error(B200_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(_r));      \
    } while (0)

}  // namespace b200

using namespace b200;

extern "C" b200_status b200_comm_unique_id(void *id128) {
    if (!id128) return set_error(B200_ERR_INVALID_VALUE, "id128 is NULL");
    B200_TRY(need_nccl());
    ncclUniqueId id;
    B200_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == B200_UNIQUE_ID_BYTES, "unique id size");
    memcpy(id128, &id, sizeof(id));
    return B200_OK;
}

extern "C" b200_status b200_comm_init_rank(void **comm, int32_t nranks, int32_t rank, const void *id128) {
    if (!comm || !id128) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(B200_ERR_INVALID_VALUE, "bad rank %d/%d", rank, nranks);
    B200_TRY(need_nccl());
    B200_TRY(ensure_device());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    B200_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank));
    *comm = (void *)c;
    return B200_OK;
}

extern "C" b200_status b200_comm_destroy(void *comm) {
    if (!comm) return B200_OK;
    B200_TRY(need_nccl());
    B200_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
    return B200_OK;
}

extern "C" b200_status b200_allreduce_sum_i32(void *comm, int32_t *buf, int64_t count) {
    if (!comm || (!buf && count > 0)) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (count <= 0) return B200_OK;
    B200_TRY(need_nccl());
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    B200_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclInt32, ncclSum, (ncclComm_t)comm, st));
    return B200_OK;
}

extern "C" b200_status b200_broadcast(void *comm, void *buf, size_t bytes, int32_t root) {
    if (!comm || (!buf && bytes > 0)) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (bytes == 0) return B200_OK;
    B200_TRY(need_nccl());
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    B200_NCCL(g_nccl.Broadcast(buf, buf, bytes, ncclUint8, root, (ncclComm_t)comm, st));
    return B200_OK;
}

// exchange! of examples/mpi.jl:24-39 without the host staging: send `bytes` to dst_rank and receive `bytes` from
// src_rank, device buffers, stream ordered (the MPI.Isend/Irecv pair becomes one NCCL group over NVLink).
extern "C" b200_status b200_sendrecv(void *comm, const void *sendbuf, int32_t dst_rank, void *recvbuf, int32_t src_rank,
                                     size_t bytes) {
    if (!comm || (bytes > 0 && (!sendbuf || !recvbuf))) return set_error(B200_ERR_INVALID_VALUE, "NULL argument");
    if (bytes == 0) return B200_OK;
    B200_TRY(need_nccl());
    cudaStream_t st;
    B200_TRY(current_stream(&st));
    B200_NCCL(g_nccl.GroupStart());
    ncclResult_t r1 = g_nccl.Send(sendbuf, bytes, ncclUint8, dst_rank, (ncclComm_t)comm, st);
    ncclResult_t r2 = g_nccl.Recv(recvbuf, bytes, ncclUint8, src_rank, (ncclComm_t)comm, st);
    B200_NCCL(g_nccl.GroupEnd());
    if (r1 != 0) return set_error(B200_ERR_NCCL, "ncclSend failed: %s", g_nccl.GetErrorString(r1));
    if (r2 != 0) return set_error(B200_ERR_NCCL, "ncclRecv failed: %s", g_nccl.GetErrorString(r2));
    return B200_OK;
}
