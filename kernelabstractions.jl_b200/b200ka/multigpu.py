"""Multi-GPU sharding of the north-star kernels: one process (rank) per GPU, work row/slab partitioned.

Pattern: reference examples/mpi.jl:24-39 (host orchestrates; device work is stream ordered; waits are
cooperative).  copy and transpose shard with NO data-path collective; histogram needs exactly one exchange
(sum of the 256 partial bins); matmul shards A/C row panels with B replicated.

The partitioning arithmetic is pure host logic (tested on CPU with gloo, world_size 2); the collective goes
through NCCL, either torch.distributed's communicator or the library's own b200_comm_* entry points.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib


def slab(n: int, world_size: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous share of n units, multiples of `align` (last rank takes the remainder)."""
    per = (n // world_size) // align * align
    start = rank * per
    stop = n if rank == world_size - 1 else start + per
    return start, stop


def copy_shard(n_elems: int, world_size: int, rank: int) -> Tuple[int, int]:
    """copy: contiguous element slabs, no collective."""
    return slab(n_elems, world_size, rank, align=4)


def transpose_shard(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """transpose a[i,j] = b[j,i]: rank owns output columns j in [j0, j1) of `a` (a contiguous n x (j1-j0) slab);
    it needs rows j0..j1 of `b`, held as a (j1-j0) x n column-major slab.  No collective."""
    return slab(n, world_size, rank, align=64)


def histogram_shard(n_elems: int, world_size: int, rank: int) -> Tuple[int, int]:
    """histogram: contiguous input slabs -> per-GPU partial bins -> one all-reduce (sum, Int32)."""
    return slab(n_elems, world_size, rank, align=4)


def matmul_shard(M: int, world_size: int, rank: int) -> Tuple[int, int]:
    """matmul: rank owns rows [m0, m1) of A and C; B is replicated.  k is not split, so no reduction and the
    FP32 result is identical to the 1-GPU result."""
    return slab(M, world_size, rank, align=128)


def allreduce_bins_torch(bins_tensor) -> None:
    """Sum partial histograms across ranks with torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    dist.all_reduce(bins_tensor, op=dist.ReduceOp.SUM)


class Comm:
    """The library's own NCCL communicator (b200_comm_*), for hosts without torch (e.g. the Julia layer + MPI.jl):
    rank 0 creates the unique id, the host's launcher broadcasts the 128 bytes, every rank calls init."""

    def __init__(self, world_size: int, rank: int, unique_id: bytes):
        self.handle = C.c_void_p()
        buf = C.create_string_buffer(unique_id, _lib.B200_UNIQUE_ID_BYTES)
        _lib.call("b200_comm_init_rank", C.byref(self.handle), world_size, rank, buf)
        self.world_size, self.rank = world_size, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.B200_UNIQUE_ID_BYTES)
        _lib.call("b200_comm_unique_id", buf)
        return buf.raw

    def allreduce_sum_i32(self, ptr: int, count: int) -> None:
        _lib.call("b200_allreduce_sum_i32", self.handle, C.c_void_p(ptr), count)

    def broadcast(self, ptr: int, nbytes: int, root: int = 0) -> None:
        _lib.call("b200_broadcast", self.handle, C.c_void_p(ptr), nbytes, root)

    def sendrecv(self, send_ptr: int, dst_rank: int, recv_ptr: int, src_rank: int, nbytes: int) -> None:
        _lib.call("b200_sendrecv", self.handle, C.c_void_p(send_ptr), dst_rank, C.c_void_p(recv_ptr), src_rank, nbytes)

    def destroy(self) -> None:
        if self.handle:
            _lib.call("b200_comm_destroy", self.handle)
            self.handle = C.c_void_p()


class PeerGroup:
    """Symmetric NVLink peer buffers for the fused compute+collective kernels (b200_p2p_*): every rank allocates its
    buffer, the 64-byte IPC handles are all-gathered by the launcher (`gather` = a callable bytes -> list of bytes, e.g.
    built on torch.distributed.all_gather_object or MPI.Allgather), and every rank maps its peers."""

    def __init__(self, world_size: int, rank: int, gather):
        self.handle = C.c_void_p()
        _lib.call("b200_p2p_alloc", C.byref(self.handle))
        buf = C.create_string_buffer(_lib.B200_IPC_HANDLE_BYTES)
        _lib.call("b200_p2p_export", self.handle, buf)
        handles = gather(buf.raw)
        assert len(handles) == world_size and all(len(h) == _lib.B200_IPC_HANDLE_BYTES for h in handles)
        allh = C.create_string_buffer(b"".join(handles), world_size * _lib.B200_IPC_HANDLE_BYTES)
        _lib.call("b200_p2p_connect", self.handle, world_size, rank, allh)
        self.world_size, self.rank = world_size, rank

    def histogram_allreduce(self, bins, input) -> None:
        """bins += histogram of every rank's `input`, on every rank, in one fused kernel per rank."""
        _lib.call("b200_histogram_i32_allreduce", self.handle, C.c_void_p(bins.ptr), bins.size, C.c_void_p(input.ptr), input.size)

    def debug_stamps(self):
        """ns timeline of the last fused launch after its slowest CTA left the counting loop: [counts_merged, slots_sent, all_summed]."""
        buf = (C.c_uint64 * 4)()
        _lib.call("b200_p2p_debug", self.handle, buf)
        return [int(buf[i]) - int(buf[0]) for i in (1, 2, 3)]

    def timeline(self):
        """ns timeline of the last fused launch relative to the moment its FASTEST CTA left the counting loop (b200_p2p_timeline)."""
        buf = (C.c_uint64 * 8)()
        _lib.call("b200_p2p_timeline", self.handle, buf)
        t0 = int(buf[4])
        return {"first_cta_resident": int(buf[5]) - t0, "slowest_cta_done": int(buf[0]) - t0, "counts_merged": int(buf[1]) - int(buf[0]),
                "all_published": int(buf[6]) - int(buf[0]), "partial_read": int(buf[7]) - int(buf[0]),
                "slots_sent": int(buf[2]) - int(buf[0]), "all_summed": int(buf[3]) - int(buf[0])}

    def free(self) -> None:
        if self.handle:
            _lib.call("b200_p2p_free", self.handle)
            self.handle = C.c_void_p()


def torch_gather(dist):
    """`gather` callable for PeerGroup on top of torch.distributed (any backend)."""
    def gather(blob: bytes):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, blob)
        return out
    return gather


def exchange_(comm: "Comm", backend, d_send_buf, d_recv_buf, src_rank: int, dst_rank: int) -> None:
    """exchange! of reference examples/mpi.jl:24-39 on device buffers: the ring neighbour's payload lands in
    d_recv_buf, stream ordered behind whatever produced d_send_buf; KA.synchronize(backend) completes it."""
    assert d_send_buf.nbytes == d_recv_buf.nbytes
    comm.sendrecv(d_send_buf.ptr, dst_rank, d_recv_buf.ptr, src_rank, d_send_buf.nbytes)


class SymBuffer:
    """A symmetric device buffer (b200_sym_*): every rank allocates `nbytes`, the IPC handles are all-gathered with `gather`
    (see PeerGroup), and every rank maps its peers, so kernels can store straight into another GPU's memory over NVLink."""

    def __init__(self, nbytes: int, world_size: int, rank: int, gather):
        self.handle = C.c_void_p()
        _lib.call("b200_sym_alloc", C.byref(self.handle), C.c_size_t(nbytes))
        buf = C.create_string_buffer(_lib.B200_IPC_HANDLE_BYTES)
        _lib.call("b200_sym_export", self.handle, buf)
        handles = gather(buf.raw)
        allh = C.create_string_buffer(b"".join(handles), world_size * _lib.B200_IPC_HANDLE_BYTES)
        _lib.call("b200_sym_connect", self.handle, world_size, rank, allh)
        self.world_size, self.rank, self.nbytes = world_size, rank, nbytes

    def ptr(self, rank: int = -1) -> int:
        p = C.c_void_p()
        _lib.call("b200_sym_ptr", self.handle, C.c_int32(rank), C.byref(p))
        return p.value

    def array(self, dtype, shape):
        """the own buffer as a (non-owning) B200Array view"""
        from .backend import B200Array
        return B200Array(dtype, shape, _ptr=self.ptr(), _owner=self)

    def barrier(self) -> None:
        _lib.call("b200_sym_barrier", self.handle)

    def free(self) -> None:
        if self.handle:
            _lib.call("b200_sym_free", self.handle)
            self.handle = C.c_void_p()


def transpose_alltoall_blocks(N0: int, N1: int, world_size: int, rank: int):
    """The block plan of the distributed transpose (pure host logic, tested on CPU): for every destination h the tuple
    (h, first source row in b_local, first destination row in rank h's slab); a block is N0/world x N1/world in b_local and its
    transpose N1/world x N0/world lands at that row of a_h (all columns)."""
    if N0 % world_size or N1 % world_size:
        raise _lib.ArgumentError("transpose_alltoall: N0 and N1 must be multiples of the world size")
    rb, cb = N0 // world_size, N1 // world_size
    return [(h, h * rb, rank * cb) for h in range(world_size)]


def transpose_alltoall_(a_sym: SymBuffer, b_local, N0: int, N1: int) -> None:
    """Distributed a = transpose(b) for a global N0 x N1 matrix sharded by columns (b_local: N0 x N1/world on this rank; afterwards
    a_sym of rank h holds a[:, h-th block]: N1 x N0/world): the fused tile transpose + all-to-all of b200_transpose_alltoall."""
    _lib.call("b200_transpose_alltoall", a_sym.handle, C.c_void_p(b_local.ptr), N0, N1, C.c_int32(b_local.dtype.itemsize))


def transpose_alltoall_nccl_(comm: "Comm", backend, a_local, b_local, N0: int, N1: int, stage_send, stage_recv) -> None:
    """The same result with library collectives, for comparison and as a parity check: pack each destination's block, exchange the
    blocks pairwise with NCCL send/recv (world-1 rounds), transpose every received block locally into a_local."""
    G, g = comm.world_size, comm.rank
    rb, cb = N0 // G, N1 // G
    es = b_local.dtype.itemsize
    blk_bytes = rb * cb * es

    for h in range(G):     # pack: contiguous rb x cb block for destination h  (copy of a strided sub-matrix = transpose of transpose ...)
        _lib.call("b200_memcpy_2d_d2d_async", C.c_void_p(stage_send.ptr + h * blk_bytes), C.c_size_t(rb * es), C.c_void_p(b_local.ptr + h * rb * es),
                  C.c_size_t(N0 * es), C.c_size_t(rb * es), C.c_size_t(cb))
    for r in range(G):
        dst, src = (g + r) % G, (g - r) % G
        if r == 0:
            _lib.call("b200_memcpy_d2d_async", C.c_void_p(stage_recv.ptr + g * blk_bytes), C.c_void_p(stage_send.ptr + g * blk_bytes), C.c_size_t(blk_bytes))
        else:
            comm.sendrecv(stage_send.ptr + dst * blk_bytes, dst, stage_recv.ptr + src * blk_bytes, src, blk_bytes)
    for s in range(G):     # block from rank s (rb x cb) -> rows c1(s) of a_local (cb x rb block, ld N1)
        _lib.call("b200_transpose", C.c_void_p(a_local.ptr + s * cb * es), C.c_void_p(stage_recv.ptr + s * blk_bytes), cb, rb, N1, rb, es)
