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

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.B200_UNIQUE_ID_BYTES)
        _lib.call("b200_comm_unique_id", buf)
        return buf.raw

    def allreduce_sum_i32(self, ptr: int, count: int) -> None:
        _lib.call("b200_allreduce_sum_i32", self.handle, C.c_void_p(ptr), count)

    def broadcast(self, ptr: int, nbytes: int, root: int = 0) -> None:
        _lib.call("b200_broadcast", self.handle, C.c_void_p(ptr), nbytes, root)

    def destroy(self) -> None:
        if self.handle:
            _lib.call("b200_comm_destroy", self.handle)
            self.handle = C.c_void_p()
