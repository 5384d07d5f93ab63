"""ctypes/numpy front end of the CPU parity oracle (oracle/ka_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under ``kernelabstractions.jl_b200/`` imports this module.

All arrays are numpy arrays in Fortran (column-major) order, mirroring Julia `Array`s; every function
cites the reference lines the C restatement follows (see ka_oracle.c for the details).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libka_oracle.so")
KA_MAX_NDIM = 4


def build(force: bool = False) -> str:
    """Compile libka_oracle.so with the Makefile beside it (gcc -O3 -fopenmp -ffp-contract=off)."""
    src = os.path.join(_HERE, "ka_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class Ctx(C.Structure):
    """Flattened CompilerMetadata + NDRange (reference src/compiler.jl:1-33, src/nditeration.jl:49-60)."""

    _fields_ = [
        ("ndim", C.c_int32),
        ("ndrange", C.c_int64 * KA_MAX_NDIM),
        ("wgs", C.c_int64 * KA_MAX_NDIM),
        ("blocks", C.c_int64 * KA_MAX_NDIM),
        ("dynamic_check", C.c_int32),
    ]

    def groups(self) -> int:
        return int(np.prod([self.blocks[d] for d in range(self.ndim)], dtype=np.int64))

    def items(self) -> int:
        return int(np.prod([self.wgs[d] for d in range(self.ndim)], dtype=np.int64))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ka_oracle_threads.restype = C.c_int
        _lib.ka_num_groups.restype = C.c_int64
        _lib.ka_group_items.restype = C.c_int64
        _lib.ka_validindex.restype = C.c_int
        _lib.ka_partition.restype = C.c_int
    return _lib


def threads() -> int:
    return int(lib().ka_oracle_threads())


def set_threads(n: int) -> None:
    """OpenMP thread count of the timed CPU baseline (torchrun exports OMP_NUM_THREADS=1 to its ranks)."""
    lib().ka_oracle_set_threads(C.c_int(int(n)))


def _i64arr(vals: Sequence[int]):
    return (C.c_int64 * max(len(vals), 1))(*[int(v) for v in vals])


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------------------------
# Host logic
# --------------------------------------------------------------------------------------------
def partition(ndrange: Sequence[int], workgroupsize: Sequence[int]) -> Tuple[Tuple[int, ...], Tuple[int, ...], bool]:
    """NDIteration.partition (reference src/nditeration.jl:143-164) -> (blocks, wgs, dynamic)."""
    n = len(ndrange)
    blocks = (C.c_int64 * KA_MAX_NDIM)()
    wgs = (C.c_int64 * KA_MAX_NDIM)()
    dyn = C.c_int32(0)
    rc = lib().ka_partition(C.c_int32(n), _i64arr(ndrange), C.c_int32(len(workgroupsize)), _i64arr(workgroupsize),
                            blocks, wgs, C.byref(dyn))
    if rc != 0:
        raise AssertionError("length(workgroupsize) <= length(ndrange)")
    return tuple(blocks[d] for d in range(n)), tuple(wgs[d] for d in range(n)), bool(dyn.value)


def threads_to_workgroupsize(nthreads: int, ndrange: Sequence[int]) -> Tuple[int, ...]:
    """reference src/pocl/backend.jl:108-115"""
    out = (C.c_int64 * KA_MAX_NDIM)()
    lib().ka_threads_to_workgroupsize(C.c_int64(nthreads), C.c_int32(len(ndrange)), _i64arr(ndrange), out)
    return tuple(out[d] for d in range(len(ndrange)))


def make_ctx(ndrange, workgroupsize, dynamic_check: Optional[bool] = True) -> Ctx:
    """partition + mkcontext.  The in-tree backend always builds a DynamicCheck context
    (reference src/pocl/backend.jl:74-76), hence the default."""
    if isinstance(ndrange, (int, np.integer)):
        ndrange = (int(ndrange),)
    if isinstance(workgroupsize, (int, np.integer)):
        workgroupsize = (int(workgroupsize),)
    blocks, wgs, dyn = partition(ndrange, workgroupsize)
    c = Ctx()
    c.ndim = len(ndrange)
    for d in range(len(ndrange)):
        c.ndrange[d] = ndrange[d]
        c.wgs[d] = wgs[d]
        c.blocks[d] = blocks[d]
    c.dynamic_check = int(dyn if dynamic_check is None else dynamic_check)
    return c


def raw_ctx(blocks: Sequence[int], wgs: Sequence[int]) -> Ctx:
    """An NDRange given directly by blocks x workitems (reference test/nditeration.jl:6-17)."""
    c = Ctx()
    c.ndim = len(blocks)
    for d in range(len(blocks)):
        c.blocks[d] = blocks[d]
        c.wgs[d] = wgs[d]
        c.ndrange[d] = blocks[d] * wgs[d]
    c.dynamic_check = 0
    return c


def expand(ctx: Ctx, g: int, l: int) -> Tuple[int, ...]:
    """expand(ndrange, groupidx::Integer, idx::Integer) (reference src/nditeration.jl:115-126), 1-based."""
    out = (C.c_int64 * KA_MAX_NDIM)()
    lib().ka_expand(C.byref(ctx), C.c_int64(g), C.c_int64(l), out)
    return tuple(out[d] for d in range(ctx.ndim))


def validindex(ctx: Ctx, g: int, l: int) -> bool:
    return bool(lib().ka_validindex(C.byref(ctx), C.c_int64(g), C.c_int64(l)))


def linear_iteration(ctx: Ctx) -> np.ndarray:
    """reference test/nditeration.jl:20-30 -> array of shape (groups*items, ndim)."""
    out = np.zeros((ctx.groups() * ctx.items(), ctx.ndim), dtype=np.int64)
    lib().ka_linear_iteration(C.byref(ctx), _p(out))
    return out


# --------------------------------------------------------------------------------------------
# KA @kernel twins
# --------------------------------------------------------------------------------------------
def _F(a: np.ndarray) -> np.ndarray:
    assert a.flags.f_contiguous or a.ndim <= 1, "oracle arrays are column-major"
    return a


def copy_kernel(A: np.ndarray, B: np.ndarray, ndrange, workgroupsize) -> None:
    """reference src/KernelAbstractions.jl:874-877 / examples/memcopy.jl:4-7"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_copy_kernel(C.byref(ctx), _p(_F(A)), _p(_F(B)), C.c_int64(A.dtype.itemsize))


def init_kernel(arr: np.ndarray, value, ndrange, workgroupsize) -> None:
    """reference src/KernelAbstractions.jl:869-872"""
    ctx = make_ctx(ndrange, workgroupsize)
    v = np.array([value], dtype=arr.dtype)
    lib().orc_init_kernel(C.byref(ctx), _p(_F(arr)), _p(v), C.c_int64(arr.dtype.itemsize))


def naive_transpose(a: np.ndarray, b: np.ndarray, ndrange=None, workgroupsize=256) -> None:
    """reference examples/naive_transpose.jl:4-21 (kernel built with workgroupsize 256, ndrange=size(a))"""
    ndrange = a.shape if ndrange is None else ndrange
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_naive_transpose(C.byref(ctx), _p(_F(a)), _p(_F(b)), C.c_int64(a.shape[0]), C.c_int64(b.shape[0]),
                              C.c_int64(a.dtype.itemsize))


def naive_transpose_wg(a: np.ndarray, b: np.ndarray, workgroupsize=256) -> None:
    """Work-group-vectorised execution of the same kernel (the timed CPU baseline); float32 only."""
    assert a.dtype == np.float32 and b.dtype == np.float32
    ctx = make_ctx(a.shape, workgroupsize)
    lib().orc_naive_transpose_wg_f32(C.byref(ctx), _p(_F(a)), _p(_F(b)), C.c_int64(a.shape[0]), C.c_int64(b.shape[0]))


def copy_wg(A: np.ndarray, B: np.ndarray, workgroupsize=1024) -> None:
    """Work-group-vectorised 1-d copy_kernel! (timed CPU baseline); float32 only."""
    assert A.dtype == np.float32 and B.dtype == np.float32 and A.ndim == 1 and A.shape == B.shape
    ctx = make_ctx(A.shape, workgroupsize)
    lib().orc_copy_wg_f32(C.byref(ctx), _p(A), _p(B))


def saxpy_wg(Z: np.ndarray, a, X: np.ndarray, Y: np.ndarray, workgroupsize=1024) -> None:
    """Work-group-vectorised 1-d saxpy_kernel! (timed CPU baseline); float32 only."""
    assert Z.dtype == X.dtype == Y.dtype == np.float32 and Z.ndim == 1
    ctx = make_ctx(Z.shape, workgroupsize)
    lib().orc_saxpy_wg_f32(C.byref(ctx), _p(Z), C.c_float(a), _p(X), _p(Y))


def matmul_naive(out: np.ndarray, a: np.ndarray, b: np.ndarray, workgroupsize=None, max_threads=1024) -> None:
    """reference examples/matmul.jl:5-27 (dynamic workgroup -> threads_to_workgroupsize)"""
    ndrange = out.shape
    if workgroupsize is None:
        workgroupsize = threads_to_workgroupsize(max_threads, ndrange)
    ctx = make_ctx(ndrange, workgroupsize)
    f = lib().orc_matmul_naive_f32 if out.dtype == np.float32 else lib().orc_matmul_naive_f64
    f(C.byref(ctx), _p(_F(out)), _p(_F(a)), _p(_F(b)), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]))


PERF_KERNELS = ("simple_copy", "simple_transpose", "lmem_copy", "lmem_transpose", "coalesced_copy", "coalesced_transpose")


def performance_kernel(which, output: np.ndarray, input: np.ndarray, ndrange, workgroupsize, bank: int = 1) -> None:
    """reference examples/performance.jl:17-138; `which` is an index into PERF_KERNELS or one of its names.
    4-byte elements; kernels 2-5 are unsafe_indices kernels (launch must stay inside the arrays)."""
    which = PERF_KERNELS.index(which) if isinstance(which, str) else int(which)
    assert output.dtype.itemsize == 4 and input.dtype.itemsize == 4
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_performance_kernel(C.byref(ctx), C.c_int(which), _p(_F(output)), _p(_F(input)), C.c_int64(output.shape[0]),
                                 C.c_int64(input.shape[0]), C.c_int64(bank))


def localmem(variant: int, A: np.ndarray, ndrange, workgroupsize) -> None:
    """reference test/localmem.jl:4-65; variant 0 localmem, 1 localmem2, 2 unsafe_indices, 3 many_localmem"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_localmem(C.byref(ctx), C.c_int(variant), _p(A), C.c_int64(A.size))


def private(A: np.ndarray, ndrange, workgroupsize) -> None:
    """reference test/private.jl:18-28"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_private(C.byref(ctx), _p(A))


def typetest(B: np.ndarray, ndrange, workgroupsize) -> None:
    """reference test/private.jl:10-16"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_typetest(C.byref(ctx), _p(B))


def forloop(A: np.ndarray, ndrange, workgroupsize) -> None:
    """reference test/private.jl:31-45"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_forloop(C.byref(ctx), _p(_F(A)), C.c_int64(A.shape[0]), C.c_int64(A.shape[1]))


def reduce_private(out: np.ndarray, A: np.ndarray, ndrange, workgroupsize) -> None:
    """reference test/private.jl:47-59"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_reduce_private_f32(C.byref(ctx), _p(out), _p(_F(A)), C.c_int64(A.shape[0]), C.c_int64(A.shape[1]))


def unroll(variant: int, a: np.ndarray, N: int = 5) -> None:
    """reference test/unroll.jl:5-37, launched as (backend, 1, 1)"""
    ctx = make_ctx((1,), (1,))
    lib().orc_unroll_f32(C.byref(ctx), C.c_int(variant), _p(a), C.c_int64(N))


def index_linear(which: int, A: np.ndarray, ndrange, workgroupsize) -> None:
    """reference test/test.jl:46-58"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_index_linear(C.byref(ctx), C.c_int(which), _p(_F(A)))


def index_cartesian(which: int, A: np.ndarray, shapeA, ndrange, workgroupsize) -> None:
    """reference test/test.jl:59-73; A is int64 of shape (n, *shapeA) in F order (CartesianIndex{n} elements)"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_index_cartesian(C.byref(ctx), C.c_int(which), _p(_F(A)), C.c_int32(len(shapeA)), _i64arr(shapeA))


def kernel_val(a: np.ndarray, m: int, ndrange, workgroupsize) -> None:
    """reference test/test.jl:216-224"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_kernel_val(C.byref(ctx), _p(a), C.c_int64(m))


def misc_i64(which: int, a: np.ndarray, ndrange, workgroupsize) -> None:
    """which 0: context_kernel (reference test/test.jl:275-282); 1: gpu_return_kernel! (:307-313)"""
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_misc_i64(C.byref(ctx), C.c_int(which), _p(a), C.c_int64(a.size))


def unaliased_accumulate(local: bool, output: np.ndarray, input: np.ndarray, n: int, ndrange, workgroupsize) -> None:
    """reference test/test.jl:339-356 (Float32)"""
    assert output.dtype == np.float32 and input.dtype == np.float32
    ctx = make_ctx(ndrange, workgroupsize)
    lib().orc_unaliased_accumulate_f32(C.byref(ctx), C.c_int(int(local)), _p(_F(output)), _p(_F(input)), C.c_int64(n),
                                       C.c_int64(output.shape[0]), C.c_int64(input.shape[0]))


def saxpy(Z, a, X, Y, ndrange, workgroupsize) -> None:
    """reference benchmark/benchmarks.jl:29-32"""
    ctx = make_ctx(ndrange, workgroupsize)
    if Z.dtype == np.float64:
        lib().orc_saxpy_f64(C.byref(ctx), _p(Z), C.c_double(a), _p(X), _p(Y))
    else:
        lib().orc_saxpy_f32(C.byref(ctx), _p(Z), C.c_float(a), _p(X), _p(Y))


# --------------------------------------------------------------------------------------------
# KI twins
# --------------------------------------------------------------------------------------------
def create_histogram(inp: np.ndarray) -> np.ndarray:
    """reference examples/histogram.jl:9-15"""
    out = np.zeros(int(inp.max()), dtype=np.int32)
    lib().orc_create_histogram_i32(_p(out), _p(inp), C.c_int64(inp.size))
    return out


def histogram(out: np.ndarray, inp: np.ndarray, groupsize: int = 256) -> None:
    """reference examples/histogram.jl:18-66 (accumulates into out)"""
    assert out.dtype == np.int32 and inp.dtype == np.int32
    lib().orc_histogram_i32(_p(out), C.c_int64(out.size), _p(inp), C.c_int64(inp.size), C.c_int64(groupsize))


def coalesced_matmul(output: np.ndarray, in1: np.ndarray, in2: np.ndarray, tdim: int = 16) -> None:
    """reference examples/performant_matmul.jl:15-92"""
    N, R = in1.shape
    M = in2.shape[1]
    lib().orc_coalesced_matmul_f32(_p(_F(output)), _p(_F(in1)), _p(_F(in2)), C.c_int64(N), C.c_int64(R),
                                   C.c_int64(M), C.c_int64(tdim))


def matmul_running(Cm: np.ndarray, A: np.ndarray, B: np.ndarray) -> None:
    M, K = A.shape
    N = B.shape[1]
    lib().orc_matmul_running_f32(_p(_F(Cm)), _p(_F(A)), _p(_F(B)), C.c_int64(M), C.c_int64(K), C.c_int64(N))


def matmul_f64_truth(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    M, K = A.shape
    N = B.shape[1]
    out = np.zeros((M, N), dtype=np.float64, order="F")
    lib().orc_matmul_f64_truth(_p(out), _p(_F(A)), _p(_F(B)), C.c_int64(M), C.c_int64(K), C.c_int64(N))
    return out


def round_bf16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bfloat16, returned as float32 (the tensor-core path's operand rounding)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape, order="A")


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16).reshape(x.shape)


def test_intrinsics(numworkgroups: int, workgroupsize: int, length: int) -> np.ndarray:
    """reference test/intrinsics.jl:11-25 -> (length, 6) int64"""
    out = np.zeros((length, 6), dtype=np.int64)
    lib().orc_test_intrinsics(_p(out), C.c_int64(length), C.c_int64(numworkgroups), C.c_int64(workgroupsize))
    return out


def launch_kernel_nd(arr: np.ndarray, numworkgroups: Sequence[int], workgroupsize: Sequence[int]) -> None:
    """reference test/intrinsics.jl:31-72"""
    ng = list(numworkgroups) + [1] * (3 - len(numworkgroups))
    ws = list(workgroupsize) + [1] * (3 - len(workgroupsize))
    dims = list(arr.shape) + [1] * (3 - arr.ndim)
    lib().orc_launch_kernel_nd(_p(_F(arr)), _i64arr(dims), _i64arr(ng), _i64arr(ws))


# --------------------------------------------------------------------------------------------
# Synthetic inputs: counter-hash streams, identical wherever they are generated.
# --------------------------------------------------------------------------------------------
def gen_uniform_f32(shape, seed: int) -> np.ndarray:
    out = np.empty(shape, dtype=np.float32, order="F")
    lib().orc_gen_uniform_f32(_p(out), C.c_int64(out.size), C.c_uint64(seed))
    return out


def gen_bins_i32(n: int, seed: int, nbins: int = 256) -> np.ndarray:
    out = np.empty(n, dtype=np.int32)
    lib().orc_gen_bins_i32(_p(out), C.c_int64(n), C.c_uint64(seed), C.c_int32(nbins))
    return out
