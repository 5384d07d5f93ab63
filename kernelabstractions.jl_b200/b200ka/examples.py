"""The reference's example drivers, written against the mirror API exactly as the Julia files are written
against KA (reference examples/*.jl).  These are what a user switching backends would run unchanged."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from . import kernels
from .backend import B200Array, B200Backend, allocate, copyto_, get_backend, zeros
from .kernel import Val, ki_kernel

TILE_DIM = 16  # reference examples/performant_matmul.jl:13


def mycopy_(A: B200Array, B: B200Array) -> None:
    """reference examples/memcopy.jl:9-17"""
    backend = get_backend(A)
    assert A.shape == B.shape
    assert get_backend(B) == backend
    kernel = kernels.copy_kernel_(backend)
    kernel(A, B, ndrange=A.size)


def mycopy_static_(A: B200Array, B: B200Array) -> None:
    """reference examples/memcopy_static.jl:9-17"""
    backend = get_backend(A)
    assert A.shape == B.shape
    assert get_backend(B) == backend
    kernel = kernels.copy_kernel_(backend, 32, A.shape)
    kernel(A, B, ndrange=A.shape)


def naive_transpose_(a: B200Array, b: B200Array):
    """reference examples/naive_transpose.jl:11-21"""
    if a.shape[0] != b.shape[1] or a.shape[1] != b.shape[0]:
        print("Matrix size mismatch!")
        return None
    backend = get_backend(a)
    assert get_backend(b) == backend
    kernel_ = kernels.naive_transpose_kernel_(backend, 256)
    kernel_(a, b, ndrange=a.shape)
    return None


def matmul_(output: B200Array, a: B200Array, b: B200Array):
    """reference examples/matmul.jl:17-27"""
    if a.shape[1] != b.shape[0]:
        print("Matrix size mismatch!")
        return None
    backend = get_backend(a)
    kernel_ = kernels.matmul_kernel_(backend)
    kernel_(output, a, b, ndrange=output.shape)
    return None


def performant_matmul_(C_: B200Array, A: B200Array, B: B200Array) -> None:
    """reference examples/performant_matmul.jl:86-89"""
    backend = get_backend(C_)
    N, R = A.shape
    M = B.shape[1]
    workgroupsize = (TILE_DIM, TILE_DIM)
    numworkgroups = (math.ceil(C_.shape[0] / TILE_DIM), math.ceil(C_.shape[1] / TILE_DIM))
    ki_kernel(backend, "coalesced_matmul_kernel!", C_, A, B, N, R, M, Val(TILE_DIM),
              workgroupsize=workgroupsize, numworkgroups=numworkgroups)


def histogram_(histogram_output: B200Array, input: B200Array, groupsize: int = 256) -> None:
    """reference examples/histogram.jl:60-66"""
    backend = get_backend(histogram_output)
    ki_kernel(backend, "histogram_kernel!", histogram_output, input, Val(groupsize),
              workgroupsize=groupsize, numworkgroups=math.ceil(input.size / groupsize))


def move(backend, input: np.ndarray) -> B200Array:
    """reference examples/histogram.jl:68-72"""
    out = allocate(backend, input.dtype, input.shape)
    copyto_(backend, out, input)
    return out


def performance_variants(backend, N: int = 2048, T=np.float32, TILE_DIM: int = 32, BLOCK_ROWS: int = 8):
    """reference examples/performance.jl:140-205: yields (label, launch) for the script's 14 launch shapes.
    `launch(output, input)` enqueues one kernel exactly as the script does (same block_dims / Val(bank) / ndrange)."""
    for block_dims in ((TILE_DIM, TILE_DIM), (TILE_DIM * TILE_DIM, 1), (1, TILE_DIM * TILE_DIM)):
        for name, ctor in (("copy", kernels.simple_copy_kernel_), ("transpose", kernels.simple_transpose_kernel_)):
            k = ctor(backend, block_dims)
            yield f"Simple {name} {block_dims}", (lambda o, i, k=k: k(o, i, ndrange=o.shape))
    for name, ctor in (("copy", kernels.lmem_copy_kernel_), ("transpose", kernels.lmem_transpose_kernel_)):
        k = ctor(backend, (TILE_DIM, TILE_DIM))
        for bank in (True, False):
            yield (f"Localmem {name} ({TILE_DIM}, {TILE_DIM}) bank={bank}",
                   (lambda o, i, k=k, bank=bank: k(o, i, Val(int(bank)), ndrange=o.shape)))
    for name, ctor in (("copy", kernels.coalesced_copy_kernel_), ("transpose", kernels.coalesced_transpose_kernel_)):
        k = ctor(backend, (TILE_DIM, BLOCK_ROWS))
        for bank in (True, False):
            block_factor = TILE_DIM // BLOCK_ROWS
            yield (f"Localmem + multiple elements {name} ({TILE_DIM}, {BLOCK_ROWS}) bank={bank}",
                   (lambda o, i, k=k, bank=bank, bf=block_factor: k(o, i, Val(int(bank)), ndrange=(o.shape[0], o.shape[1] // bf))))


def measure_membw(backend, N: int = 1024 * 500_000, dtype=np.float32, evals: int = 2, samples: int = 10, X=None, Y=None):
    """reference examples/numa_aware.jl:21-56: time the in-place SAXPY `Y = a*X + Y`, return (GB/s, GFLOP/s)."""
    import time
    from .backend import synchronize, to_device
    nbytes = 3 * np.dtype(dtype).itemsize * N
    flops = 2 * N
    a = np.dtype(dtype).type(3.1415)
    if X is None:
        rng = np.random.default_rng(0)
        X = to_device(rng.random(N, dtype=dtype))
        Y = to_device(rng.random(N, dtype=dtype))
    best = float("inf")
    for _ in range(samples):
        synchronize(backend)
        t0 = time.perf_counter()
        for _ in range(evals):
            kernel = kernels.saxpy_kernel(backend, 1024, Y.shape)
            kernel(a, X, Y, ndrange=Y.shape)
            synchronize(backend)
        best = min(best, (time.perf_counter() - t0) / evals)
    return nbytes * 1.0e-9 / best, flops * 1.0e-9 / best


def matmul_bf16_(C_: B200Array, A: B200Array, B: B200Array, workspace: B200Array = None) -> None:
    """Opt-in tensor-core path (no reference twin): FP32 in/out, BF16 tcgen05 arithmetic."""
    M, K = A.shape
    N = B.shape[1]
    ws = workspace if workspace is not None else allocate(B200Backend(), np.uint16, (M * K + K * N,))
    _lib.call("b200_matmul_f32_via_bf16", C.c_void_p(C_.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), M, K, N, M, K,
              C_.shape[0], C.c_void_p(ws.ptr))


def prepare_bf16_b(B: B200Array) -> B200Array:
    """B (K x N, FP32, column-major) -> its resident BF16 image for matmul_bf16_resident_b_ (converted once, e.g. after the
    set-up broadcast that replicates B on every rank)."""
    K, N = B.shape
    out = allocate(B200Backend(), np.uint16, (K * N,))
    _lib.call("b200_convert_f32_bf16", C.c_void_p(out.ptr), C.c_void_p(B.ptr), K, N, K, 0)
    return out


def matmul_bf16_resident_b_(C_: B200Array, A: B200Array, Bbf: B200Array, N: int, workspace: B200Array = None) -> None:
    """C = A * B with FP32 A and C, BF16 tcgen05 arithmetic, B already converted (prepare_bf16_b): only the A (row panel) is
    converted per call.  Same bits as matmul_bf16_."""
    M, K = A.shape
    if Bbf.size != K * N or C_.shape != (M, N):
        raise _lib.ArgumentError("matmul_bf16_resident_b_: shapes do not conform")
    ws = workspace if workspace is not None else allocate(B200Backend(), np.uint16, (M * K,))
    _lib.call("b200_matmul_f32_bf16_resident_b", C.c_void_p(C_.ptr), C.c_void_p(A.ptr), C.c_void_p(Bbf.ptr), M, K, N, M, C_.shape[0],
              C.c_void_p(ws.ptr))


def matmul_tf32_(C_: B200Array, A: B200Array, B: B200Array) -> None:
    """Opt-in tensor-core path (no reference twin): tcgen05 kind::tf32 straight from the FP32 column-major operands --
    no conversion pass, no workspace; the tensor core uses the top 19 bits of every operand, FP32 accumulate."""
    M, K = A.shape
    N = B.shape[1]
    if B.shape[0] != K or C_.shape != (M, N):
        raise _lib.ArgumentError("matmul_tf32_: shapes do not conform")
    _lib.call("b200_matmul_tf32", C.c_void_p(C_.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), M, K, N, M, K, M)


def matmul_tf32x3_(C_: B200Array, A: B200Array, B: B200Array, workspace: B200Array = None) -> None:
    """Opt-in FP32-grade tensor-core product ("3xTF32", no reference twin): split operands, three tcgen05 MMAs per k step, K
    accumulated in chunks that are added to C in FP32.  rtol 1e-5 against Float64 for the examples' inputs, not bit-exact."""
    M, K = A.shape
    N = B.shape[1]
    if B.shape[0] != K or C_.shape != (M, N):
        raise _lib.ArgumentError("matmul_tf32x3_: shapes do not conform")
    ws = workspace if workspace is not None else allocate(B200Backend(), np.float32, (M * K + K * N,))
    _lib.call("b200_matmul_f32_tf32x3", C.c_void_p(C_.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), M, K, N, M, K, M, C.c_void_p(ws.ptr))


# ---- host-array forms: what `naive_transpose!(a, b)` / `histogram!(out, in)` are for a CPU() user -- plain host
# ---- arrays in and out -- executed by B200Backend as a slab pipeline (h2d | kernel | d2h overlapped).
def _host_f(arr: np.ndarray, name: str) -> np.ndarray:
    if not isinstance(arr, np.ndarray) or not (arr.flags.f_contiguous or arr.ndim == 1):
        raise _lib.ArgumentError(f"{name} must be a column-major (Fortran-ordered) host array")
    return arr


def naive_transpose_host_(backend: B200Backend, a: np.ndarray, b: np.ndarray):
    """reference examples/naive_transpose.jl:9-21 on host Arrays; stream ordered (KA.synchronize(backend) completes it)."""
    _host_f(a, "a"), _host_f(b, "b")
    if a.shape[0] != b.shape[1] or a.shape[1] != b.shape[0]:
        print("Matrix size mismatch!")
        return None
    if a.dtype != b.dtype or a.dtype.itemsize not in (1, 2, 4, 8, 16):
        raise _lib.ArgumentError("naive_transpose_host_: element types must match and be 1, 2, 4, 8 or 16 bytes wide")
    m, n = a.shape
    _lib.call("b200_transpose_host", C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), m, n, m, n, a.dtype.itemsize)
    return None


def histogram_host_(backend: B200Backend, histogram_output: np.ndarray, input: np.ndarray) -> None:
    """reference examples/histogram.jl:60-66 on host Arrays (Int32 values, accumulates into histogram_output)."""
    if histogram_output.dtype != np.int32 or input.dtype != np.int32:
        raise _lib.ArgumentError("histogram_host_: Int32 input and bins only")
    _lib.call("b200_histogram_i32_host", C.c_void_p(histogram_output.ctypes.data), histogram_output.size,
              C.c_void_p(np.ascontiguousarray(input).ctypes.data), input.size)


def performant_matmul_host_(backend: B200Backend, C_: np.ndarray, A: np.ndarray, B: np.ndarray, mode: int = _lib.B200_MATMUL_TILE16) -> None:
    """reference examples/performant_matmul.jl:79-92 on host Arrays (Float32, column-major): B is uploaded once, row panels of
    A stream in and row panels of C stream out while the panels in between are multiplied; stream ordered."""
    for name, x in (("C", C_), ("A", A), ("B", B)):
        _host_f(x, name)
        if x.dtype != np.float32 or x.ndim != 2:
            raise _lib.ArgumentError("performant_matmul_host_: Float32 matrices only")
    M, K = A.shape
    N = B.shape[1]
    if B.shape[0] != K or C_.shape != (M, N):
        raise _lib.ArgumentError("performant_matmul_host_: shapes do not conform")
    _lib.call("b200_matmul_f32_host", C.c_void_p(C_.ctypes.data), C.c_void_p(A.ctypes.data), C.c_void_p(B.ctypes.data), M, K, N, M, K, M,
              C.c_int32(mode))
