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


def matmul_bf16_(C_: B200Array, A: B200Array, B: B200Array, workspace: B200Array = None) -> None:
    """Opt-in tensor-core path (no reference twin): FP32 in/out, BF16 tcgen05 arithmetic."""
    M, K = A.shape
    N = B.shape[1]
    ws = workspace if workspace is not None else allocate(B200Backend(), np.uint16, (M * K + K * N,))
    _lib.call("b200_matmul_f32_via_bf16", C.c_void_p(C_.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), M, K, N, M, K,
              C_.shape[0], C.c_void_p(ws.ptr))
