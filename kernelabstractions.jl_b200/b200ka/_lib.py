"""ctypes binding of libb200ka.so -- the only door from the host layer into native code.

Mirrors how the reference binds libpocl: one `@ccall` per entry point plus a `@checked` wrapper that turns a
non-zero status into an exception (reference src/pocl/nanoOpenCL.jl:22-52,392-408,492-682).  There is no
CPU fallback: if the shared library is missing the import fails loudly, and without a CUDA device every
compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libb200ka.so")
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
B200_MAX_NDIM = 4

# status codes (include/b200ka.h)
B200_OK = 0
B200_NOT_READY = 1
B200_ERR_CUDA = -1
B200_ERR_INVALID_VALUE = -2
B200_ERR_UNKNOWN_KERNEL = -3
B200_ERR_BAD_ARGS = -4
B200_ERR_LAUNCH_DIMS = -5
B200_ERR_LENGTH_MISMATCH = -6
B200_ERR_ALIAS = -7
B200_ERR_UNSUPPORTED = -8
B200_ERR_NO_DEVICE = -9
B200_ERR_OOM = -10
B200_ERR_PARTITION = -11
B200_ERR_BAD_DEVICE = -12
B200_ERR_NCCL = -13

B200_LAUNCH_KA, B200_LAUNCH_KI = 0, 1
B200_ARG_ARRAY, B200_ARG_SCALAR, B200_ARG_VAL = 0, 1, 2
(B200_T_OPAQUE, B200_T_I8, B200_T_U8, B200_T_I16, B200_T_U16, B200_T_I32, B200_T_U32, B200_T_I64, B200_T_U64,
 B200_T_F16, B200_T_BF16, B200_T_F32, B200_T_F64, B200_T_BOOL) = range(14)
B200_MATMUL_TILE16, B200_MATMUL_RUNNING, B200_MATMUL_UNFUSED = 0, 1, 2
B200_UNIQUE_ID_BYTES = 128
B200_IPC_HANDLE_BYTES = 64


class ErrorException(RuntimeError):
    """Julia's ErrorException (what `error(...)` throws)."""


class ArgumentError(ValueError):
    """Julia's ArgumentError."""


class MethodError(TypeError):
    """Julia's MethodError."""


class B200Error(ErrorException):
    """A non-zero status from libb200ka.so (the CLError analogue, reference src/pocl/nanoOpenCL.jl:392-408)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{message} [{status_name(status)}]")
        self.status = status


class LaunchDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("ndim", C.c_int32),
        ("ndrange", C.c_int64 * B200_MAX_NDIM),
        ("wgs", C.c_int64 * B200_MAX_NDIM),
        ("blocks", C.c_int64 * B200_MAX_NDIM),
        ("dynamic_check", C.c_int32),
        ("flags", C.c_int32),
    ]


class Arg(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("eltype", C.c_int32),
        ("elsize", C.c_int32),
        ("ndim", C.c_int32),
        ("dims", C.c_int64 * B200_MAX_NDIM),
        ("ptr", C.c_void_p),
        ("scalar", C.c_int64),
        ("fscalar", C.c_double),
    ]


# every symbol include/b200ka.h declares: name -> (restype, argtypes)
_VP, _I32, _I64, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_PI32, _PI64, _PSZ, _PVP = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_status_name": (C.c_char_p, [_I32]),
    "b200_version": (_I32, []),
    "b200_init": (_I32, []),
    "b200_device_count": (_I32, [_PI32]),
    "b200_set_device": (_I32, [_I32]),
    "b200_get_device": (_I32, [_PI32]),
    "b200_sm_count": (_I32, [_PI32]),
    "b200_max_threads_per_block": (_I32, [_PI32]),
    "b200_mem_info": (_I32, [_PSZ, _PSZ]),
    "b200_device_name": (_I32, [C.c_char_p, _SZ]),
    "b200_malloc": (_I32, [_PVP, _SZ, _I32]),
    "b200_free": (_I32, [_VP]),
    "b200_host_alloc": (_I32, [_PVP, _SZ]),
    "b200_host_free": (_I32, [_VP]),
    "b200_host_register": (_I32, [_VP, _SZ]),
    "b200_host_unregister": (_I32, [_VP]),
    "b200_memcpy_h2d_async": (_I32, [_VP, _VP, _SZ]),
    "b200_memcpy_d2h_async": (_I32, [_VP, _VP, _SZ]),
    "b200_memcpy_d2d_async": (_I32, [_VP, _VP, _SZ]),
    "b200_memset_async": (_I32, [_VP, _I32, _SZ]),
    "b200_stream_query": (_I32, []),
    "b200_synchronize": (_I32, []),
    "b200_device_synchronize": (_I32, []),
    "b200_set_priority": (_I32, [_I32]),
    "b200_get_stream": (_I32, [_PVP]),
    "b200_set_stream": (_I32, [_VP]),
    "b200_stream_create": (_I32, [_PVP, _I32]),
    "b200_stream_destroy": (_I32, [_VP]),
    "b200_bind": (_I32, [_I32, _VP]),
    "b200_event_create": (_I32, [_PVP]),
    "b200_event_record": (_I32, [_VP]),
    "b200_event_synchronize": (_I32, [_VP]),
    "b200_event_elapsed_ms": (_I32, [_VP, _VP, C.POINTER(C.c_float)]),
    "b200_event_destroy": (_I32, [_VP]),
    "b200_launch": (_I32, [C.c_char_p, C.POINTER(LaunchDesc), C.POINTER(Arg), _I32]),
    "b200_memcpy_2d_d2d_async": (_I32, [_VP, _SZ, _VP, _SZ, _SZ, _SZ]),
    "b200_sym_alloc": (_I32, [_PVP, _SZ]),
    "b200_sym_export": (_I32, [_VP, _VP]),
    "b200_sym_connect": (_I32, [_VP, _I32, _I32, _VP]),
    "b200_sym_ptr": (_I32, [_VP, _I32, _PVP]),
    "b200_sym_barrier": (_I32, [_VP]),
    "b200_sym_free": (_I32, [_VP]),
    "b200_transpose_alltoall": (_I32, [_VP, _VP, _I64, _I64, _I32]),
    "b200_graph_begin": (_I32, []),
    "b200_graph_end": (_I32, [_PVP]),
    "b200_graph_launch": (_I32, [_VP]),
    "b200_graph_destroy": (_I32, [_VP]),
    "b200_register_kernel": (_I32, [C.c_char_p, _I32, _VP, _VP]),
    "b200_unregister_kernel": (_I32, [C.c_char_p]),
    "b200_kernel_count": (_I32, [_PI32]),
    "b200_kernel_name": (C.c_char_p, [_I32]),
    "b200_kernel_max_work_group_size": (_I32, [C.c_char_p, _I64, _PI64]),
    "b200_launch_counter": (_I32, [_PI64, _I32]),
    "b200_tune_set": (_I32, [C.c_char_p, _I32]),
    "b200_tune_get": (_I32, [C.c_char_p, _I32, _PI32]),
    "b200_copy": (_I32, [_VP, _VP, _SZ]),
    "b200_fill": (_I32, [_VP, _SZ, _I32, _VP]),
    "b200_transpose": (_I32, [_VP, _VP, _I64, _I64, _I64, _I64, _I32]),
    "b200_histogram_i32": (_I32, [_VP, _I64, _VP, _I64]),
    "b200_histogram_i64": (_I32, [_VP, _I64, _VP, _I64]),
    "b200_saxpy": (_I32, [_VP, C.c_double, _VP, _VP, _SZ, _I32]),
    "b200_matmul_f32": (_I32, [_VP, _VP, _VP, _I64, _I64, _I64, _I64, _I64, _I64, _I32]),
    "b200_matmul_bf16": (_I32, [_VP, _VP, _VP, _I64, _I64, _I64, _I64]),
    "b200_matmul_tf32": (_I32, [_VP, _VP, _VP, _I64, _I64, _I64, _I64, _I64, _I64]),
    "b200_matmul_f32_tf32x3": (_I32, [_VP, _VP, _VP, _I64, _I64, _I64, _I64, _I64, _I64, _VP]),
    "b200_convert_f32_bf16": (_I32, [_VP, _VP, _I64, _I64, _I64, _I32]),
    "b200_matmul_f32_via_bf16": (_I32, [_VP, _VP, _VP, _I64, _I64, _I64, _I64, _I64, _I64, _VP]),
    "b200_matmul_f32_bf16_resident_b": (_I32, [_VP, _VP, _VP, _I64, _I64, _I64, _I64, _I64, _VP]),
    "b200_reduce": (_I32, [_I32, _I32, _VP, _SZ, _VP]),
    "b200_count_mismatch": (_I32, [_VP, _VP, _SZ, _I32, _I32, _VP]),
    "b200_norms": (_I32, [_VP, _VP, _SZ, _I32, _VP]),
    "b200_rand_uniform_f32": (_I32, [_VP, _SZ, C.c_uint64]),
    "b200_rand_bins_i32": (_I32, [_VP, _SZ, C.c_uint64, _I32]),
    "b200_transpose_host": (_I32, [_VP, _VP, _I64, _I64, _I64, _I64, _I32]),
    "b200_matmul_f32_host": (_I32, [_VP, _VP, _VP, _I64, _I64, _I64, _I64, _I64, _I64, _I32]),
    "b200_histogram_i32_host": (_I32, [_VP, _I64, _VP, _I64]),
    "b200_comm_unique_id": (_I32, [_VP]),
    "b200_comm_init_rank": (_I32, [_PVP, _I32, _I32, _VP]),
    "b200_comm_destroy": (_I32, [_VP]),
    "b200_allreduce_sum_i32": (_I32, [_VP, _VP, _I64]),
    "b200_broadcast": (_I32, [_VP, _VP, _SZ, _I32]),
    "b200_sendrecv": (_I32, [_VP, _VP, _I32, _VP, _I32, _SZ]),
    "b200_p2p_alloc": (_I32, [_PVP]),
    "b200_p2p_export": (_I32, [_VP, _VP]),
    "b200_p2p_connect": (_I32, [_VP, _I32, _I32, _VP]),
    "b200_p2p_free": (_I32, [_VP]),
    "b200_sym_connect_ptrs": (_I32, [_VP, _I32, _I32, C.POINTER(C.c_void_p)]),
    "b200_p2p_connect_ptrs": (_I32, [_VP, _I32, _I32, C.POINTER(C.c_void_p)]),
    "b200_p2p_debug_prefill": (_I32, [_VP, _I32, C.POINTER(C.c_uint32), _I64]),
    "b200_p2p_timeline": (_I32, [_VP, C.POINTER(C.c_uint64)]),
    "b200_p2p_debug": (_I32, [_VP, _VP]),
    "b200_histogram_i32_allreduce": (_I32, [_VP, _VP, _I64, _VP, _I64]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libb200ka.so for sm_100a with the Makefile in csrc/ (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    if not verbose:
        cmd.append("-s")
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    """The loaded library.  Raises if it has not been built: there is nothing to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()). "
                "B200Backend has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def status_name(status: int) -> str:
    return lib().b200_status_name(status).decode()


def last_error() -> str:
    return lib().b200_last_error().decode(errors="replace")


def check(status: int) -> None:
    """The `@checked` wrapper: non-zero status -> the exception type the reference's tests expect."""
    if status == B200_OK:
        return
    msg = last_error()
    if status in (B200_ERR_LAUNCH_DIMS, B200_ERR_BAD_DEVICE):
        raise ArgumentError(f"{msg} [{status_name(status)}]")
    raise B200Error(status, msg)


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))
