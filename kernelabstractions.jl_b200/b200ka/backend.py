"""B200Backend and B200Array: the Python mirror of `struct B200Backend <: KA.GPU` (julia/B200Backend).

Method-for-method twin of the in-tree backend glue (reference src/pocl/backend.jl:19-69) on top of the C ABI.
Julia's `f!` is spelled `f_` here.  All device work is asynchronous and ordered on the calling thread's stream;
`synchronize` is cooperative (it polls and yields instead of blocking in the driver), as the reference requires
of GPU backends (docs/src/implementations.md:3-11).
"""
from __future__ import annotations

import ctypes as C
import threading
import time
import weakref
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ArgumentError, ErrorException, MethodError, call, lib

_DTYPE_CODES = {
    np.dtype(np.int8): _lib.B200_T_I8, np.dtype(np.uint8): _lib.B200_T_U8,
    np.dtype(np.int16): _lib.B200_T_I16, np.dtype(np.uint16): _lib.B200_T_U16,
    np.dtype(np.int32): _lib.B200_T_I32, np.dtype(np.uint32): _lib.B200_T_U32,
    np.dtype(np.int64): _lib.B200_T_I64, np.dtype(np.uint64): _lib.B200_T_U64,
    np.dtype(np.float16): _lib.B200_T_F16, np.dtype(np.float32): _lib.B200_T_F32,
    np.dtype(np.float64): _lib.B200_T_F64, np.dtype(np.bool_): _lib.B200_T_BOOL,
}


def eltype_code(dtype) -> int:
    return _DTYPE_CODES.get(np.dtype(dtype), _lib.B200_T_OPAQUE)


def CartesianIndex(n: int) -> np.dtype:
    """eltype of an array of CartesianIndex{n}: n consecutive Int64 (reference test/test.jl:131-139)."""
    return np.dtype([("I", np.int64, (n,))])


KernelData = np.dtype([(k, np.int64) for k in
                       ("global_size", "global_id", "local_size", "local_id", "num_groups", "group_id")])
"""reference test/intrinsics.jl:3-10"""


class B200Backend:
    """`struct B200Backend <: KA.GPU end` -- stateless, all instances compare equal (reference src/pocl/backend.jl:19-20)."""

    def __eq__(self, other):
        return isinstance(other, B200Backend)

    def __hash__(self):
        return hash("B200Backend")

    def __repr__(self):
        return "B200Backend()"


class CPU:
    """Stand-in for KA.CPU() used only as the target of `adapt(CPU(), x)` (reference test/test.jl:107-113)."""

    def __eq__(self, other):
        return isinstance(other, CPU)

    def __hash__(self):
        return hash("CPU")


# Host buffers of in-flight async copies (caller-preserve rule, KernelAbstractions.jl:127-142), keyed by the stream the copy
# was enqueued on: synchronize() waits for ONE stream (the calling thread's, on its current device) and releases only that
# stream's references -- a copy still in flight on another thread's stream or on another device keeps its buffer.
_pending_host_refs: dict = {}
_pending_lock = threading.Lock()


def _stream_key() -> int:
    s = C.c_void_p()
    call("b200_get_stream", C.byref(s))
    return s.value or 0


def _hold(ref) -> None:
    key = _stream_key()
    with _pending_lock:
        _pending_host_refs.setdefault(key, []).append(ref)


class B200Array:
    """Device array: pointer + dims (column-major) + eltype, owner of its allocation (finalizer -> b200_free).

    The Julia twin is `mutable struct B200Array{T,N} <: AbstractArray{T,N}`."""

    __slots__ = ("ptr", "shape", "dtype", "unified", "_owner", "_freed")

    def __init__(self, dtype, shape: Sequence[int], unified: bool = False, _ptr: Optional[int] = None, _owner=None):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        if any(s < 0 for s in self.shape):
            raise ArgumentError("invalid Array dimensions")
        self.unified = bool(unified)
        self._owner = _owner
        self._freed = False
        if _ptr is not None:
            self.ptr = _ptr
        else:
            p = C.c_void_p()
            call("b200_malloc", C.byref(p), C.c_size_t(self.nbytes), C.c_int32(1 if unified else 0))
            self.ptr = p.value or 0

    # -- Base array interface -------------------------------------------------------------------
    @property
    def size(self) -> int:  # length(A)
        n = 1
        for s in self.shape:
            n *= s
        return n

    def __len__(self) -> int:
        return self.size

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    def __getitem__(self, idx):
        """Scalar indexing: ErrorException unless the array is unified memory (reference test/test.jl:82-87)."""
        if not self.unified:
            raise ErrorException("Scalar indexing of a B200Array is disallowed; copy it to the host with Array(A)")
        synchronize(B200Backend())
        return Array(self)[idx]

    def __del__(self):
        try:
            unsafe_free_(self)
        except Exception:
            pass

    def __repr__(self):
        return f"B200Array({self.dtype}, {self.shape}, ptr=0x{self.ptr:x})"

    def view(self, offset_elems: int, shape: Sequence[int]) -> "B200Array":
        """A contiguous sub-array sharing memory (keeps the parent alive), e.g. one slab of a sharded matrix."""
        v = B200Array(self.dtype, shape, self.unified, _ptr=self.ptr + offset_elems * self.dtype.itemsize, _owner=self)
        assert offset_elems + v.size <= self.size
        return v

    def fill_(self, value) -> "B200Array":
        """fill!(A, x) / `A .= x` (reference test/private.jl:75), on the device."""
        v = np.array([value], dtype=self.dtype)
        if self.size:
            call("b200_fill", C.c_void_p(self.ptr), C.c_size_t(self.size), C.c_int32(self.dtype.itemsize),
                 v.ctypes.data_as(C.c_void_p))
        return self


def unsafe_free_(A: B200Array) -> None:
    """KA.unsafe_free! (reference src/KernelAbstractions.jl:193-195)"""
    if isinstance(A, B200Array) and not A._freed and A._owner is None and A.ptr:
        A._freed = True
        lib().b200_free(C.c_void_p(A.ptr))
        A.ptr = 0


# ---- memory operations (reference src/pocl/backend.jl:25-57) ---------------------------------------
def _dims(dims) -> Tuple[int, ...]:
    if isinstance(dims, (int, np.integer)):
        return (int(dims),)
    return tuple(int(d) for d in dims)


def allocate(backend, T, *dims, unified: bool = False) -> B200Array:
    """KA.allocate(::B200Backend, T, dims; unified) (reference src/pocl/backend.jl:25).
    A non-type first argument is a MethodError (reference test/test.jl:300-305)."""
    if not isinstance(backend, B200Backend):
        raise MethodError(f"no method matching allocate(::{type(backend).__name__}, ...)")
    try:
        dt = np.dtype(T)
    except TypeError as e:
        raise MethodError(f"no method matching allocate(::B200Backend, ::{type(T).__name__}, ...)") from e
    if isinstance(T, (np.ndarray, B200Array, list, tuple)) and not isinstance(T, type):
        raise MethodError("allocate: eltype must be a type")
    shape = _dims(dims[0]) if len(dims) == 1 and not isinstance(dims[0], (int, np.integer)) else _dims(dims)
    return B200Array(dt, shape, unified=unified)


def zeros(backend, T, *dims, **kw) -> B200Array:
    """KA.zeros = allocate + init_kernel(arr, zero, T) (reference src/pocl/backend.jl:27-32)."""
    return allocate(backend, T, *dims, **kw).fill_(0)


def ones(backend, T, *dims, **kw) -> B200Array:
    """KA.ones (reference src/pocl/backend.jl:33-38)."""
    return allocate(backend, T, *dims, **kw).fill_(1)


def _host_ptr(a: np.ndarray):
    """Device arrays are column-major (Julia layout); a host array's bytes are copied as they lie, so a matrix must be
    Fortran-ordered (vectors are both)."""
    if not a.flags.f_contiguous:
        raise ArgumentError("host arrays of more than one dimension must be column-major (np.asfortranarray): device arrays "
                            "are column-major and copyto! moves raw bytes")
    return a.ctypes.data_as(C.c_void_p)


def copyto_(backend, A, B):
    """KA.copyto!(backend, A, B) (reference src/pocl/backend.jl:40-54): device<-device through the copy kernel
    (with the length and alias checks), host<->device as stream-ordered async copies.  Returns A."""
    a_dev, b_dev = isinstance(A, B200Array), isinstance(B, B200Array)
    if a_dev and b_dev:
        if A.size != B.size:
            raise ErrorException("Arrays must match in length")
        if A.dtype.itemsize != B.dtype.itemsize:
            raise ErrorException("Arrays must match in element size")
        try:
            call("b200_copy", C.c_void_p(A.ptr), C.c_void_p(B.ptr), C.c_size_t(A.nbytes))
        except _lib.B200Error as e:
            if e.status == _lib.B200_ERR_ALIAS:
                raise ErrorException("Arrays may not alias") from e
            raise
        return A
    if a_dev and isinstance(B, np.ndarray):
        if A.size != B.size:
            raise ErrorException("Arrays must match in length")
        # element (i, j, ...) of the host array must land on element (i, j, ...) of the column-major device array: a C-ordered
        # matrix (numpy's default) is re-laid out first; the temporary stays alive until the stream has been waited for
        src = B if B.dtype == A.dtype else B.astype(A.dtype)
        if not src.flags.f_contiguous:
            src = np.asfortranarray(src)
        call("b200_memcpy_h2d_async", C.c_void_p(A.ptr), _host_ptr(src), C.c_size_t(A.nbytes))
        _hold(src)
        return A
    if isinstance(A, np.ndarray) and b_dev:
        if A.size != B.size:
            raise ErrorException("Arrays must match in length")
        if A.dtype != B.dtype:
            raise ErrorException("copyto!: host destination must have the device array's eltype")
        if not A.flags.f_contiguous:
            raise ArgumentError("copyto!: a host destination of more than one dimension must be column-major "
                                "(np.asfortranarray / order='F'); the device array's bytes are copied as they lie")
        call("b200_memcpy_d2h_async", _host_ptr(A), C.c_void_p(B.ptr), C.c_size_t(B.nbytes))
        _hold(A)
        return A
    raise MethodError("copyto!(::B200Backend, A, B): at least one side must be a B200Array")


def synchronize(backend) -> None:
    """KA.synchronize (reference src/KernelAbstractions.jl:176): cooperative wait -- poll b200_stream_query and yield
    to other host threads, never block inside the driver (docs/src/implementations.md:3-11)."""
    if not isinstance(backend, B200Backend):
        raise MethodError(f"no method matching synchronize(::{type(backend).__name__})")
    spins = 0
    while True:
        st = lib().b200_stream_query()
        if st == _lib.B200_OK:
            break
        if st != _lib.B200_NOT_READY:
            _lib.check(st)
        spins += 1
        time.sleep(0 if spins < 2000 else 50e-6)
    key = _stream_key()
    with _pending_lock:
        _pending_host_refs.pop(key, None)      # only the stream that was waited for


def get_backend(A):
    """KA.get_backend (reference src/KernelAbstractions.jl:548-557, src/pocl/backend.jl:59)."""
    if isinstance(A, B200Array):
        return B200Backend()
    if isinstance(A, np.ndarray):
        return CPU()
    raise ArgumentError(f"Implement KernelAbstractions.get_backend(::{type(A).__name__})")


def functional(backend) -> bool:
    n = C.c_int32(0)
    return lib().b200_device_count(C.byref(n)) == 0 and n.value > 0


def pagelock_(backend, x: np.ndarray) -> None:
    """KA.pagelock! (reference src/KernelAbstractions.jl:166)"""
    if not (x.flags.f_contiguous or x.flags.c_contiguous):
        raise ArgumentError("pagelock!: host arrays must be contiguous")
    addr = x.ctypes.data
    call("b200_host_register", C.c_void_p(addr), C.c_size_t(x.nbytes))
    # A registration that outlives the array poisons the address range: the allocator hands the pages to the next array, and a
    # copy from a buffer that lies PARTLY inside a stale registration fails with "invalid argument" (seen as a flaky failure of
    # whichever test came later).  Julia's pagelock! leaves this to the GC-owned Array; here the registration is dropped when the
    # array that owns the memory is collected.
    owner = x
    while isinstance(getattr(owner, "base", None), np.ndarray):
        owner = owner.base
    if addr not in _pagelocked:
        _pagelocked[addr] = weakref.finalize(owner, _unregister, addr)


def _unregister(addr: int) -> None:
    _pagelocked.pop(addr, None)
    try:
        lib().b200_host_unregister(C.c_void_p(addr))
    except Exception:   # interpreter shutdown
        pass


_pagelocked: dict = {}


def unpagelock_(backend, x: np.ndarray) -> None:
    """Undo pagelock_ explicitly (it also happens when the array that owns the memory is collected)."""
    fin = _pagelocked.pop(x.ctypes.data, None)
    if fin is not None:
        fin.detach()
    call("b200_host_unregister", x.ctypes.data_as(C.c_void_p))


def supports_float64(backend) -> bool:
    return True


def supports_atomics(backend) -> bool:
    return True


def supports_unified(backend) -> bool:
    return True


def priority_(backend, prio: str) -> None:
    """KA.priority!(backend, :high | :normal | :low) (reference src/KernelAbstractions.jl:658-663)."""
    if prio not in ("high", "normal", "low"):
        raise ErrorException(f"priority must be one of :high, :normal, :low (got {prio!r})")
    call("b200_set_priority", C.c_int32({"low": -1, "normal": 0, "high": 1}[prio]))


def ndevices(backend) -> int:
    n = C.c_int32(0)
    call("b200_device_count", C.byref(n))
    return n.value


def device(backend) -> int:
    """1-based like KA.device (reference src/KernelAbstractions.jl:670-680)."""
    d = C.c_int32(0)
    call("b200_get_device", C.byref(d))
    return d.value + 1


def device_(backend, id: int) -> None:
    """KA.device!(backend, id): ArgumentError when out of range (reference src/KernelAbstractions.jl:686-691)."""
    if id < 1 or id > ndevices(backend):
        raise ArgumentError(f"Device id {id} out of bounds.")
    call("b200_set_device", C.c_int32(id - 1))


# ---- conversions the testsuite relies on (reference test/test.jl:107-113, test/private.jl:75-91) -----------
def Array(A: B200Array) -> np.ndarray:
    """Array(::B200Array): synchronous device->host copy, column-major numpy array."""
    out = np.empty(A.shape, dtype=A.dtype, order="F")
    if A.size:
        call("b200_memcpy_d2h_async", out.ctypes.data_as(C.c_void_p), C.c_void_p(A.ptr), C.c_size_t(A.nbytes))
        call("b200_synchronize")
    return out


def to_device(host: np.ndarray, unified: bool = False) -> B200Array:
    """B200Array(::Array): synchronous host->device copy (column-major)."""
    h = np.asfortranarray(host)
    A = B200Array(h.dtype, h.shape, unified=unified)
    if A.size:
        call("b200_memcpy_h2d_async", C.c_void_p(A.ptr), h.ctypes.data_as(C.c_void_p), C.c_size_t(A.nbytes))
        call("b200_synchronize")
    return A


def similar(A: B200Array, dtype=None, shape=None) -> B200Array:
    return B200Array(A.dtype if dtype is None else dtype, A.shape if shape is None else shape, unified=A.unified)


def adapt(backend, x):
    """Adapt.adapt_storage for the three cases the testsuite uses (reference test/test.jl:107-113)."""
    if isinstance(backend, CPU):
        return Array(x) if isinstance(x, B200Array) else x
    if isinstance(backend, B200Backend):
        return x if isinstance(x, B200Array) else to_device(x)
    raise MethodError("adapt: unknown backend")


# ---- device-array conveniences: what Base / LinearAlgebra give a CPU() user on plain Arrays -----------
_RESULT_KIND = {"i": np.int64, "u": np.uint64, "f": np.float64}


def _reduce(A: B200Array, op: int):
    code = eltype_code(A.dtype)
    out = np.zeros(1, dtype=(_RESULT_KIND[A.dtype.kind] if op == 0 else A.dtype))
    call("b200_reduce", C.c_int32(op), C.c_int32(code), C.c_void_p(A.ptr), C.c_size_t(A.size), out.ctypes.data_as(C.c_void_p))
    return out[0]


def sum(A: B200Array):  # noqa: A001 -- Base.sum
    """sum(A): integers are carried in Int64/UInt64 like Julia's widening, floats in Float64."""
    return _reduce(A, 0)


def maximum(A: B200Array):
    """maximum(A) on the device (reference examples/histogram.jl:10,88-90); empty -> ArgumentError like Julia."""
    if A.size == 0:
        raise ArgumentError("reducing over an empty collection is not allowed")
    return _reduce(A, 1)


def minimum(A: B200Array):
    if A.size == 0:
        raise ArgumentError("reducing over an empty collection is not allowed")
    return _reduce(A, 2)


def isequal(A: B200Array, B: B200Array) -> bool:
    """`A == B` for device arrays (reference examples/memcopy.jl:23, test/copyto.jl:17-19): same shape and every
    element equal (floats numerically: NaN != NaN, -0.0 == 0.0)."""
    if tuple(A.shape) != tuple(B.shape):
        return False
    if A.dtype != B.dtype:
        raise MethodError("isequal: element types differ")
    bad = np.zeros(1, dtype=np.int64)
    call("b200_count_mismatch", C.c_void_p(A.ptr), C.c_void_p(B.ptr), C.c_size_t(A.size), C.c_int32(eltype_code(A.dtype)),
         C.c_int32(A.dtype.itemsize), bad.ctypes.data_as(C.c_void_p))
    return int(bad[0]) == 0


def isapprox(A: B200Array, B: B200Array, rtol=None, atol: float = 0.0) -> bool:
    """isapprox(A, B) for Float32/Float64 device arrays (reference examples/matmul.jl:36): norm(A - B) <=
    max(atol, rtol * max(norm(A), norm(B))) with Julia's default rtol = sqrt(eps(T)) when atol == 0."""
    if tuple(A.shape) != tuple(B.shape) or A.dtype != B.dtype or A.dtype.kind != "f" or A.dtype.itemsize not in (4, 8):
        raise MethodError("isapprox: two Float32 or two Float64 arrays of the same shape")
    if rtol is None:
        rtol = float(np.sqrt(np.finfo(A.dtype).eps)) if atol == 0.0 else 0.0
    out = np.zeros(3, dtype=np.float64)
    call("b200_norms", C.c_void_p(A.ptr), C.c_void_p(B.ptr), C.c_size_t(A.size), C.c_int32(A.dtype.itemsize == 8),
         out.ctypes.data_as(C.c_void_p))
    nd, nx, ny = np.sqrt(out)
    return bool(nd <= max(atol, rtol * max(nx, ny)))


def rand_uniform_(A: B200Array, seed: int) -> B200Array:
    """Fill a Float32 device array with the deterministic uniform [0,1) stream of SURVEY.md 8(d) (same bits as the
    oracle's generator) -- the device-side stand-in for `rand(Float32, dims)` + copyto!."""
    if A.dtype != np.float32:
        raise MethodError("rand_uniform_: Float32 arrays only")
    call("b200_rand_uniform_f32", C.c_void_p(A.ptr), C.c_size_t(A.size), C.c_uint64(seed))
    return A


def rand_bins_(A: B200Array, seed: int, nbins: int) -> B200Array:
    """Fill an Int32 device array with values uniform in 1..nbins (reference examples/histogram.jl:76)."""
    if A.dtype != np.int32:
        raise MethodError("rand_bins_: Int32 arrays only")
    call("b200_rand_bins_i32", C.c_void_p(A.ptr), C.c_size_t(A.size), C.c_uint64(seed), C.c_int32(nbins))
    return A


def mul_(C_: B200Array, A: B200Array, B: B200Array) -> B200Array:
    """LinearAlgebra.mul!(C, A, B) for Float32 device matrices -> the FP32 SIMT GEMM (k ascending, fused multiply-add)."""
    if A.ndim != 2 or B.ndim != 2 or A.shape[1] != B.shape[0] or tuple(C_.shape) != (A.shape[0], B.shape[1]):
        raise _lib.ErrorException("DimensionMismatch: mul_")
    if not (A.dtype == B.dtype == C_.dtype == np.float32):
        raise MethodError("mul_: Float32 matrices only")
    M, K = A.shape
    N = B.shape[1]
    if M and N:
        if K == 0:
            C_.fill_(0)
        else:
            call("b200_matmul_f32", C.c_void_p(C_.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), M, K, N, M, K, M, _lib.B200_MATMUL_TILE16)
    return C_


def matmul(A: B200Array, B: B200Array) -> B200Array:
    """`A * B` (reference examples/matmul.jl:36 compares the kernel's output with it)."""
    return mul_(B200Array(A.dtype, (A.shape[0], B.shape[1])), A, B)


class Graph:
    """`with KA.Graph(backend) as g: <launches>` records the stream-ordered calls of the block (nothing executes), `g.launch()`
    replays them with one submission.  For launch-bound sequences of small kernels (b200_graph_*)."""

    def __init__(self, backend):
        self.backend = backend
        self.handle = C.c_void_p()

    def __enter__(self):
        call("b200_graph_begin")
        return self

    def __exit__(self, et, ev, tb):
        if et is not None:                       # leave capture mode, drop the partial recording
            h = C.c_void_p()
            _lib.lib().b200_graph_end(C.byref(h))
            if h:
                _lib.lib().b200_graph_destroy(h)
            return False
        call("b200_graph_end", C.byref(self.handle))
        return False

    def launch(self) -> None:
        call("b200_graph_launch", self.handle)

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().b200_graph_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass
