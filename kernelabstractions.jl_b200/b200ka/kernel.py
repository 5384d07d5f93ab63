"""KA.Kernel and KI.Kernel callables for B200Backend: the launch path.

Mirrors (reference):
  * `@kernel` constructors  f(dev), f(dev, size), f(dev, size, range)        src/macros.jl:38-48
  * KA.launch_config / mkcontext / (obj::KA.Kernel{B})(args...; ndrange, workgroupsize)   src/pocl/backend.jl:74-147
  * KI.kernel_function / (obj::KI.Kernel{B})(args...; numworkgroups, workgroupsize)        src/pocl/backend.jl:149-168
  * KI.check_launch_args                                                      src/intrinsics.jl:182-188
There is no JIT: the kernel *name* selects a precompiled sm_100a kernel inside libb200ka.so (b200_launch); an
unknown name raises ErrorException.  Launches are asynchronous on the calling thread's stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ArgumentError, ErrorException, call, lib
from .backend import B200Array, B200Backend, eltype_code
from .ndrange import (DYNAMIC, DynamicCheck, DynamicSize, NDRange, StaticSize, _as_tuple, ka_partition,
                      threads_to_workgroupsize)

MAX_WORK_GROUP_SIZE = 1024


class Val:
    """Julia's Val{x}() -- a compile-time constant argument (e.g. Val(groupsize), Val(TILE_DIM))."""

    def __init__(self, value):
        self.value = value

    def __repr__(self):
        return f"Val({self.value})"


def argconvert(arg) -> _lib.Arg:
    """KA.argconvert / KI.argconvert (reference src/pocl/backend.jl:149,242): host value -> C argument record."""
    a = _lib.Arg()
    if isinstance(arg, B200Array):
        if arg.ndim > _lib.B200_MAX_NDIM:
            raise ArgumentError(f"arrays of more than {_lib.B200_MAX_NDIM} dimensions are not supported")
        a.kind = _lib.B200_ARG_ARRAY
        a.eltype = eltype_code(arg.dtype)
        a.elsize = arg.dtype.itemsize
        a.ndim = arg.ndim
        for d, s in enumerate(arg.shape):
            a.dims[d] = s
        a.ptr = arg.ptr
        return a
    if isinstance(arg, np.ndarray):
        raise ErrorException("host arrays cannot be passed to a B200Backend kernel; move them with KA.copyto!/adapt")
    if isinstance(arg, Val):
        a.kind = _lib.B200_ARG_VAL
        v = arg.value
    else:
        a.kind = _lib.B200_ARG_SCALAR
        v = arg
    if isinstance(v, (type, np.dtype)):
        a.eltype = eltype_code(v)
        a.scalar = a.eltype
    elif isinstance(v, (bool, np.bool_)):
        a.eltype = _lib.B200_T_BOOL
        a.scalar = int(v)
        a.fscalar = float(v)
    elif isinstance(v, (int, np.integer)):
        a.eltype = _lib.B200_T_I64
        a.scalar = int(v)
        a.fscalar = float(v)
    elif isinstance(v, (float, np.floating)):
        a.eltype = _lib.B200_T_F32 if isinstance(v, np.float32) else _lib.B200_T_F64
        a.fscalar = float(v)
        a.scalar = int(v) if float(v).is_integer() else 0
    else:
        raise ErrorException(f"Don't know how to convert arguments of type {type(v).__name__} for Kernel{{B200Backend}}")
    return a


def _launch(name: str, desc: _lib.LaunchDesc, args: Sequence) -> None:
    cargs = (_lib.Arg * max(len(args), 1))(*[argconvert(a) for a in args])
    st = lib().b200_launch(name.encode(), C.byref(desc), cargs, C.c_int32(len(args)))
    if st == _lib.B200_ERR_UNKNOWN_KERNEL:
        raise ErrorException(_lib.last_error())
    _lib.check(st)


class Kernel:
    """KA.Kernel{B200Backend, WorkgroupSize, NDRange, Fun} (reference src/KernelAbstractions.jl:735-746)."""

    def __init__(self, backend, name: str, workgroupsize=DYNAMIC, ndrange=DYNAMIC):
        self.backend = backend
        self.name = name          # nameof(obj.f) == "gpu_" * kernel function name (src/macros.jl:32)
        self.workgroupsize = workgroupsize
        self.ndrange = ndrange

    def launch_config(self, ndrange, workgroupsize):
        """KA.launch_config (reference src/pocl/backend.jl:84-106)."""
        ndrange = _as_tuple(ndrange)
        workgroupsize = _as_tuple(workgroupsize)
        if isinstance(self.ndrange, StaticSize):
            # "partition checked that the ndrange's agreed" (:93-96): keep that check, then drop the runtime value
            if ndrange is not None and ndrange != self.ndrange.S:
                raise ErrorException(f"Static NDRange ({self.ndrange}) and launch NDRange ({ndrange}) differ")
            ndrange = None
        if isinstance(self.workgroupsize, DynamicSize) and workgroupsize is None:
            # use ndrange as preliminary workgroupsize for autotuning (:98-101)
            iterspace, dynamic = ka_partition(self.ndrange, self.workgroupsize, ndrange, ndrange)
        else:
            iterspace, dynamic = ka_partition(self.ndrange, self.workgroupsize, ndrange, workgroupsize)
        return ndrange, workgroupsize, iterspace, dynamic

    def mkcontext(self, ndrange, iterspace: NDRange) -> _lib.LaunchDesc:
        """KA.mkcontext (reference src/pocl/backend.jl:74-76): the flattened CompilerMetadata.  Like the in-tree
        backend the context is always built with DynamicCheck."""
        full_ndrange = ndrange if ndrange is not None else self.ndrange.S
        d = _lib.LaunchDesc()
        d.kind = _lib.B200_LAUNCH_KA
        d.ndim = iterspace.N
        for i in range(_lib.B200_MAX_NDIM):
            inside = i < iterspace.N
            d.ndrange[i] = full_ndrange[i] if inside else 1
            d.wgs[i] = iterspace.workitems[i] if inside else 1
            d.blocks[i] = iterspace.blocks[i] if inside else 1
        d.dynamic_check = 1
        d.flags = (1 if iterspace.static_blocks else 0) | (2 if iterspace.static_workitems else 0)
        return d

    def __call__(self, *args, ndrange=None, workgroupsize=None):
        """(obj::KA.Kernel{B200Backend})(args...; ndrange, workgroupsize) (reference src/pocl/backend.jl:117-147)."""
        nd, wgs, iterspace, dynamic = self.launch_config(ndrange, workgroupsize)
        if iterspace.N > _lib.B200_MAX_NDIM:
            raise ArgumentError(f"ndranges of more than {_lib.B200_MAX_NDIM} dimensions are not supported")
        if isinstance(self.workgroupsize, DynamicSize) and wgs is None:
            # figure out the workgroupsize automatically (:126-131): greedy fill of the device maximum
            full = nd if nd is not None else self.ndrange.S
            wg_size_nd = threads_to_workgroupsize(MAX_WORK_GROUP_SIZE, full)
            iterspace, dynamic = ka_partition(self.ndrange, self.workgroupsize, nd, wg_size_nd)
        if len(iterspace) == 0:
            return None  # groups == 0 (:136-138)
        if iterspace.nitems() > MAX_WORK_GROUP_SIZE:
            raise ArgumentError(f"workgroupsize {iterspace.workitems} exceeds {MAX_WORK_GROUP_SIZE} work-items")
        _launch(self.name, self.mkcontext(nd, iterspace), args)
        return None


class KernelFunction:
    """What `@kernel function name(...) end` leaves behind: a constructor with the four methods of
    reference src/macros.jl:38-48."""

    def __init__(self, name: str):
        self.fname = name
        self.gpu_name = "gpu_" + name

    def __call__(self, backend, workgroupsize=None, ndrange=None) -> Kernel:
        if not isinstance(backend, B200Backend):
            raise ErrorException(f"kernel {self.fname} is only precompiled for B200Backend")
        ws = DYNAMIC if workgroupsize is None else (workgroupsize if isinstance(workgroupsize, (StaticSize, DynamicSize))
                                                    else StaticSize(_as_tuple(workgroupsize)))
        nr = DYNAMIC if ndrange is None else (ndrange if isinstance(ndrange, (StaticSize, DynamicSize))
                                              else StaticSize(_as_tuple(ndrange)))
        return Kernel(backend, self.gpu_name, ws, nr)


# ---- KernelIntrinsics --------------------------------------------------------------------------------------
def check_launch_args(numworkgroups, workgroupsize) -> None:
    """KI.check_launch_args (reference src/intrinsics.jl:182-188)."""
    if len(numworkgroups) > 3:
        raise ArgumentError("`numworkgroups` only accepts up to 3 dimensions")
    if len(workgroupsize) > 3:
        raise ArgumentError("`workgroupsize` only accepts up to 3 dimensions")


class KIKernel:
    """KI.Kernel{B200Backend, Kern} (reference src/intrinsics.jl:177-180)."""

    def __init__(self, backend, name: str):
        self.backend = backend
        self.name = name

    def __call__(self, *args, numworkgroups=1, workgroupsize=1):
        """(obj::KI.Kernel{B200Backend})(args...; numworkgroups, workgroupsize) (reference src/pocl/backend.jl:156-168):
        a true <=3-D launch, no implicit bounds check."""
        ng, ws = _as_tuple(numworkgroups), _as_tuple(workgroupsize)
        check_launch_args(ng, ws)
        d = _lib.LaunchDesc()
        d.kind = _lib.B200_LAUNCH_KI
        d.ndim = max(len(ng), len(ws))
        for i in range(_lib.B200_MAX_NDIM):
            d.wgs[i] = ws[i] if i < len(ws) else 1
            d.blocks[i] = ng[i] if i < len(ng) else 1
            d.ndrange[i] = d.wgs[i] * d.blocks[i]
        _launch(self.name, d, args)
        return None


def ki_kernel(backend, f_name: str, *args, numworkgroups=None, workgroupsize=None, launch: bool = True):
    """`KI.@kernel backend [launch=false] numworkgroups=... workgroupsize=... f(args...)`
    (reference src/intrinsics.jl:300-373)."""
    if not isinstance(backend, B200Backend):
        raise ErrorException("KI.@kernel: B200Backend expected")
    k = kernel_function(backend, f_name)
    if not launch:
        return k
    kw = {}
    if numworkgroups is not None:
        kw["numworkgroups"] = numworkgroups
    if workgroupsize is not None:
        kw["workgroupsize"] = workgroupsize
    k(*args, **kw)
    return k


def kernel_function(backend, f_name: str) -> KIKernel:
    """KI.kernel_function (reference src/pocl/backend.jl:151-154): look the precompiled kernel up by name."""
    out = C.c_int64(0)
    st = lib().b200_kernel_max_work_group_size(f_name.encode(), C.c_int64(MAX_WORK_GROUP_SIZE), C.byref(out))
    if st == _lib.B200_ERR_UNKNOWN_KERNEL:
        raise ErrorException(_lib.last_error())
    _lib.check(st)
    return KIKernel(backend, f_name)


def kernel_max_work_group_size(kernel, max_work_items: int = 2 ** 63 - 1) -> int:
    """KI.kernel_max_work_group_size (reference src/pocl/backend.jl:170-173; test/intrinsics.jl:95-96)."""
    out = C.c_int64(0)
    call("b200_kernel_max_work_group_size", kernel.name.encode(), C.c_int64(max_work_items), C.byref(out))
    return int(out.value)


def max_work_group_size(backend) -> int:
    """KI.max_work_group_size (reference src/pocl/backend.jl:174-176)."""
    v = C.c_int32(0)
    call("b200_max_threads_per_block", C.byref(v))
    return int(v.value)


def multiprocessor_count(backend) -> int:
    """KI.multiprocessor_count (reference src/pocl/backend.jl:177-179)."""
    v = C.c_int32(0)
    call("b200_sm_count", C.byref(v))
    return int(v.value)


def registered_kernels():
    n = C.c_int32(0)
    call("b200_kernel_count", C.byref(n))
    return [lib().b200_kernel_name(i).decode() for i in range(n.value)]


def register_kernel(name: str, handler, kind: str = "ka", user=None) -> None:
    """Add a precompiled kernel from the caller's own shared library to the catalogue (b200_register_kernel).
    `name` is what b200_launch is called with: "gpu_" + function name for `@kernel`s (reference src/macros.jl:32), the
    function name itself for `KI.@kernel`.  `handler` is the address of a native `b200_kernel_handler` (an int, a
    ctypes function pointer, or a ctypes CFUNCTYPE object the caller keeps alive)."""
    addr = handler if isinstance(handler, int) else C.cast(handler, C.c_void_p).value
    call("b200_register_kernel", name.encode(), C.c_int32(_lib.B200_LAUNCH_KA if kind == "ka" else _lib.B200_LAUNCH_KI),
         C.c_void_p(addr), C.c_void_p(user))


def unregister_kernel(name: str) -> None:
    st = lib().b200_unregister_kernel(name.encode())
    if st == _lib.B200_ERR_UNKNOWN_KERNEL:
        raise ErrorException(_lib.last_error())
    _lib.check(st)
