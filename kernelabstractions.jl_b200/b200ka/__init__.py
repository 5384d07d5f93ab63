"""b200ka -- Python mirror of the Julia host layer `B200Backend <: KA.GPU` over libb200ka.so.

The names follow KernelAbstractions (KA) and KernelAbstractions.KernelIntrinsics (KI); a Julia `f!` is `f_`.
Everything here is thin: argument marshalling + the unchanged KA launch logic; all compute happens in
hand-written sm_100a CUDA kernels behind the C ABI declared in include/b200ka.h.
"""
from . import _lib
from ._lib import ArgumentError, B200Error, ErrorException, MethodError, build
from .backend import (CPU, Array, B200Array, B200Backend, CartesianIndex, KernelData, adapt, allocate, copyto_,
                      device, device_, functional, get_backend, ndevices, ones, pagelock_, unpagelock_, priority_, similar,
                      supports_atomics, supports_float64, supports_unified, synchronize, to_device, unsafe_free_,
                      zeros, maximum, minimum, isequal, isapprox, rand_uniform_, rand_bins_, mul_, matmul, Graph)
from .backend import sum as sum  # noqa: A004 -- Base.sum for device arrays (KA.sum)
from .kernel import (KIKernel, Kernel, KernelFunction, Val, check_launch_args, kernel_function,
                     kernel_max_work_group_size, ki_kernel, max_work_group_size, multiprocessor_count,
                     registered_kernels, register_kernel, unregister_kernel)
from .ndrange import (DYNAMIC, DynamicCheck, DynamicSize, NDRange, NoDynamicCheck, StaticSize, expand, ka_partition,
                      nd_partition, threads_to_workgroupsize)
from . import kernels, examples, multigpu  # noqa: E402

__all__ = [n for n in dir() if not n.startswith("_")]
