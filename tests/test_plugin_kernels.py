"""Run-time kernel registration (b200_register_kernel): the catalogue of precompiled kernels is open -- a user ships the CUDA
twin of their `@kernel` in their own shared library and launches it by name through the unchanged KA front end.  (No-JIT
counterpart of KI.kernel_function, reference src/pocl/backend.jl:151-154.)"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))

import b200ka as KA  # noqa: E402
from b200ka import _lib  # noqa: E402

HANDLER = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p)


def test_registration_bookkeeping_without_a_device():
    calls = []
    cb = HANDLER(lambda d, a, n, s, u: calls.append(n) or 0)
    before = KA.registered_kernels()
    assert "gpu_user_noop!" not in before
    KA.register_kernel("gpu_user_noop!", cb)
    assert KA.registered_kernels() == before + ["gpu_user_noop!"]
    with pytest.raises(_lib.B200Error):          # already registered
        KA.register_kernel("gpu_user_noop!", cb)
    with pytest.raises(_lib.B200Error):          # built-in names cannot be overridden
        KA.register_kernel("gpu_copy_kernel!", cb)
    assert KA.kernel_max_work_group_size(KA.KIKernel(KA.B200Backend(), "gpu_user_noop!"), 7) == 7
    KA.unregister_kernel("gpu_user_noop!")
    assert KA.registered_kernels() == before
    with pytest.raises(KA.ErrorException):
        KA.unregister_kernel("gpu_user_noop!")


@pytest.mark.gpu
def test_user_kernels_launch_through_the_ka_front_end(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    so = str(tmp_path / "libscale_plugin.so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "plugin", "scale_plugin.cu")])
    plug = C.CDLL(so)
    be = KA.B200Backend()
    counter = C.c_longlong(0)
    KA.register_kernel("gpu_scale_kernel!", plug.scale_handler, user=C.addressof(counter))
    KA.register_kernel("gpu_rowsum_kernel!", plug.rowsum_handler)
    try:
        scale = KA.KernelFunction("scale_kernel!")
        h = np.arange(1, 1001, dtype=np.float32)
        A = KA.to_device(h)
        scale(be, 64)(A, np.float32(2.5), ndrange=(1000,))            # 1000 is not a multiple of 64: dynamic bounds check
        scale(be)(A, np.float32(2.0), ndrange=(10,))                   # default workgroup size, partial ndrange
        KA.synchronize(be)
        want = h * np.float32(2.5)
        want[:10] *= np.float32(2.0)
        assert np.array_equal(KA.Array(A), want) and counter.value == 2
        rowsum = KA.KernelFunction("rowsum_kernel!")
        M = np.asfortranarray(np.random.default_rng(0).standard_normal((300, 17)))
        out = KA.zeros(be, np.float64, 300)
        rowsum(be, 128)(out, KA.to_device(M), ndrange=(300,))
        KA.synchronize(be)
        ref = np.zeros(300)
        for j in range(17):
            ref += M[:, j]
        assert np.array_equal(KA.Array(out), ref)
        # the handler's own argument check surfaces as an error, and a KI launch of a KA kernel is refused
        with pytest.raises(_lib.B200Error):
            scale(be, 64)(KA.zeros(be, np.float64, 8), np.float32(1), ndrange=(8,))
    finally:
        KA.unregister_kernel("gpu_scale_kernel!")
        KA.unregister_kernel("gpu_rowsum_kernel!")
    with pytest.raises(KA.ErrorException):                             # gone again: unknown kernel -> ErrorException
        KA.KernelFunction("scale_kernel!")(be, 64)(A, np.float32(1), ndrange=(8,))
