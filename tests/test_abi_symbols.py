"""CPU-only: the C-ABI library loads, exports every symbol include/b200ka.h declares, and refuses to compute
without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200ka.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)) - {"b200_status", "b200_launch_desc", "b200_arg"})


def test_header_symbols_are_exported():
    import b200ka
    from b200ka import _lib

    handle = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/b200ka.h but not exported by libb200ka.so"
    # the ctypes signature table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_exports_are_plain_c():
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "kernelabstractions.jl_b200", "libb200ka.so")]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in declared_symbols():
        assert n in exported


def test_struct_layout_matches_header():
    from b200ka import _lib

    # b200_launch_desc: 2*int32 + 3*4*int64 + 2*int32 ; b200_arg: 4*int32 + 4*int64 + ptr + int64 + double
    assert C.sizeof(_lib.LaunchDesc) == 8 + 96 + 8
    assert C.sizeof(_lib.Arg) == 16 + 32 + 8 + 8 + 8


def test_version_and_status_names():
    from b200ka import _lib

    assert _lib.lib().b200_version() == 100
    assert _lib.status_name(0) == "B200_OK"
    assert _lib.status_name(-3) == "B200_ERR_UNKNOWN_KERNEL"


def test_registry_lists_the_catalogue():
    import b200ka as KA

    names = set(KA.registered_kernels())
    for k in ["gpu_copy_kernel!", "gpu_copy_kernel", "gpu_init_kernel", "gpu_naive_transpose_kernel!",
              "gpu_matmul_kernel!", "histogram_kernel!", "coalesced_matmul_kernel!", "gpu_localmem", "gpu_localmem2",
              "gpu_localmem_unsafe_indices", "gpu_many_localmem", "gpu_stmt_form", "gpu_typetest", "gpu_private",
              "gpu_forloop", "gpu_reduce_private", "gpu_kernel_unroll!", "gpu_kernel_unroll2!",
              "test_intrinsics_kernel", "launch_kernel1d", "launch_kernel2d", "launch_kernel3d",
              "gpu_simple_copy_kernel!", "gpu_simple_transpose_kernel!", "gpu_lmem_copy_kernel!",
              "gpu_lmem_transpose_kernel!", "gpu_coalesced_copy_kernel!", "gpu_coalesced_transpose_kernel!"]:
        assert k in names


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every compute entry point must fail loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import b200ka as KA
    from b200ka import _lib

    assert KA.functional(KA.B200Backend()) is False
    with pytest.raises(KA.B200Error) as e:
        KA.allocate(KA.B200Backend(), "float32", (4,))
    assert e.value.status == _lib.B200_ERR_NO_DEVICE
    buf = (C.c_float * 4)()
    assert _lib.lib().b200_copy(buf, buf, 0) == 0  # zero bytes: nothing to do
    assert _lib.lib().b200_transpose(buf, buf, 2, 2, 2, 2, 4) == _lib.B200_ERR_NO_DEVICE
    assert _lib.lib().b200_histogram_i32(buf, 4, buf, 4) == _lib.B200_ERR_NO_DEVICE
    assert _lib.lib().b200_matmul_f32(buf, buf, buf, 2, 2, 2, 2, 2, 2, 0) == _lib.B200_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "kernelabstractions.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "ka_oracle" not in text and "libka_oracle" not in text, f"{f} references the oracle"
