"""GPU parity: the CUDA hot path (through the C ABI / host mirror) against the CPU oracle on the same seeded inputs.

Bar: bit-exact for copy / fill / transpose / histogram (integer, byte and index work) and for the FP32 GEMM in the
reference's summation order; rtol 1e-2 for the opt-in BF16 tensor-core GEMM (tolerance stated in the test).
Full BASELINE sizes are covered through size-independent properties plus sampled oracle blocks.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import b200ka as KA
from b200ka import _lib
from b200ka import examples as EX
from b200ka import kernels as K
from b200ka.kernel import Val, ki_kernel

pytestmark = pytest.mark.gpu
be = KA.B200Backend()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tune(**kv):
    for k, v in kv.items():
        _lib.call("b200_tune_set", k.replace("__", ".").encode(), int(v))


@pytest.fixture(autouse=True)
def _reset_tuning():
    yield
    _lib.call("b200_tune_set", b"transpose.tma_stages", 3)
    _lib.call("b200_tune_set", b"transpose.tma_ctas_per_sm", 4)
    for k in ["registry.force_twins", "transpose.order", "transpose.force_generic", "transpose.impl"]:
        _lib.call("b200_tune_set", k.encode(), 0)
    for k, v in [("hist.impl", 1), ("hist.interleave", -1), ("hist.tma_stages", 4), ("copy.impl", 1), ("hist.threads", 512), ("hist.batch", 8), ("hist.ilp", 2), ("copy.unroll", 8),
                 ("copy.ctas_per_sm", 32), ("copy.bulk_stage_kb", 16), ("copy.bulk_stages", 0)]:
        _lib.call("b200_tune_set", k.encode(), v)


def F(shape, dtype):
    return np.zeros(shape, dtype=dtype, order="F")


# =====================================================================================================
# copy / fill
# =====================================================================================================
@pytest.mark.parametrize("n", [1, 3, 17, 1000, 4097, (1 << 20) + 3, 1 << 22])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8])
def test_memcopy_matches_oracle(orc, n, dtype):
    src = (orc.gen_uniform_f32((n,), 1234) * 255).astype(dtype)
    B = KA.to_device(src)
    A = KA.zeros(be, dtype, n)
    EX.mycopy_(A, B)
    KA.synchronize(be)
    ref = np.zeros(n, dtype=dtype)
    orc.copy_kernel(ref, src, n, KA.threads_to_workgroupsize(1024, (n,)))
    assert np.array_equal(KA.Array(A).view(np.uint8), ref.view(np.uint8))


@pytest.mark.parametrize("force_twins", [0, 1])
def test_memcopy_examples(orc, force_twins):
    # examples/memcopy.jl:19-23 and memcopy_static.jl:19-23
    tune(registry__force_twins=force_twins)
    for fn in (EX.mycopy_, EX.mycopy_static_):
        A = KA.zeros(be, np.float64, 128, 128)
        B = KA.ones(be, np.float64, 128, 128)
        fn(A, B)
        KA.synchronize(be)
        assert np.array_equal(KA.Array(A), KA.Array(B))
        assert np.all(KA.Array(A) == 1.0)


def test_copy_partial_ndrange_and_unaligned(orc):
    src = orc.gen_uniform_f32((5000,), 7)
    B = KA.to_device(src)
    for force in (0, 1):
        tune(registry__force_twins=force)
        # only the first 1234 elements are in the ndrange: the rest of A must stay untouched
        A = KA.allocate(be, np.float32, 5000).fill_(-1.0)
        K.copy_kernel_(be, 64)(A, B, ndrange=1234)
        ref = np.full(5000, -1.0, dtype=np.float32)
        orc.copy_kernel(ref, src, 1234, 64)
        assert np.array_equal(KA.Array(A), ref)
        # misaligned views (offset by 1 and 3 elements: 4- and 12-byte misalignment)
        A = KA.allocate(be, np.float32, 5000).fill_(-1.0)
        K.copy_kernel_(be, 256)(A.view(1, (4000,)), B.view(3, (4000,)), ndrange=4000)
        ref = np.full(5000, -1.0, dtype=np.float32)
        ref[1:4001] = src[3:4003]
        assert np.array_equal(KA.Array(A), ref)
    # N-d ndrange that is not a multiple of the workgroup: literal twin semantics (padded linear ids)
    A = KA.allocate(be, np.float32, 64).fill_(-1.0)
    K.copy_kernel_(be, (4, 2))(A, B, ndrange=(6, 5))
    ref = np.full(64, -1.0, dtype=np.float32)
    orc.copy_kernel(ref, src[:64].copy(), (6, 5), (4, 2))
    assert np.array_equal(KA.Array(A), ref)


@pytest.mark.parametrize("impl,knobs", [(0, dict(copy__unroll=1)), (0, dict(copy__unroll=2)), (0, dict(copy__unroll=8)),
                                        (1, dict(copy__bulk_stages=2)), (1, dict(copy__bulk_stages=4)),
                                        (1, dict(copy__bulk_stages=8, copy__bulk_stage_kb=16))])
def test_copy_implementations_bit_exact(orc, impl, knobs):
    tune(copy__impl=impl, **knobs)
    for n in [(1 << 22), (1 << 22) + 1028, (3 << 20) + 4]:
        src = orc.gen_uniform_f32((n,), 99)
        B = KA.to_device(src)
        A = KA.zeros(be, np.float32, n)
        KA.copyto_(be, A, B)
        KA.synchronize(be)
        assert np.array_equal(KA.Array(A), src)


def test_copy_c1_size_property(orc):
    # BASELINE config: 2^26 Float32.  Property: copy is the identity on bytes (compare checksums and samples).
    n = 1 << 26
    src = orc.gen_uniform_f32((n,), 1234)
    B = KA.to_device(src)
    A = KA.zeros(be, np.float32, n)
    EX.mycopy_(A, B)
    KA.synchronize(be)
    got = KA.Array(A)
    assert np.array_equal(got.view(np.uint32), src.view(np.uint32))


@pytest.mark.parametrize("dtype,shape", [(np.float32, (7, 5)), (np.float64, (1000,)), (np.int64, (64, 64)), (np.uint8, (12345,)),
                                         (np.int32, (1 << 20,)), (np.float32, (3, 1))])
def test_zeros_ones_fill(orc, dtype, shape):
    Z = KA.zeros(be, dtype, shape)
    O = KA.ones(be, dtype, shape)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(Z), np.zeros(shape, dtype))
    assert np.array_equal(KA.Array(O), np.ones(shape, dtype))
    # init_kernel through the generic launch, fast path and literal twin
    for force in (0, 1):
        tune(registry__force_twins=force)
        A = KA.allocate(be, dtype, shape).fill_(5)
        n = A.size
        K.init_kernel(be)(A, dtype(1), dtype, ndrange=n)
        ref = np.full(shape, 5, dtype=dtype, order="F")
        orc.init_kernel(ref, 1, n, KA.threads_to_workgroupsize(1024, (n,)))
        assert np.array_equal(KA.Array(A), ref)
    # unaligned fill (views)
    A = KA.allocate(be, dtype, 100).fill_(9)
    A.view(3, (90,)).fill_(2)
    ref = np.full(100, 9, dtype=dtype)
    ref[3:93] = 2
    assert np.array_equal(KA.Array(A), ref)


# =====================================================================================================
# transpose
# =====================================================================================================
@pytest.mark.parametrize("shape", [(1024, 1024), (64, 128), (128, 64), (2048, 512), (37, 300), (1, 1), (65, 63), (1000, 1000)])
@pytest.mark.parametrize("order", [0, 1])
def test_transpose_matches_oracle(orc, shape, order):
    tune(transpose__order=order)
    m, n = shape  # a is m x n, b is n x m
    b = orc.gen_uniform_f32((n, m), 1235)
    a_ref = F((m, n), np.float32)
    orc.naive_transpose(a_ref, b)
    dA = KA.zeros(be, np.float32, m, n)
    EX.naive_transpose_(dA, KA.to_device(b))
    KA.synchronize(be)
    assert np.array_equal(KA.Array(dA).view(np.uint32), a_ref.view(np.uint32))


@pytest.mark.parametrize("stages,cps", [(3, 4), (2, 1), (4, 3), (6, 2)])
@pytest.mark.parametrize("shape", [(64, 64), (128, 64), (64, 128), (1024, 1024), (2048, 512), (4096, 192), (64, 9600)])
def test_transpose_tma_variant_matches_oracle(orc, shape, stages, cps):
    """transpose.impl = 1: TMA tensor loads with the tile's columns permuted on the way in (3-d map, 128-byte swizzle), register
    transpose, coalesced stores -- same bits as the definition, for every ring depth / residency, incl. leading dimensions
    larger than the matrix (views)."""
    tune(transpose__impl=1, transpose__tma_stages=stages, transpose__tma_ctas_per_sm=cps)
    m, n = shape
    b = orc.gen_uniform_f32((n, m), 1235)
    a_ref = F((m, n), np.float32)
    orc.naive_transpose(a_ref, b)
    dA = KA.zeros(be, np.float32, m, n)
    EX.naive_transpose_(dA, KA.to_device(b))
    KA.synchronize(be)
    assert np.array_equal(KA.Array(dA).view(np.uint32), a_ref.view(np.uint32))
    # leading dimensions larger than the shapes: the C entry point on sub-matrices of bigger arrays
    big_b = orc.gen_uniform_f32((n + 64, m), 99)
    dBig, dOut = KA.to_device(big_b), KA.allocate(be, np.float32, m + 128, n).fill_(-1.0)
    _lib.call("b200_transpose", C.c_void_p(dOut.ptr + 4 * 64), C.c_void_p(dBig.ptr + 4 * 32), m, n, m + 128, n + 64, 4)
    KA.synchronize(be)
    got = KA.Array(dOut)
    assert np.array_equal(got[64:64 + m, :], big_b[32:32 + n, :].T)
    assert np.all(got[:64, :] == -1.0) and np.all(got[64 + m:, :] == -1.0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
@pytest.mark.parametrize("shape", [(132, 68), (68, 132), (1000, 1004), (4, 4), (2, 6), (64, 64), (192, 320), (100, 36), (2050, 130)])
def test_transpose_any_shape_vector_path(shape, dtype):
    # leading dimensions that are multiples of 16 bytes take transpose64_any_kernel (interior + predicated edge tiles)
    m, n = shape
    rng = np.random.default_rng(m * 131 + n)
    b = np.asfortranarray((rng.standard_normal((n, m)) * 1e3).astype(dtype))
    dA = KA.allocate(be, dtype, m, n).fill_(7)
    dB = KA.to_device(b)
    EX.naive_transpose_(dA, dB)
    assert np.array_equal(KA.Array(dA), b.T)
    pad = 4 if np.dtype(dtype).itemsize == 4 else 2             # keeps lda a multiple of 16 bytes whenever m is
    guard = KA.allocate(be, dtype, m + pad, n).fill_(7)         # a with lda = m + pad: the extra rows must stay untouched
    _lib.call("b200_transpose", C.c_void_p(guard.ptr), C.c_void_p(dB.ptr), m, n, m + pad, n, np.dtype(dtype).itemsize)
    KA.synchronize(be)
    g = KA.Array(guard)
    assert np.array_equal(g[:m, :], b.T) and np.all(g[m:, :] == 7)


@pytest.mark.parametrize("dtype", [np.int8, np.float16, np.complex128])
@pytest.mark.parametrize("shape", [(64, 64), (37, 300), (1, 9), (1000, 129)])
@pytest.mark.parametrize("twins", [0, 1])
def test_transpose_other_element_sizes(orc, shape, dtype, twins):
    """naive_transpose_kernel! is a permutation of isbits elements: 1-, 2- and 16-byte element types (Int8, Float16, ComplexF64)
    through the tile kernel and the literal twin, against the oracle's byte-wise restatement, incl. a partial ndrange."""
    m, n = shape
    rng = np.random.default_rng(m * 7 + n)
    raw = rng.integers(0, 256, size=n * m * np.dtype(dtype).itemsize, dtype=np.uint8)
    b = np.asfortranarray(raw.view(dtype).reshape((n, m), order="F"))
    a_ref = np.zeros((m, n), dtype=dtype, order="F")
    orc.naive_transpose(a_ref, b)
    assert bits_equal(a_ref, np.asfortranarray(b.T))
    tune(registry__force_twins=twins)
    try:
        dA = KA.zeros(be, dtype, m, n)
        dB = KA.to_device(b)
        EX.naive_transpose_(dA, dB)
        assert bits_equal(KA.Array(dA), a_ref)
        if m > 8 and n > 8:
            part = (m - 5, n - 3)
            dA = KA.zeros(be, dtype, m, n)
            K.naive_transpose_kernel_(be, 256)(dA, dB, ndrange=part)
            ref = np.zeros((m, n), dtype=dtype, order="F")
            orc.naive_transpose(ref, b, ndrange=part)
            assert bits_equal(KA.Array(dA), ref)
    finally:
        tune(registry__force_twins=0)


def test_transpose_variants(orc):
    b = orc.gen_uniform_f32((256, 192), 5)
    a_ref = F((192, 256), np.float32)
    orc.naive_transpose(a_ref, b)
    for knobs in (dict(registry__force_twins=1), dict(transpose__force_generic=1)):
        tune(**knobs)
        dA = KA.zeros(be, np.float32, 192, 256)
        EX.naive_transpose_(dA, KA.to_device(b))
        assert np.array_equal(KA.Array(dA), a_ref)
        tune(registry__force_twins=0, transpose__force_generic=0)
    # Float64 (the example's f_type when supports_float64) goes through the generic tile kernel
    b64 = np.asfortranarray(np.random.default_rng(1).random((96, 130)))
    dA = KA.zeros(be, np.float64, 130, 96)
    EX.naive_transpose_(dA, KA.to_device(b64))
    assert np.array_equal(KA.Array(dA), b64.T)
    # size mismatch: prints and returns nothing, `a` untouched (examples/naive_transpose.jl:12-15)
    dA = KA.zeros(be, np.float32, 10, 10)
    assert EX.naive_transpose_(dA, KA.to_device(b)) is None
    assert np.all(KA.Array(dA) == 0)
    # partial ndrange: only the top-left block is written
    dA = KA.allocate(be, np.float32, 192, 256).fill_(-1)
    K.naive_transpose_kernel_(be, 256)(dA, KA.to_device(b), ndrange=(100, 70))
    ref = np.full((192, 256), -1, dtype=np.float32, order="F")
    orc.naive_transpose(ref, b, ndrange=(100, 70))
    assert np.array_equal(KA.Array(dA), ref)


def test_transpose_c2_size_properties(orc):
    # BASELINE config: 16384 x 16384 Float32.  Properties: (1) involution, (2) sampled rows against the definition.
    n = 16384
    b = orc.gen_uniform_f32((n, n), 1235)
    dB = KA.to_device(b)
    dA = KA.zeros(be, np.float32, n, n)
    EX.naive_transpose_(dA, dB)
    dB2 = KA.zeros(be, np.float32, n, n)
    EX.naive_transpose_(dB2, dA)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(dB2).view(np.uint32), b.view(np.uint32))
    a = KA.Array(dA)
    for i in [0, 1, 63, 64, 4097, n - 1]:
        assert np.array_equal(a[i, :], b[:, i])
        assert np.array_equal(a[:, i], b[i, :])


# =====================================================================================================
# histogram
# =====================================================================================================
def _hist_check(orc, inp, nbins, gs=256):
    base = np.zeros(nbins, dtype=np.int32)
    orc.histogram(base, inp, gs)
    out = KA.zeros(be, np.int32, nbins)
    EX.histogram_(out, EX.move(be, inp), gs)
    KA.synchronize(be)
    got = KA.Array(out)
    assert np.array_equal(got, base), f"max |diff| = {np.abs(got.astype(np.int64) - base).max()}"
    return got


def test_histogram_examples(orc):
    # examples/histogram.jl:74-100
    rng = np.random.default_rng(1236)
    rand_input = rng.integers(1, 129, size=1000).astype(np.int32)
    linear_input = np.arange(1, 1025, dtype=np.int32)
    all_two = np.full(512, 2, dtype=np.int32)
    for force in (0, 1):
        tune(registry__force_twins=force)
        for inp, gs in ((rand_input, 6), (linear_input, 256), (all_two, 256)):
            got = _hist_check(orc, inp, int(inp.max()), gs)
            assert np.array_equal(got, orc.create_histogram(inp))


@pytest.mark.parametrize("cfg", [dict(hist__impl=0, hist__threads=512, hist__batch=8, hist__ilp=2),
                                 dict(hist__impl=0, hist__threads=512, hist__batch=8, hist__ilp=1),
                                 dict(hist__impl=0, hist__threads=512, hist__batch=4, hist__ilp=2),
                                 dict(hist__impl=0, hist__threads=768, hist__batch=4, hist__ilp=2),
                                 dict(hist__impl=0, hist__threads=256, hist__batch=8, hist__ilp=2),
                                 dict(hist__impl=1), dict(hist__impl=1, hist__interleave=0), dict(hist__impl=1, hist__interleave=1),
                                 dict(hist__impl=2, hist__tma_stages=4), dict(hist__impl=2, hist__tma_stages=2)])
@pytest.mark.parametrize("dist", ["uniform", "all_equal", "ramp", "out_of_range"])
def test_histogram_variants_and_distributions(orc, cfg, dist):
    tune(**cfg)
    n = (1 << 22) + 3  # not a multiple of 4: exercises the scalar tail
    if dist == "uniform":
        inp = orc.gen_bins_i32(n, 1236, 256)
    elif dist == "all_equal":
        inp = np.full(n, 2, dtype=np.int32)  # examples/histogram.jl:78 at scale: worst case for atomics
    elif dist == "ramp":
        inp = (np.arange(n, dtype=np.int64) * 256 // n + 1).astype(np.int32)  # sorted: long equal runs
    else:
        inp = orc.gen_bins_i32(n, 5, 256)
        inp[::7] = 0
        inp[1::11] = -5
        inp[2::13] = 300
        inp[3::17] = np.iinfo(np.int32).min
    _hist_check(orc, inp, 256)


@pytest.mark.parametrize("impl,interleave", [(1, 0), (1, 1), (2, 1)])
@pytest.mark.parametrize("n,offset", [(1, 0), (3, 1), (5, 0), (127, 3), (1024 * 4, 0), (1024 * 4 + 1, 2), (148 * 4096 + 77, 1), (3 * (1 << 20) + 5, 3),
                                      (1 << 22, 0), (100003, 2)])
def test_histogram_unit_plans_sizes_and_alignments(orc, impl, interleave, n, offset):
    """contiguous slices, interleaved 16 KiB units and the TMA-fed ring: every size class (less than one int4, less than one unit,
    fewer units than CTAs, ragged last unit) and every misalignment of the input against 16 bytes / 512-byte lines."""
    tune(hist__impl=impl, hist__interleave=interleave)
    whole = orc.gen_bins_i32(n + offset, 77 + n % 13, 256)
    d_all = EX.move(be, whole)
    want = np.zeros(256, dtype=np.int32)
    orc.histogram(want, np.ascontiguousarray(whole[offset:]), 256)
    out = KA.zeros(be, np.int32, 256)
    EX.histogram_(out, d_all.view(offset, (n,)))
    KA.synchronize(be)
    assert np.array_equal(KA.Array(out), want)


def test_histogram_accumulates_and_other_bin_counts(orc):
    inp = orc.gen_bins_i32(100000, 3, 100)
    # `+=` semantics: a second launch adds to the existing bins (examples/histogram.jl:53)
    out = KA.zeros(be, np.int32, 100)
    dinp = EX.move(be, inp)
    EX.histogram_(out, dinp)
    EX.histogram_(out, dinp)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(out), 2 * np.bincount(inp, minlength=101)[1:])
    for nbins in (1, 37, 255, 257, 1000, 5000, 20000):
        inp = orc.gen_bins_i32(50000, nbins, nbins)
        _hist_check(orc, inp, nbins)
    # unaligned input pointer
    inp = orc.gen_bins_i32(10001, 9, 256)
    d = EX.move(be, inp)
    out = KA.zeros(be, np.int32, 256)
    v = d.view(1, (10000,))
    EX.histogram_(out, v)
    assert np.array_equal(KA.Array(out), np.bincount(inp[1:], minlength=257)[1:])


def test_histogram_c3_full_size(orc):
    # BASELINE config: 2^30 Int32 into 256 bins.  Checked against the serial reference baseline on the whole input.
    n = 1 << 30
    inp = orc.gen_bins_i32(n, 1236, 256)
    want = orc.create_histogram(inp)  # the reference's serial baseline (examples/histogram.jl:9-15)
    d = EX.move(be, inp)
    out = KA.zeros(be, np.int32, 256)
    EX.histogram_(out, d)
    KA.synchronize(be)
    got = KA.Array(out)
    assert int(got.astype(np.int64).sum()) == n  # conservation
    assert np.array_equal(got, want)


# =====================================================================================================
# matmul
# =====================================================================================================
def test_matmul_example_bit_exact(orc):
    # examples/matmul.jl:29-36
    a = orc.gen_uniform_f32((256, 123), 1237)
    b = orc.gen_uniform_f32((123, 45), 1238)
    out = KA.zeros(be, np.float32, 256, 45)
    EX.matmul_(out, KA.to_device(a), KA.to_device(b))
    KA.synchronize(be)
    ref = F((256, 45), np.float32)
    orc.matmul_naive(ref, a, b)
    assert np.array_equal(KA.Array(out).view(np.uint32), ref.view(np.uint32))
    # Float64 variant
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    out64 = KA.zeros(be, np.float64, 256, 45)
    EX.matmul_(out64, KA.to_device(a64), KA.to_device(b64))
    ref64 = F((256, 45), np.float64)
    orc.matmul_naive(ref64, np.asfortranarray(a64), np.asfortranarray(b64))
    assert np.array_equal(KA.Array(out64), ref64)


@pytest.mark.parametrize("shape", [(256, 124, 48), (128, 16, 128), (132, 20, 260), (4, 4, 1), (512, 256, 1000)])
def test_matmul_example_tiled_route_bit_exact(orc, shape):
    # examples/matmul.jl:5-27 with TMA-addressable operands takes the tiled engine in its UNFUSED mode: same bits as the
    # oracle's literal `tmp += a*b` (separate roundings, k ascending) and as the literal GPU twin
    M, Kd, N = shape
    a = orc.gen_uniform_f32((M, Kd), 1237) - 0.5
    b = orc.gen_uniform_f32((Kd, N), 1238) - 0.5
    dA, dB = KA.to_device(a), KA.to_device(b)
    out = KA.allocate(be, np.float32, M, N).fill_(-7.0)
    EX.matmul_(out, dA, dB)
    ref = F((M, N), np.float32)
    orc.matmul_naive(ref, a, b)
    assert np.array_equal(KA.Array(out).view(np.uint32), ref.view(np.uint32))
    tune(registry__force_twins=1)
    out2 = KA.allocate(be, np.float32, M, N).fill_(-7.0)
    EX.matmul_(out2, dA, dB)
    tune(registry__force_twins=0)
    assert KA.isequal(out, out2)


def test_matmul_example_c4_size_route():
    # 2048^3 through the tiled route against the literal twin (GPU vs GPU, every bit), and within rtol 1e-5 of Float64
    n = 2048
    dA = KA.rand_uniform_(KA.allocate(be, np.float32, n, n), 1237)
    dB = KA.rand_uniform_(KA.allocate(be, np.float32, n, n), 1238)
    out, out2 = KA.zeros(be, np.float32, n, n), KA.zeros(be, np.float32, n, n)
    EX.matmul_(out, dA, dB)
    tune(registry__force_twins=1)
    EX.matmul_(out2, dA, dB)
    tune(registry__force_twins=0)
    assert KA.isequal(out, out2)
    hA, hB = KA.Array(dA)[:64, :].astype(np.float64), KA.Array(dB)[:, :64].astype(np.float64)
    assert np.allclose(KA.Array(out)[:64, :64], hA @ hB, rtol=1e-5, atol=0)


@pytest.mark.parametrize("shape", [(1024, 512, 2048), (128, 16, 128), (256, 123, 45), (37, 29, 21), (130, 17, 129), (16, 16, 16)])
@pytest.mark.parametrize("force_twins", [0, 1])
def test_performant_matmul_bit_exact(orc, shape, force_twins):
    # examples/performant_matmul.jl:79-92 (first shape) + ragged shapes; FP32 parity: bit-identical to the oracle,
    # and within rtol 1e-5 of FP64 truth for the example's uniform [0,1) inputs
    tune(registry__force_twins=force_twins)
    N, R, M = shape
    A = orc.gen_uniform_f32((N, R), 1237)
    B = orc.gen_uniform_f32((R, M), 1238)
    dC = KA.zeros(be, np.float32, N, M)
    EX.performant_matmul_(dC, KA.to_device(A), KA.to_device(B))
    KA.synchronize(be)
    ref = F((N, M), np.float32)
    orc.coalesced_matmul(ref, A, B)
    got = KA.Array(dC)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.allclose(got, orc.matmul_f64_truth(A, B), rtol=1e-5, atol=0)


def test_matmul_signed_and_running_mode(orc):
    rng = np.random.default_rng(8)
    A = np.asfortranarray(rng.standard_normal((256, 200)).astype(np.float32))
    B = np.asfortranarray(rng.standard_normal((200, 384)).astype(np.float32))
    dA, dB = KA.to_device(A), KA.to_device(B)
    dC = KA.zeros(be, np.float32, 256, 384)
    EX.performant_matmul_(dC, dA, dB)
    ref = F((256, 384), np.float32)
    orc.coalesced_matmul(ref, A, B)
    assert np.array_equal(KA.Array(dC).view(np.uint32), ref.view(np.uint32))
    # typed entry point, running-accumulator mode, aligned and ragged
    for (M, Kd, N) in [(256, 200, 384), (256, 192, 384), (100, 50, 70)]:
        A2 = np.asfortranarray(A[:M, :Kd]); B2 = np.asfortranarray(B[:Kd, :N])
        dC2 = KA.zeros(be, np.float32, M, N)
        dA2, dB2 = KA.to_device(A2), KA.to_device(B2)
        _lib.call("b200_matmul_f32", C.c_void_p(dC2.ptr), C.c_void_p(dA2.ptr), C.c_void_p(dB2.ptr),
                  M, Kd, N, M, Kd, M, _lib.B200_MATMUL_RUNNING)
        ref2 = F((M, N), np.float32)
        orc.matmul_running(ref2, A2, B2)
        assert np.array_equal(KA.Array(dC2).view(np.uint32), ref2.view(np.uint32))


@pytest.mark.parametrize("pw", [1, 0])
@pytest.mark.parametrize("nj", [6, 7, 8])
@pytest.mark.parametrize("shape", [(256, 200, 1000), (130, 17, 129), (384, 64, 112)])
def test_matmul_tile_width_variants_bit_exact(orc, nj, shape, pw):
    # the wave-filling tile widths (128 x 16 nj, csrc/kernels/matmul_f32_tma.cu pick_nj) and the producer placement
    # (dedicated warp group vs thread 0) must not change a single bit
    M, Kd, N = shape
    A = orc.gen_uniform_f32((M, Kd), 1237) - 0.5
    B = orc.gen_uniform_f32((Kd, N), 1238) - 0.5
    dA, dB = KA.to_device(A), KA.to_device(B)
    tune(sgemm__nj=nj, sgemm__producer_warp=pw)
    try:
        for mode, ref_fn in ((_lib.B200_MATMUL_TILE16, orc.coalesced_matmul), (_lib.B200_MATMUL_RUNNING, orc.matmul_running)):
            dC = KA.allocate(be, np.float32, M, N).fill_(-7.0)
            _lib.call("b200_matmul_f32", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, Kd, N, M, Kd, M, mode)
            ref = F((M, N), np.float32)
            ref_fn(ref, A, B)
            assert np.array_equal(KA.Array(dC).view(np.uint32), ref.view(np.uint32))
    finally:
        tune(sgemm__nj=0, sgemm__producer_warp=1)


def test_matmul_c4_size_sampled_blocks(orc):
    # BASELINE config: 8192^3 FP32.  The oracle cannot run 1.1 TFLOP in seconds, so bit-exactness is checked on
    # three 128x128 output blocks (each needs only a row panel of A and a column panel of B) and rtol 1e-5 vs FP64.
    n = 8192
    A = orc.gen_uniform_f32((n, n), 1237)
    B = orc.gen_uniform_f32((n, n), 1238)
    dC = KA.zeros(be, np.float32, n, n)
    EX.performant_matmul_(dC, KA.to_device(A), KA.to_device(B))
    KA.synchronize(be)
    got = KA.Array(dC)
    for (r0, c0) in [(0, 0), (4096 - 64, 4096 + 64), (n - 128, n - 128)]:
        Ap = np.asfortranarray(A[r0:r0 + 128, :])
        Bp = np.asfortranarray(B[:, c0:c0 + 128])
        ref = F((128, 128), np.float32)
        orc.coalesced_matmul(ref, Ap, Bp)
        blk = got[r0:r0 + 128, c0:c0 + 128]
        assert np.array_equal(blk.view(np.uint32), ref.view(np.uint32))
        assert np.allclose(blk, orc.matmul_f64_truth(Ap, Bp), rtol=1e-5, atol=0)


def test_matmul_bf16_tensor_core():
    # Runs in a child process with a hard timeout: a mis-synchronised tcgen05 pipeline must not hang the suite.
    script = os.path.join(ROOT, "tests", "gpu_tc_check.py")
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=300)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0, "tensor-core GEMM check failed:\n" + r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("plan", [-1, 0, 1])
@pytest.mark.parametrize("n,nbins", [(1000, 128), (1, 1), (2, 7), (100003, 256), (1 << 20, 1000), (50000, 14336), (50000, 20000), ((1 << 22) + 1, 256),
                                     (148 * 2048 * 3 + 7, 256)])
def test_histogram_int64_matches_int32_oracle(orc, n, nbins, plan):
    tune(hist__interleave=plan)      # unit plan of the counting phase shared with the Int32 kernel: by size / slices / interleaved
    # histogram_kernel! is generic in the element type (examples/histogram.jl:18); Julia's `rand(1:128, n)` is Int64.
    # The Int64 path must give the counts of the pinned Int32 oracle on the same values, incl. out-of-range ones.
    inp32 = orc.gen_bins_i32(n, 99, nbins + 3) - 1          # values 0 .. nbins+2: 0 and > nbins are ignored
    want = np.zeros(nbins, dtype=np.int32)
    orc.histogram(want, inp32, 256)
    inp64 = inp32.astype(np.int64)
    if n > 3:
        inp64[1] = -5                                        # negative and huge values are out of range too
        inp64[2] = (1 << 40) + 1
        want[:] = 0
        orc.histogram(want, np.where((inp64 >= 1) & (inp64 <= nbins), inp64, 0).astype(np.int32), 256)
    d = KA.to_device(inp64)
    bins = KA.ones(be, np.int64, nbins)                      # accumulates like the reference's +=
    EX.histogram_(bins, d)
    assert np.array_equal(KA.Array(bins), want.astype(np.int64) + 1)
    if n > 8:                                                # 8-byte aligned but not 16-byte aligned input
        bins2 = KA.zeros(be, np.int64, nbins)
        EX.histogram_(bins2, d.view(1, (n - 1,)))
        w2 = np.zeros(nbins, dtype=np.int32)
        orc.histogram(w2, np.where((inp64[1:] >= 1) & (inp64[1:] <= nbins), inp64[1:], 0).astype(np.int32), 256)
        assert np.array_equal(KA.Array(bins2), w2.astype(np.int64))


@pytest.mark.parametrize("shape,slab_mib", [((256, 64, 128), 32), ((1000, 300, 260), 1), ((130, 17, 129), 1), ((2048, 512, 1024), 1), ((129, 1, 3), 32)])
def test_matmul_host_pipeline_matches_oracle_and_device(orc, shape, slab_mib):
    # performant_matmul! on host Arrays: B uploaded once, A / C row panels streamed; every bit equal to the device-resident product
    M, Kd, N = shape
    tune(hostpipe__slab_mib=slab_mib)
    try:
        A = orc.gen_uniform_f32((M, Kd), 1237) - 0.5
        B = orc.gen_uniform_f32((Kd, N), 1238) - 0.5
        Ch = np.full((M, N), -7.0, dtype=np.float32, order="F")
        KA.pagelock_(be, A)
        EX.performant_matmul_host_(be, Ch, A, B)
        KA.synchronize(be)
        dC = KA.zeros(be, np.float32, M, N)
        EX.performant_matmul_(dC, KA.to_device(A), KA.to_device(B))
        assert np.array_equal(Ch.view(np.uint32), KA.Array(dC).view(np.uint32))
        if M * Kd * N <= 1 << 26:
            ref = F((M, N), np.float32)
            orc.coalesced_matmul(ref, A, B)
            assert np.array_equal(Ch.view(np.uint32), ref.view(np.uint32))
        Ch2 = np.zeros((M, N), dtype=np.float32, order="F")       # a second call right behind the first reuses the workspace
        EX.performant_matmul_host_(be, Ch2, A, B)
        EX.performant_matmul_host_(be, Ch, A, B)
        KA.synchronize(be)
        assert np.array_equal(Ch2, Ch)
    finally:
        KA.unpagelock_(be, A)
        tune(hostpipe__slab_mib=32)


# =====================================================================================================
# host-buffer pipeline (b200_transpose_host / b200_histogram_i32_host): same bits as the oracle, host arrays in and out
# =====================================================================================================
def bits_equal(x, y):
    return x.shape == y.shape and np.array_equal(x.ravel(order="K").view(np.uint8), y.ravel(order="K").view(np.uint8))


@pytest.mark.parametrize("shape,dtype,slab_mib", [((256, 128), np.float32, 32), ((300, 1000), np.float32, 1), ((1000, 333), np.float64, 1),
                                                  ((4096, 2048), np.float32, 4), ((64, 64), np.float32, 32), ((1, 7), np.float32, 32),
                                                  ((515, 260), np.int16, 1), ((130, 70), np.complex128, 1)])
def test_transpose_host_pipeline_matches_oracle(orc, shape, dtype, slab_mib):
    tune(hostpipe__slab_mib=slab_mib)
    b = None
    try:
        m, n = shape
        a = np.full((m, n), -1, dtype=dtype, order="F")
        b = np.asfortranarray((orc.gen_uniform_f32((n, m), 1235) * 1000).astype(dtype))
        KA.pagelock_(be, b)
        EX.naive_transpose_host_(be, a, b)          # pageable destination, pinned source: both must work
        KA.synchronize(be)
        ref = F((m, n), dtype)
        orc.naive_transpose(ref, b)
        assert bits_equal(a, ref)
        # a second call right behind the first one reuses the workspace: stream order must keep it correct
        a2 = np.full((n, m), -1, dtype=dtype, order="F")
        EX.naive_transpose_host_(be, a2, a)
        EX.naive_transpose_host_(be, a, a2)
        KA.synchronize(be)
        assert bits_equal(a2, b) and bits_equal(a, ref)
    finally:
        KA.synchronize(be)
        if b is not None:
            KA.unpagelock_(be, b)
        tune(hostpipe__slab_mib=32)


def test_transpose_host_mismatch_prints_and_returns_none(capsys):
    a, b = F((4, 5), np.float32), F((4, 5), np.float32)
    assert EX.naive_transpose_host_(be, a, b) is None
    assert "Matrix size mismatch!" in capsys.readouterr().out


@pytest.mark.parametrize("n,slab_mib", [(100003, 32), (3 * (1 << 20) + 5, 1), (1, 32)])
def test_histogram_host_pipeline_matches_oracle(orc, n, slab_mib):
    tune(hostpipe__slab_mib=slab_mib)
    try:
        inp = orc.gen_bins_i32(n, 1236, 256)
        bins = np.arange(256, dtype=np.int32)       # accumulates into existing counts like the reference's atomic +=
        EX.histogram_host_(be, bins, inp)
        KA.synchronize(be)
        want = np.arange(256, dtype=np.int32)
        orc.histogram(want, inp, 256)
        assert np.array_equal(bins, want)
    finally:
        tune(hostpipe__slab_mib=32)


# =====================================================================================================
# examples/performance.jl: six copy/transpose variants x the script's launch shapes, tuned route and literal twins
# =====================================================================================================
@pytest.mark.parametrize("force_twins", [0, 1])
@pytest.mark.parametrize("N", [256, 2048])
def test_performance_jl_variants_match_oracle(orc, force_twins, N):
    if N == 2048 and force_twins:
        N = 512   # the literal twins are slow by construction; the script's own size is covered by the tuned route
    tune(registry__force_twins=force_twins)
    inp = np.asfortranarray(orc.gen_uniform_f32((N, N), 77))
    d_in = KA.to_device(inp)
    labels = []
    shapes = {"Simple": None}
    for label, launch in EX.performance_variants(be, N=N):
        out = KA.allocate(be, np.float32, N, N).fill_(-1.0)
        launch(out, d_in)
        KA.synchronize(be)
        which = ("simple_" if label.startswith("Simple") else "lmem_" if "multiple" not in label else "coalesced_") + \
                ("transpose" if "transpose" in label else "copy")
        if label.startswith("Simple"):
            wg = eval(label.split(" ", 2)[2])
            nd, bank = (N, N), 1
        elif "multiple" in label:
            wg, nd, bank = (32, 8), (N, N // 4), int("bank=True" in label)
        else:
            wg, nd, bank = (32, 32), (N, N), int("bank=True" in label)
        ref = np.full((N, N), -1.0, dtype=np.float32, order="F")
        orc.performance_kernel(which, ref, inp, nd, wg, bank)
        assert np.array_equal(KA.Array(out), ref), label
        assert np.array_equal(ref, inp.T if "transpose" in label else inp), label
        labels.append(label)
    assert len(labels) == 14


@pytest.mark.parametrize("force_twins", [0, 1])
def test_performance_jl_partial_and_rectangular(orc, force_twins):
    """ndrange smaller than the arrays / non-square input: only the launched rectangle may be written."""
    tune(registry__force_twins=force_twins)
    inp = np.asfortranarray(orc.gen_uniform_f32((160, 96), 5))
    d_in = KA.to_device(inp)
    # bounds-checked kernels with a ragged ndrange
    for name, ctor, oshape in (("simple_copy", K.simple_copy_kernel_, (160, 96)), ("simple_transpose", K.simple_transpose_kernel_, (96, 160))):
        out = KA.allocate(be, np.float32, *oshape).fill_(-2.0)
        ctor(be, (32, 8))(out, d_in, ndrange=(100, 50))
        KA.synchronize(be)
        ref = np.full(oshape, -2.0, dtype=np.float32, order="F")
        orc.performance_kernel(name, ref, inp, (100, 50), (32, 8), 1)
        assert np.array_equal(KA.Array(out), ref), name
    # unsafe_indices kernels on a sub-rectangle of whole groups
    for name, ctor, wg, nd, oshape in (("lmem_copy", K.lmem_copy_kernel_, (32, 32), (64, 96), (160, 96)),
                                       ("lmem_transpose", K.lmem_transpose_kernel_, (16, 16), (160, 48), (96, 160)),
                                       ("coalesced_copy", K.coalesced_copy_kernel_, (32, 8), (128, 16), (160, 96)),
                                       ("coalesced_transpose", K.coalesced_transpose_kernel_, (32, 4), (160, 12), (96, 160))):
        for bank in (0, 1):
            out = KA.allocate(be, np.float32, *oshape).fill_(-3.0)
            ctor(be, wg)(out, d_in, Val(bank), ndrange=nd)
            KA.synchronize(be)
            ref = np.full(oshape, -3.0, dtype=np.float32, order="F")
            orc.performance_kernel(name, ref, inp, nd, wg, bank)
            assert np.array_equal(KA.Array(out), ref), (name, bank)


def test_performance_jl_out_of_bounds_launch_is_rejected():
    d_in = KA.zeros(be, np.float32, 64, 64)
    out = KA.zeros(be, np.float32, 64, 64)
    with pytest.raises(_lib.B200Error):
        K.lmem_copy_kernel_(be, (32, 32))(out, d_in, ndrange=(96, 64))     # third group row would read past the array
    with pytest.raises(_lib.B200Error):
        K.lmem_transpose_kernel_(be, (32, 8))(out, d_in, Val(1), ndrange=(64, 64))   # tile[j, i] needs a square group


# =====================================================================================================
# SAXPY (benchmark/benchmarks.jl:29-70, examples/numa_aware.jl:10-56): unfused a*x + y, bit-exact
# =====================================================================================================
@pytest.mark.parametrize("force_twins", [0, 1])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_saxpy_benchmark_sizes_match_oracle(orc, dtype, force_twins):
    tune(registry__force_twins=force_twins)
    rng = np.random.default_rng(11)
    for N in (64, 256, 512, 1024, 2048, 4096, 16384, 32768, 65536, 262144, 1048576, 1000003):
        X, Y = rng.random(N).astype(dtype), rng.random(N).astype(dtype)
        dX, dY = KA.to_device(X), KA.to_device(Y)
        for static in (True, False):
            Z = KA.zeros(be, dtype, N)
            kernel = K.saxpy_kernel_(be, 1024) if static else K.saxpy_kernel_(be)
            kernel(Z, dtype(2.0), dX, dY, ndrange=Z.shape)
            KA.synchronize(be)
            ref = np.zeros(N, dtype=dtype)
            orc.saxpy(ref, 2.0, X, Y, (N,), (1024,))
            assert np.array_equal(KA.Array(Z), ref), (N, static)


def test_saxpy_inplace_numa_aware_form(orc):
    rng = np.random.default_rng(12)
    N = 3 * 1024 * 1024 + 7
    X, Y = rng.random(N, dtype=np.float32), rng.random(N, dtype=np.float32)
    dX, dY = KA.to_device(X), KA.to_device(Y)
    a = np.float32(3.1415)
    K.saxpy_kernel(be, 1024, dY.shape)(a, dX, dY, ndrange=dY.shape)
    KA.synchronize(be)
    ref = np.zeros(N, dtype=np.float32)
    orc.saxpy(ref, float(a), X, Y, (N,), (1024,))
    assert np.array_equal(KA.Array(dY), ref)
    # partial ndrange: the tail of Y stays untouched
    dY2 = KA.to_device(Y)
    K.saxpy_kernel(be, 256)(a, dX, dY2, ndrange=1000)
    KA.synchronize(be)
    got = KA.Array(dY2)
    assert np.array_equal(got[:1000], ref[:1000]) and np.array_equal(got[1000:], Y[1000:])
    bw, fl = EX.measure_membw(be, N=1 << 22, samples=2)
    assert bw > 0 and fl > 0
