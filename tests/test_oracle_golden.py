"""Pins the CPU oracle against every known-answer value the reference's own tests hold for the hot path
(SURVEY.md section 4 / 8c).  Each test cites the reference test it replays."""
import itertools

import numpy as np
import pytest


def F(shape, dtype):
    return np.zeros(shape, dtype=dtype, order="F")


# ---- test/test.jl:14-44 "partition" ---------------------------------------------------------
def test_partition_golden(orc):
    blocks, wgs, dyn = orc.partition((128,), (64,))
    assert int(np.prod(blocks)) == 2 and dyn is False
    blocks, wgs, dyn = orc.partition((129,), (64,))
    assert int(np.prod(blocks)) == 3 and dyn is True


def test_partition_padding_and_zero(orc):
    # src/nditeration.jl:146-153: pad with ones; a 0 entry becomes 1
    blocks, wgs, dyn = orc.partition((16384, 16384), (256,))
    assert wgs == (256, 1) and blocks == (64, 16384) and dyn is False
    blocks, wgs, dyn = orc.partition((10, 7), (0, 2))
    assert wgs == (1, 2) and blocks == (10, 4) and dyn is True
    # zero-size ndrange -> zero groups (test/test.jl:232-236 launches it and nothing runs)
    blocks, wgs, dyn = orc.partition((0,), (1,))
    assert blocks == (0,)


def test_threads_to_workgroupsize(orc):
    # src/pocl/backend.jl:108-115 greedy fill
    assert orc.threads_to_workgroupsize(1024, (256, 45)) == (256, 4)
    assert orc.threads_to_workgroupsize(1024, (4096,)) == (1024,)
    assert orc.threads_to_workgroupsize(1024, (8, 8, 100)) == (8, 8, 16)


# ---- test/nditeration.jl:6-17 "iteration" ------------------------------------------------------
def test_ndrange_iteration_order(orc):
    ctx = orc.raw_ctx((256, 256), (32, 32))
    assert ctx.groups() == 256 * 256
    # iterating blocks is column-major CartesianIndices((256,256))
    expect = [(i, j) for j in range(1, 4) for i in range(1, 257)]
    got = []
    import ctypes as C
    out = (C.c_int64 * 4)()
    for g in range(1, 3 * 256 + 1):
        orc.lib().ka_group_cartesian(C.byref(ctx), C.c_int64(g), out)
        got.append((out[0], out[1]))
    assert got == expect


# ---- test/nditeration.jl:20-95 "linear_iteration" ------------------------------------------------
@pytest.mark.parametrize("blocks,items", [((4, 4), (32, 32)), ((4, 4 * 32), (32, 1)), ((4 * 32, 4), (1, 32))])
def test_linear_iteration(orc, blocks, items):
    ctx = orc.raw_ctx(blocks, items)
    idx = orc.linear_iteration(ctx)
    N = items[0] * items[1]
    gi = 0
    for by in range(blocks[1]):          # CartesianIndices(blocks): dim 1 fastest
        for bx in range(blocks[0]):
            chunk = idx[gi * N:(gi + 1) * N]
            expect = np.array([(bx * items[0] + x + 1, by * items[1] + y + 1)
                               for y in range(items[1]) for x in range(items[0])], dtype=np.int64)
            assert np.array_equal(chunk, expect)
            gi += 1


# ---- test/test.jl:116-160 "indextest" -------------------------------------------------------------
def test_indextest(orc):
    A = F((16, 16), np.int64)
    orc.index_linear(0, A, A.size, 8)
    assert np.array_equal(A.ravel(order="F"), np.arange(1, 257))
    for shape in [(8,), (16,), (8, 2)]:
        A = F(shape, np.int64)
        orc.index_linear(1, A, A.size, 8)
        flat = A.ravel(order="F")
        for s in range(0, flat.size, 8):
            assert np.array_equal(flat[s:s + 8], np.arange(1, 9))
    # Global Cartesian with a 2-D ndrange
    A = np.zeros((2, 16, 16), dtype=np.int64, order="F")
    orc.index_cartesian(0, A, (16, 16), (16, 16), 8)
    for i, j in itertools.product(range(16), range(16)):
        assert tuple(A[:, i, j]) == (i + 1, j + 1)
    # CartesianIndex{1} with a linear ndrange over a 16x16 array
    A = np.zeros((1, 16, 16), dtype=np.int64, order="F")
    orc.index_cartesian(0, A, (16, 16), 256, 8)
    assert np.array_equal(A.ravel(order="F"), np.arange(1, 257))
    # non-multiples of the workgroupsize
    A = F((7, 7), np.int64)
    orc.index_linear(0, A, A.size, 8)
    assert np.array_equal(A.ravel(order="F"), np.arange(1, 50))
    A = F((5,), np.int64)
    orc.index_linear(1, A, A.size, 8)
    assert np.array_equal(A, np.arange(1, 6))


def test_kernel_val(orc):
    # test/test.jl:216-224
    a = np.zeros(1024, dtype=np.int64)
    orc.kernel_val(a, 3, a.shape, orc.threads_to_workgroupsize(1024, a.shape))
    assert np.all(a == 3)


# ---- test/localmem.jl:67-79 ---------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_localmem_golden(orc, variant):
    A = np.full(64, -1, dtype=np.int64)
    orc.localmem(variant, A, A.shape, 16)
    for s in range(0, 64, 16):
        assert np.array_equal(A[s:s + 16], np.arange(16, 0, -1))


# ---- test/private.jl:61-91 ------------------------------------------------------------------------------
def test_private_golden(orc):
    A = np.full(64, -1, dtype=np.int64)
    orc.private(A, A.shape, 16)
    for s in range(0, 64, 16):
        assert np.array_equal(A[s:s + 16], np.arange(16, 0, -1))

    A = np.ones((64, 64), dtype=np.int64, order="F")
    orc.forloop(A, A.shape[0], A.shape[0])
    assert np.all(A[:, 0] == 64)
    assert np.all(A[:, 1:] == 1)

    B = np.zeros((64, 64), dtype=np.uint8, order="F")
    orc.typetest(B, B.shape, 16)
    assert np.all(B == 1)

    A = np.ones((64, 3), dtype=np.float32, order="F")
    out = np.zeros(64, dtype=np.float32)
    orc.reduce_private(out, A, out.shape, 8)
    assert np.all(out == 3.0)


# ---- test/unroll.jl:39-48 (no value asserts upstream; values follow from the kernel bodies) ---------
def test_unroll(orc):
    a = np.zeros(5, dtype=np.float32)
    orc.unroll(0, a)
    assert np.array_equal(a, [1, 2, 3, 4, 5])
    orc.unroll(1, a, 5)
    assert np.array_equal(a, [6, 7, 8, 9, 10])
    orc.unroll(2, a)
    assert a[0] == 9.0 and np.array_equal(a[1:], [7, 8, 9, 10])


# ---- examples/memcopy.jl:19-23, memcopy_static.jl:19-23 --------------------------------------------------
def test_memcopy_examples(orc):
    A = F((128, 128), np.float64)
    B = np.ones((128, 128), dtype=np.float64, order="F")
    orc.copy_kernel(A, B, A.size, orc.threads_to_workgroupsize(1024, (A.size,)))
    assert np.array_equal(A, B)
    A = F((128, 128), np.float64)
    orc.copy_kernel(A, B, A.shape, (32,))  # static variant: wgs (32,) padded to (32,1)
    assert np.array_equal(A, B)


def test_zeros_ones(orc):
    # src/pocl/backend.jl:27-38
    arr = np.full((7, 5), 3.5, dtype=np.float32, order="F")
    orc.init_kernel(arr, 0, arr.size, (35,))
    assert np.all(arr == 0)
    orc.init_kernel(arr, 1, arr.size, (35,))
    assert np.all(arr == 1)


# ---- examples/naive_transpose.jl:24-32 ---------------------------------------------------------------------
def test_naive_transpose_example(orc):
    rng = np.random.default_rng(1235)
    b = np.asfortranarray(rng.random((1024, 1024), dtype=np.float32))
    a = F((1024, 1024), np.float32)
    orc.naive_transpose(a, b)
    assert np.array_equal(a, b.T)


def test_copy_and_saxpy_wg_equal_generic(orc):
    # the work-group-vectorised forms timed as CPU baselines give the bits of the literal per-item forms
    for n in (1, 1023, 1024, 4099):
        B = orc.gen_uniform_f32((n,), 5)
        A1, A2 = np.zeros(n, np.float32), np.zeros(n, np.float32)
        orc.copy_kernel(A1, B, (n,), (1024,))
        orc.copy_wg(A2, B, 1024)
        assert np.array_equal(A1.view(np.uint32), A2.view(np.uint32)) and np.array_equal(A2, B)
        X, Y = orc.gen_uniform_f32((n,), 6), orc.gen_uniform_f32((n,), 7)
        Z1, Z2 = np.zeros(n, np.float32), np.zeros(n, np.float32)
        orc.saxpy(Z1, np.float32(3.1415), X, Y, (n,), 1024)
        orc.saxpy_wg(Z2, np.float32(3.1415), X, Y, 1024)
        assert np.array_equal(Z1.view(np.uint32), Z2.view(np.uint32))


def test_naive_transpose_wg_equals_generic(orc):
    for shape in [(300, 37), (512, 256), (1, 5)]:
        b = orc.gen_uniform_f32((shape[1], shape[0]), 3)
        a1, a2 = F(shape, np.float32), F(shape, np.float32)
        orc.naive_transpose(a1, b)
        orc.naive_transpose_wg(a2, b)
        assert np.array_equal(a1, a2) and np.array_equal(a1, b.T)


def test_naive_transpose_rect_and_ragged(orc):
    rng = np.random.default_rng(7)
    b = np.asfortranarray(rng.random((37, 300), dtype=np.float32))
    a = F((300, 37), np.float32)
    orc.naive_transpose(a, b)          # 300 % 256 != 0 -> DynamicCheck path
    assert np.array_equal(a, b.T)


# ---- examples/histogram.jl:74-100 -----------------------------------------------------------------------------
@pytest.mark.parametrize("which,wg,div", [("simple_copy", (32, 32), 1), ("simple_transpose", (32, 32), 1), ("simple_copy", (1024, 1), 1),
                                          ("simple_transpose", (1, 1024), 1), ("lmem_copy", (32, 32), 1), ("lmem_transpose", (32, 32), 1),
                                          ("coalesced_copy", (32, 8), 4), ("coalesced_transpose", (32, 8), 4)])
@pytest.mark.parametrize("bank", [0, 1])
def test_performance_jl_kernels(orc, which, wg, div, bank):
    """examples/performance.jl:17-205: every variant, at the script's launch shapes, is a copy resp. a transpose."""
    N = 256
    inp = np.asfortranarray(orc.gen_uniform_f32((N, N), 3))
    out = np.full((N, N), -1.0, dtype=np.float32, order="F")
    orc.performance_kernel(which, out, inp, (N, N // div), wg, bank)
    assert np.array_equal(out, inp.T if "transpose" in which else inp)


def test_performance_jl_ragged_simple(orc):
    """The bounds-checked variants honour a ragged ndrange (lanes outside it are masked by __validindex)."""
    inp = np.asfortranarray(orc.gen_uniform_f32((70, 50), 4))
    out = np.full((50, 70), -1.0, dtype=np.float32, order="F")
    orc.performance_kernel("simple_transpose", out, inp, (33, 21), (32, 8), 1)
    want = np.full((50, 70), -1.0, dtype=np.float32, order="F")
    want[:21, :33] = inp[:33, :21].T
    assert np.array_equal(out, want)


def test_context_return_and_unaliased_accumulate(orc):
    """test/test.jl:284-289 (all ones), :314-321 (first half ones), :358-385 (triangular row sums)."""
    a = np.zeros(1024, dtype=np.int64)
    orc.misc_i64(0, a, 1024, 256)
    assert np.all(a == 1)
    a = np.zeros(1024, dtype=np.int64)
    orc.misc_i64(1, a, 1024, 256)
    assert np.all(a[:512] == 1) and np.all(a[512:] == 0)
    N, M = 8, 5
    inp = np.asfortranarray(np.array([[i + k for k in range(1, N + 1)] for i in range(1, M + 1)], dtype=np.float32))
    reference = np.zeros((M, N), dtype=np.float32)
    for i in range(M):
        for j in range(N):
            for k in range(j, N):
                reference[i, j] += inp[i, k]
    for local in (False, True):
        out = np.zeros((M, N), dtype=np.float32, order="F")
        orc.unaliased_accumulate(local, out, inp, N, (M, N), (M, N))
        assert np.array_equal(out, reference)


def test_histogram_examples(orc):
    rng = np.random.default_rng(1236)
    rand_input = rng.integers(1, 129, size=1000).astype(np.int32)
    linear_input = np.arange(1, 1025, dtype=np.int32)
    all_two = np.full(512, 2, dtype=np.int32)
    for inp, gs in ((rand_input, 6), (linear_input, 256), (all_two, 256)):
        base = orc.create_histogram(inp)
        assert np.array_equal(base, np.bincount(inp, minlength=int(inp.max()) + 1)[1:])
        out = np.zeros(int(inp.max()), dtype=np.int32)
        orc.histogram(out, inp, gs)
        assert np.array_equal(out, base)


# ---- examples/matmul.jl:29-36, performant_matmul.jl:79-92 ---------------------------------------------------------
def _isapprox(x, y):  # Julia isapprox(::Array, ::Array): norm(x-y) <= rtol*max(norm(x),norm(y)), rtol = sqrt(eps(Float32))
    x = x.astype(np.float64); y = y.astype(np.float64)
    return np.linalg.norm(x - y) <= np.sqrt(np.finfo(np.float32).eps) * max(np.linalg.norm(x), np.linalg.norm(y))


def test_matmul_example(orc):
    rng = np.random.default_rng(1237)
    a = np.asfortranarray(rng.random((256, 123), dtype=np.float32))
    b = np.asfortranarray(rng.random((123, 45), dtype=np.float32))
    out = F((256, 45), np.float32)
    orc.matmul_naive(out, a, b)
    assert _isapprox(out, a.astype(np.float64) @ b.astype(np.float64))


def test_performant_matmul_example(orc):
    rng = np.random.default_rng(1238)
    N, R, M = 1024, 512, 2048
    A = np.asfortranarray(rng.random((N, R), dtype=np.float32))
    B = np.asfortranarray(rng.random((R, M), dtype=np.float32))
    Cm = F((N, M), np.float32)
    orc.coalesced_matmul(Cm, A, B)
    truth = orc.matmul_f64_truth(A, B)
    assert _isapprox(Cm, truth)
    assert np.allclose(Cm, truth, rtol=1e-5, atol=0)


def test_coalesced_matmul_edges(orc):
    rng = np.random.default_rng(5)
    A = np.asfortranarray(rng.standard_normal((37, 29)).astype(np.float32))
    B = np.asfortranarray(rng.standard_normal((29, 21)).astype(np.float32))
    Cm = F((37, 21), np.float32)
    orc.coalesced_matmul(Cm, A, B)
    assert _isapprox(Cm, A.astype(np.float64) @ B.astype(np.float64))


# ---- test/intrinsics.jl:27-124 --------------------------------------------------------------------------------------
def test_ki_launch_golden(orc):
    arr = np.zeros(4, dtype=np.float32)
    orc.launch_kernel_nd(arr, (2,), (2,))
    assert np.all(arr == 1)
    arr = F((4, 4), np.float32)
    orc.launch_kernel_nd(arr, (2, 2), (2, 2))
    assert np.all(arr == 1)
    arr = F((4, 4, 4), np.float32)
    orc.launch_kernel_nd(arr, (2, 2, 2), (2, 2, 2))
    assert np.all(arr == 1)

    res = orc.test_intrinsics(4, 4, 16)
    for i in range(1, 17):
        gsz, gid, lsz, lid, ng, grp = res[i - 1]
        assert gid == i and gsz == 16 and lsz == 4 and ng == 4
        assert grp == (i - 1) // 4 + 1 and lid == (i - 1) % 4 + 1


# ---- generators ----------------------------------------------------------------------------------------------------
def test_generators(orc):
    x = orc.gen_uniform_f32((1000,), 1234)
    assert x.min() >= 0 and x.max() < 1 and len(np.unique(x)) > 990
    assert np.array_equal(x, orc.gen_uniform_f32((1000,), 1234))
    v = orc.gen_bins_i32(100000, 1236, 256)
    assert v.min() == 1 and v.max() == 256
    counts = np.bincount(v, minlength=257)[1:]
    assert counts.min() > 250 and counts.max() < 550


def test_bf16_rounding(orc):
    import torch
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(orc.round_bf16(x), want)
    assert np.array_equal(orc.to_bf16_bits(x), torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))
