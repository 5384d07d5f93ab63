"""Device-array conveniences: maximum / minimum / sum, `==`, isapprox, mul!, and the device-side input
generators -- the operations the reference's examples apply to their arrays around the kernels (examples/histogram.jl:10,
88-99; memcopy.jl:23; naive_transpose.jl:32; matmul.jl:36; performant_matmul.jl:92).  Integer results are exact; float
sums are compared with a Float64 host sum."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu

import b200ka as KA  # noqa: E402
from b200ka import examples as EX  # noqa: E402

be = KA.B200Backend()


@pytest.fixture(scope="module")
def orc():
    import ka_oracle
    return ka_oracle


SIZES = [1, 3, 4, 5, 255, 1024, 1025, 100003, (1 << 20) + 7, (1 << 24) + 1]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64])
def test_reductions_match_host(n, dtype):
    rng = np.random.default_rng(n)
    if np.dtype(dtype).kind == "f":
        h = (rng.standard_normal(n) * 1000).astype(dtype)
    else:
        info = np.iinfo(dtype)
        h = rng.integers(max(info.min, -(1 << 40)), min(info.max, 1 << 40), n, dtype=dtype)
    d = KA.to_device(h)
    assert KA.maximum(d) == h.max() and KA.maximum(d).dtype == np.dtype(dtype)
    assert KA.minimum(d) == h.min()
    if np.dtype(dtype).kind == "f":
        assert np.isclose(KA.sum(d), h.astype(np.float64).sum(), rtol=1e-12, atol=1e-6)
    else:
        wide = np.int64 if np.dtype(dtype).kind == "i" else np.uint64
        assert KA.sum(d) == h.astype(wide).sum()


def test_reduction_julia_float_semantics():
    h = np.array([1.0, -0.0, 0.0, -3.0], dtype=np.float32)
    assert not np.signbit(KA.maximum(KA.to_device(np.array([-0.0, 0.0, -0.0], dtype=np.float32))))   # max(-0.0, 0.0) == 0.0
    assert np.signbit(KA.minimum(KA.to_device(np.array([0.0, -0.0, 0.0], dtype=np.float64))))
    h[2] = np.nan
    assert np.isnan(KA.maximum(KA.to_device(h))) and np.isnan(KA.minimum(KA.to_device(h)))           # NaN propagates
    with pytest.raises(KA.ArgumentError):
        KA.maximum(KA.zeros(be, np.int32, 0))
    assert KA.sum(KA.zeros(be, np.int32, 0)) == 0


@pytest.mark.parametrize("n", [1, 17, 4096, 100003])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64, np.uint8, np.bool_])
def test_isequal(n, dtype):
    rng = np.random.default_rng(n + 1)
    h = (rng.integers(0, 2, n) if dtype == np.bool_ else rng.integers(0, 100, n)).astype(dtype)
    a, b = KA.to_device(h), KA.to_device(h.copy())
    assert KA.isequal(a, b)
    h2 = h.copy()
    h2[n // 2] = 1 - h2[n // 2] if dtype == np.bool_ else h2[n // 2] + 1
    assert not KA.isequal(a, KA.to_device(h2))
    h2 = h.copy()
    h2[-1] = 1 - h2[-1] if dtype == np.bool_ else h2[-1] + 1      # last element: the scalar tail
    assert not KA.isequal(a, KA.to_device(h2))
    assert not KA.isequal(a, KA.to_device(np.concatenate([h, h[:1]])))


def test_isequal_float_semantics():
    z = KA.to_device(np.array([0.0, 1.0], dtype=np.float32))
    nz = KA.to_device(np.array([-0.0, 1.0], dtype=np.float32))
    assert KA.isequal(z, nz)                                            # -0.0 == 0.0
    nan = KA.to_device(np.array([np.nan, 1.0], dtype=np.float32))
    assert not KA.isequal(nan, nan)                                     # NaN != NaN


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_isapprox(dtype):
    rng = np.random.default_rng(3)
    h = rng.standard_normal((300, 41)).astype(dtype)
    a = KA.to_device(h)
    eps = np.finfo(dtype).eps
    assert KA.isapprox(a, KA.to_device(h * (1 + 4 * eps)))
    assert not KA.isapprox(a, KA.to_device(h * (1 + 10 * np.sqrt(eps))))
    assert KA.isapprox(a, KA.to_device(h + 1e-3), atol=1.0) and not KA.isapprox(a, KA.to_device(h + 1e-3), atol=1e-6)


def test_examples_run_on_device_arrays_without_host_round_trips(orc):
    # examples/histogram.jl:74-100: the histogram is sized by maximum(input) of the DEVICE array
    for inp in (orc.gen_bins_i32(1000, 1236, 128), np.arange(1, 1025, dtype=np.int32), np.full(512, 2, dtype=np.int32)):
        d_in = EX.move(be, inp)
        hist = KA.zeros(be, inp.dtype, int(KA.maximum(d_in)))
        EX.histogram_(hist, d_in)
        want = np.zeros(int(inp.max()), dtype=np.int32)
        orc.histogram(want, inp, want.size)
        assert KA.isequal(hist, KA.to_device(want))
        assert KA.sum(hist) == inp.size
    # examples/memcopy.jl:19-23:  A == B
    A, B = KA.zeros(be, np.float64, 128, 128), KA.ones(be, np.float64, 128, 128)
    assert not KA.isequal(A, B)
    EX.mycopy_(A, B)
    assert KA.isequal(A, B)
    # examples/naive_transpose.jl:24-32:  a == transpose(b), checked as transpose(a) == b on the device
    b = KA.rand_uniform_(KA.allocate(be, np.float32, 96, 160), 1235)
    a, back = KA.zeros(be, np.float32, 160, 96), KA.zeros(be, np.float32, 96, 160)
    EX.naive_transpose_(a, b)
    EX.naive_transpose_(back, a)
    assert KA.isequal(back, b) and not KA.isequal(a, KA.zeros(be, np.float32, 160, 96))
    # examples/matmul.jl:29-36 and performant_matmul.jl:79-92:  isapprox(output, a * b)
    a = KA.rand_uniform_(KA.allocate(be, np.float32, 256, 123), 1237)
    bm = KA.rand_uniform_(KA.allocate(be, np.float32, 123, 45), 1238)
    out = KA.zeros(be, np.float32, 256, 45)
    EX.matmul_(out, a, bm)
    assert KA.isapprox(out, KA.matmul(a, bm))
    out2 = KA.zeros(be, np.float32, 256, 45)
    EX.performant_matmul_(out2, a, bm)
    assert KA.isapprox(out2, KA.matmul(a, bm)) and KA.isequal(out2, KA.matmul(a, bm))


@pytest.mark.parametrize("n", [1, 5, 4096, 100003, (1 << 22) + 3])
def test_device_generators_bit_identical_to_oracle(orc, n):
    u = KA.rand_uniform_(KA.allocate(be, np.float32, n), 1234)
    assert np.array_equal(KA.Array(u).view(np.uint32), orc.gen_uniform_f32((n,), 1234).view(np.uint32))
    v = KA.rand_bins_(KA.allocate(be, np.int32, n), 1236, 256)
    assert np.array_equal(KA.Array(v), orc.gen_bins_i32(n, 1236, 256))


@pytest.mark.parametrize("off", [1, 2, 3, 5])
def test_unaligned_views(orc, off):
    n = 100003
    h = orc.gen_uniform_f32((n,), 3) - 0.5
    d = KA.to_device(h)
    v = d.view(off, (n - off,))                       # pointer is 4-byte but not 16-byte aligned
    assert KA.maximum(v) == h[off:].max() and KA.minimum(v) == h[off:].min()
    assert np.isclose(KA.sum(v), h[off:].astype(np.float64).sum(), rtol=1e-12)
    e = KA.to_device(h)
    assert KA.isequal(v, e.view(off, (n - off,)))     # equally misaligned: vector path with a scalar head
    assert KA.isequal(v, KA.to_device(h[off:].copy()))   # differently aligned operands: scalar path
    h2 = h.copy(); h2[off] += 1
    assert not KA.isequal(v, KA.to_device(h2).view(off, (n - off,)))
    assert KA.isapprox(v, KA.to_device(h[off:].copy()))
    g = KA.allocate(be, np.float32, n)
    KA.rand_uniform_(g.view(off, (n - off,)), 1234)
    assert np.array_equal(KA.Array(g)[off:].view(np.uint32), orc.gen_uniform_f32((n - off,), 1234).view(np.uint32))


def test_full_size_histogram_input_checks():
    # BASELINE config 3 input generated on the device: 2^30 values in 1..256; size-independent properties
    n = 1 << 30
    d = KA.rand_bins_(KA.allocate(be, np.int32, n), 1236, 256)
    assert KA.maximum(d) == 256 and KA.minimum(d) == 1
    bins = KA.zeros(be, np.int32, 256)
    EX.histogram_(bins, d)
    assert KA.sum(bins) == n                                            # every element counted once
    hb = KA.Array(bins).astype(np.int64)
    assert KA.sum(d) == int((hb * np.arange(1, 257)).sum())             # checksum of checksums: sum(values) from the counts
