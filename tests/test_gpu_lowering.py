"""GPU: the three lowerings of the NDRange -> grid mapping (csrc/ka_device.cuh: native <=3-d grid, multiply-high unravel, the
reference's 64-bit div/mod) and the scalar / 128-bit vectorised forms of the element-wise twins produce the SAME indices and
the same bytes as the oracle's restatement of `expand` / `__validindex` (reference src/nditeration.jl:74-82,115-134,
src/pocl/backend.jl:207-214), for ragged, rectangular, 3-d and 4-d iteration spaces."""
import numpy as np
import pytest

import b200ka as KA
from b200ka import _lib, kernels as K

pytestmark = pytest.mark.gpu
be = KA.B200Backend()
MODES = {"auto": -1, "grid": 0, "magic": 1, "div64": 2}


def tune(**kv):
    for k, v in kv.items():
        _lib.call("b200_tune_set", k.replace("__", ".").encode(), int(v))


@pytest.fixture(autouse=True)
def _reset():
    yield
    tune(registry__ka_mode=-1, registry__twin_vec=1, registry__force_twins=0)


SPACES = [((40,), (8,)), ((37,), (8,)), ((12, 10), (4, 2)), ((13, 9), (4, 2)), ((7, 5, 3), (2, 2, 2)), ((6, 4, 4), (3, 2, 2)),
          ((5, 3, 2, 3), (2, 2, 1, 2)), ((1000,), (1000,)), ((33, 3), (32, 1)), ((3, 70000), (1, 1))]


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("ndrange,wgs", SPACES)
def test_cartesian_indices_match_oracle_in_every_lowering(orc, mode, ndrange, wgs):
    tune(registry__ka_mode=MODES[mode])
    n = len(ndrange)
    if n > 3:
        pytest.skip("the CartesianIndex twins are built for <= 3 dimensions; 4-d spaces are covered by the element-wise test")
    for which, kern in [(0, K.index_cartesian_global), (1, K.index_cartesian_local), (2, K.index_cartesian_group)]:
        A = KA.to_device(np.zeros(ndrange, dtype=KA.CartesianIndex(n), order="F"))
        kern(be, wgs)(A, ndrange=ndrange)
        ref = np.zeros((n,) + tuple(ndrange), dtype=np.int64, order="F")
        orc.index_cartesian(which, ref, ndrange, ndrange, wgs)
        got = np.moveaxis(KA.Array(A)["I"], -1, 0)
        assert np.array_equal(got, ref), (mode, which)


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("ndrange,wgs", [s for s in SPACES if len(s[0]) <= 2])
def test_linear_indices_match_oracle_in_every_lowering(orc, mode, ndrange, wgs):
    tune(registry__ka_mode=MODES[mode])
    ctx = orc.make_ctx(ndrange, wgs)
    total = ctx.groups() * ctx.items()
    for which, kern in [(0, K.index_linear_global), (1, K.index_linear_local), (2, K.index_linear_group)]:
        A = KA.allocate(be, np.int64, total).fill_(-5)
        kern(be, wgs)(A, ndrange=ndrange)
        ref = np.full(total, -5, dtype=np.int64)
        orc.index_linear(which, ref, ndrange, wgs)
        assert np.array_equal(KA.Array(A), ref), (mode, which)


@pytest.mark.parametrize("vec", [0, 1])
@pytest.mark.parametrize("mode", ["auto", "magic", "div64"])
@pytest.mark.parametrize("ndrange,wgs,dtype", [((1 << 20,), (1024,), np.float32), ((100003,), (256,), np.float32), ((4099,), (64,), np.float64),
                                               ((64, 64), (32, 8), np.float32), ((6, 5), (4, 2), np.float32), ((1000,), (1000,), np.uint8),
                                               ((12345,), (128,), np.uint16), ((7,), (8,), np.float64),
                                               ((5, 3, 2, 3), (2, 2, 1, 2), np.float32), ((8, 4, 2, 2), (4, 2, 2, 1), np.float32),
                                               ((7, 5, 3), (2, 2, 2), np.float64), ((3, 70000), (1, 1), np.float32)])
def test_elementwise_twins_scalar_and_vector_match_oracle(orc, vec, mode, ndrange, wgs, dtype):
    """copy_kernel!, init_kernel (zeros/ones) and saxpy_kernel! through the literal twins: one work-item per thread and the
    128-bit vectorised form, in every lowering, against the oracle -- including the words in front of / behind the ndrange."""
    tune(registry__force_twins=1, registry__twin_vec=vec, registry__ka_mode=MODES[mode])
    ctx = orc.make_ctx(ndrange, wgs)
    total = ctx.groups() * ctx.items() + 9
    src = (orc.gen_uniform_f32((total,), 77) * 200).astype(dtype)
    A = KA.allocate(be, dtype, total).fill_(3)
    K.copy_kernel_(be, wgs)(A, KA.to_device(src), ndrange=ndrange)
    ref = np.full(total, 3, dtype=dtype)
    orc.copy_kernel(ref, src, ndrange, wgs)
    assert np.array_equal(KA.Array(A).view(np.uint8), ref.view(np.uint8))
    if np.dtype(dtype).kind == "f":
        X, Y = (orc.gen_uniform_f32((total,), 5)).astype(dtype), (orc.gen_uniform_f32((total,), 6)).astype(dtype)
        Z = KA.allocate(be, dtype, total).fill_(-1)
        K.saxpy_kernel_(be, wgs)(Z, dtype(2.5), KA.to_device(X), KA.to_device(Y), ndrange=ndrange)
        zref = np.full(total, -1, dtype=dtype)
        orc.saxpy(zref, dtype(2.5), X, Y, ndrange, wgs)
        assert np.array_equal(KA.Array(Z).view(np.uint8), zref.view(np.uint8))


@pytest.mark.parametrize("vec", [0, 1])
def test_unaligned_views_take_the_scalar_twin(orc, vec):
    tune(registry__force_twins=1, registry__twin_vec=vec)
    src = orc.gen_uniform_f32((5000,), 7)
    B = KA.to_device(src)
    A = KA.allocate(be, np.float32, 5000).fill_(-1.0)
    K.copy_kernel_(be, 256)(A.view(1, (4000,)), B.view(3, (4000,)), ndrange=4000)
    ref = np.full(5000, -1.0, dtype=np.float32)
    ref[1:4001] = src[3:4003]
    assert np.array_equal(KA.Array(A), ref)
