"""tests/golden/vectors.json against the oracle (CPU, always) and against the CUDA path through the C ABI (`-m gpu`).
Provenance of the vectors: tests/golden/make_golden.py (reference assertions written out literally + frozen oracle
outputs on the seeded inputs of SURVEY.md 8(d); the Julia reference cannot be run here)."""
import hashlib
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = {v["name"]: v for v in json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["vectors"]}


def sha(a):
    return hashlib.sha256(np.asfortranarray(a).tobytes(order="F")).hexdigest()


def F(shape, dtype):
    return np.zeros(shape, dtype=dtype, order="F")


def check_output(got, spec):
    if "values" in spec:
        assert got.tolist() == spec["values"]
    if "sha256" in spec:
        assert got.ravel(order="F")[:8].view(np.uint32).tolist() == spec["first8_bits"]
        assert sha(got) == spec["sha256"]


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_reference_assertions(orc):
    A = np.full(64, -1, dtype=np.int64)
    orc.localmem(0, A, (64,), 16)
    assert A.tolist() == VEC["localmem"]["expect_i64"]
    A = np.full(64, -1, dtype=np.int64)
    orc.private(A, (64,), 16)
    assert A.tolist() == VEC["private"]["expect_i64"]
    A = np.ones((64, 64), dtype=np.int64, order="F")
    orc.forloop(A, (64,), 64)
    assert A[:, 0].tolist() == VEC["forloop_first_column"]["expect_i64"]
    A = F((16, 16), np.int64)
    orc.index_linear(0, A, 256, 8)
    assert A.ravel(order="F").tolist() == VEC["index_linear_global_16x16"]["expect_i64"]
    A = F((8, 2), np.int64)
    orc.index_linear(1, A, 16, 8)
    assert A.ravel(order="F").tolist() == VEC["index_linear_local_8x2"]["expect_i64"]
    for name, inp in (("histogram_all_two", np.full(512, 2, dtype=np.int32)), ("histogram_linear", np.arange(1, 1025, dtype=np.int32))):
        want = VEC[name]["expect_i32"]
        out = np.zeros(len(want), dtype=np.int32)
        orc.histogram(out, inp, 256)
        assert out.tolist() == want and orc.create_histogram(inp).tolist() == want
    for case in VEC["partition"]["cases"]:
        blocks, _, dyn = orc.partition(case["ndrange"], case["workgroupsize"])
        assert int(np.prod(blocks)) == case["groups"] and dyn is case["dynamic"]


def test_oracle_regression_vectors(orc):
    v = VEC["naive_transpose_256x128_f32"]
    b = orc.gen_uniform_f32((256, 128), 1235)
    assert sha(b) == v["input"]["sha256"]
    a = F((128, 256), np.float32)
    orc.naive_transpose(a, b)
    check_output(a, v["output"])
    v = VEC["histogram_100003_256bins"]
    inp = orc.gen_bins_i32(100003, 1236, 256)
    assert sha(inp) == v["input"]["sha256"]
    bins = np.zeros(256, dtype=np.int32)
    orc.histogram(bins, inp, 256)
    check_output(bins, v["output"])
    A, B = orc.gen_uniform_f32((64, 48), 1237), orc.gen_uniform_f32((48, 32), 1238)
    Cm = F((64, 32), np.float32)
    orc.coalesced_matmul(Cm, A, B)
    check_output(Cm, VEC["coalesced_matmul_64x48x32_f32"]["output"])
    Cn = F((64, 32), np.float32)
    orc.matmul_naive(Cn, A, B)
    check_output(Cn, VEC["matmul_naive_64x48x32_f32"]["output"])
    z = np.zeros(4099, dtype=np.float32)
    orc.saxpy(z, np.float32(3.1415), orc.gen_uniform_f32((4099,), 1), orc.gen_uniform_f32((4099,), 2), (4099,), 1024)
    check_output(z, VEC["saxpy_4099_f32"]["output"])
    check_output(orc.gen_uniform_f32((1 << 16,), 1234), VEC["gen_uniform_f32_65536_seed1234"]["output"])


# ------------------------------------------------------------------ CUDA path through the C ABI
@pytest.mark.gpu
def test_cuda_path_matches_golden_vectors():
    # inputs are generated ON THE DEVICE (b200_rand_*), so this test does not touch the oracle at all
    import b200ka as KA
    from b200ka import examples as EX, kernels

    be = KA.B200Backend()
    v = VEC["naive_transpose_256x128_f32"]
    b = KA.rand_uniform_(KA.allocate(be, np.float32, 256, 128), 1235)
    assert sha(KA.Array(b)) == v["input"]["sha256"]
    a = KA.zeros(be, np.float32, 128, 256)
    EX.naive_transpose_(a, b)
    check_output(KA.Array(a), v["output"])

    v = VEC["histogram_100003_256bins"]
    inp = KA.rand_bins_(KA.allocate(be, np.int32, 100003), 1236, 256)
    assert sha(KA.Array(inp)) == v["input"]["sha256"]
    bins = KA.zeros(be, np.int32, 256)
    EX.histogram_(bins, inp)
    check_output(KA.Array(bins), v["output"])
    for name, host in (("histogram_all_two", np.full(512, 2, dtype=np.int32)), ("histogram_linear", np.arange(1, 1025, dtype=np.int32))):
        d = EX.move(be, host)
        out = KA.zeros(be, np.int32, int(KA.maximum(d)))
        EX.histogram_(out, d)
        assert KA.Array(out).tolist() == VEC[name]["expect_i32"]

    A = KA.rand_uniform_(KA.allocate(be, np.float32, 64, 48), 1237)
    B = KA.rand_uniform_(KA.allocate(be, np.float32, 48, 32), 1238)
    Cm = KA.zeros(be, np.float32, 64, 32)
    EX.performant_matmul_(Cm, A, B)
    check_output(KA.Array(Cm), VEC["coalesced_matmul_64x48x32_f32"]["output"])
    Cn = KA.zeros(be, np.float32, 64, 32)
    EX.matmul_(Cn, A, B)
    check_output(KA.Array(Cn), VEC["matmul_naive_64x48x32_f32"]["output"])

    x = KA.rand_uniform_(KA.allocate(be, np.float32, 4099), 1)
    y = KA.rand_uniform_(KA.allocate(be, np.float32, 4099), 2)
    z = KA.zeros(be, np.float32, 4099)
    kernels.saxpy_kernel_(be, 1024)(z, np.float32(3.1415), x, y, ndrange=(4099,))
    KA.synchronize(be)
    check_output(KA.Array(z), VEC["saxpy_4099_f32"]["output"])
    check_output(KA.Array(KA.rand_uniform_(KA.allocate(be, np.float32, 1 << 16), 1234)), VEC["gen_uniform_f32_65536_seed1234"]["output"])

    A64 = KA.to_device(np.full(64, -1, dtype=np.int64))
    kernels.localmem(be, 16)(A64, ndrange=(64,))
    KA.synchronize(be)
    assert KA.Array(A64).tolist() == VEC["localmem"]["expect_i64"]
    P = KA.to_device(np.full(64, -1, dtype=np.int64))
    kernels.private(be, 16)(P, ndrange=(64,))
    KA.synchronize(be)
    assert KA.Array(P).tolist() == VEC["private"]["expect_i64"]
