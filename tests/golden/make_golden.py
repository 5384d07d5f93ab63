#!/usr/bin/env python
"""Writes tests/golden/vectors.json: seeded input -> output vectors for every north-star kernel.

Provenance (stated because it bounds what these fixtures prove): the reference is Julia + PoCL and cannot be run in
this environment, so NO vector here was produced by the reference itself.  The vectors are of two kinds:
  * "reference_assertion": the expected value is what the reference's own test asserts (file:line cited), written out
    literally -- these pin the oracle AND the CUDA path to the reference;
  * "oracle_regression": the output of oracle/ka_oracle.c on a seeded input, frozen here (small ones literally, large
    ones as SHA-256 of the little-endian bytes) -- these pin the CUDA path and guard the oracle against drift.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ka_oracle as orc  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.asfortranarray(a).tobytes(order="F")).hexdigest()


def F(shape, dtype):
    return np.zeros(shape, dtype=dtype, order="F")


def main():
    orc.build()
    v = {"format": 1, "generator": "tests/golden/make_golden.py", "vectors": []}
    add = v["vectors"].append

    # ---- reference assertions, literally ----
    add({"name": "localmem", "kind": "reference_assertion", "cite": "test/localmem.jl:67-79",
         "launch": {"ndrange": [64], "workgroupsize": 16}, "expect_i64": [16 - (i % 16) for i in range(64)]})
    add({"name": "private", "kind": "reference_assertion", "cite": "test/private.jl:61-70",
         "launch": {"ndrange": [64], "workgroupsize": 16}, "expect_i64": [16 - (i % 16) for i in range(64)]})
    add({"name": "forloop_first_column", "kind": "reference_assertion", "cite": "test/private.jl:73-79",
         "launch": {"ndrange": [64], "workgroupsize": 64}, "input": "A = ones(Int64, 64, 64)", "expect_i64": [64] * 64})
    add({"name": "index_linear_global_16x16", "kind": "reference_assertion", "cite": "test/test.jl:116-122",
         "launch": {"ndrange": [256], "workgroupsize": 8}, "expect_i64": list(range(1, 257))})
    add({"name": "index_linear_local_8x2", "kind": "reference_assertion", "cite": "test/test.jl:124-129",
         "launch": {"ndrange": [16], "workgroupsize": 8}, "expect_i64": list(range(1, 9)) * 2})
    add({"name": "histogram_all_two", "kind": "reference_assertion", "cite": "examples/histogram.jl:78,99",
         "input": "fill(Int32(2), 512)", "expect_i32": [0, 512]})
    add({"name": "histogram_linear", "kind": "reference_assertion", "cite": "examples/histogram.jl:77,98",
         "input": "Int32.(1:1024)", "expect_i32": [1] * 1024})
    add({"name": "partition", "kind": "reference_assertion", "cite": "test/test.jl:14-44",
         "cases": [{"ndrange": [128], "workgroupsize": [64], "groups": 2, "dynamic": False},
                   {"ndrange": [129], "workgroupsize": [64], "groups": 3, "dynamic": True}]})

    # ---- oracle regression vectors on the seeded inputs of SURVEY.md 8(d) ----
    b = orc.gen_uniform_f32((256, 128), 1235)
    a = F((128, 256), np.float32)
    orc.naive_transpose(a, b)
    add({"name": "naive_transpose_256x128_f32", "kind": "oracle_regression", "cite": "examples/naive_transpose.jl:4-21",
         "input": {"gen": "uniform_f32", "shape": [256, 128], "seed": 1235, "sha256": sha(b)},
         "output": {"shape": [128, 256], "dtype": "float32", "sha256": sha(a), "first8_bits": a.ravel(order="F")[:8].view(np.uint32).tolist()}})

    inp = orc.gen_bins_i32(100003, 1236, 256)
    bins = np.zeros(256, dtype=np.int32)
    orc.histogram(bins, inp, 256)
    add({"name": "histogram_100003_256bins", "kind": "oracle_regression", "cite": "examples/histogram.jl:18-66",
         "input": {"gen": "bins_i32", "n": 100003, "seed": 1236, "nbins": 256, "sha256": sha(inp)},
         "output": {"dtype": "int32", "values": bins.tolist()}})
    assert np.array_equal(bins, orc.create_histogram(inp)[:256])

    A = orc.gen_uniform_f32((64, 48), 1237)
    B = orc.gen_uniform_f32((48, 32), 1238)
    Cm = F((64, 32), np.float32)
    orc.coalesced_matmul(Cm, A, B)
    Cn = F((64, 32), np.float32)
    orc.matmul_naive(Cn, A, B)
    add({"name": "coalesced_matmul_64x48x32_f32", "kind": "oracle_regression", "cite": "examples/performant_matmul.jl:15-77",
         "input": {"gen": "uniform_f32", "A": {"shape": [64, 48], "seed": 1237}, "B": {"shape": [48, 32], "seed": 1238}},
         "output": {"shape": [64, 32], "dtype": "float32", "sha256": sha(Cm), "first8_bits": Cm.ravel(order="F")[:8].view(np.uint32).tolist()}})
    add({"name": "matmul_naive_64x48x32_f32", "kind": "oracle_regression", "cite": "examples/matmul.jl:5-27",
         "input": {"gen": "uniform_f32", "A": {"shape": [64, 48], "seed": 1237}, "B": {"shape": [48, 32], "seed": 1238}},
         "output": {"shape": [64, 32], "dtype": "float32", "sha256": sha(Cn), "first8_bits": Cn.ravel(order="F")[:8].view(np.uint32).tolist()}})

    x = orc.gen_uniform_f32((4099,), 1)
    y = orc.gen_uniform_f32((4099,), 2)
    z = np.zeros(4099, dtype=np.float32)
    orc.saxpy(z, np.float32(3.1415), x, y, (4099,), 1024)
    add({"name": "saxpy_4099_f32", "kind": "oracle_regression", "cite": "benchmark/benchmarks.jl:29-32",
         "input": {"gen": "uniform_f32", "n": 4099, "seeds": [1, 2], "a": 3.1415},
         "output": {"dtype": "float32", "sha256": sha(z), "first8_bits": z[:8].view(np.uint32).tolist()}})

    src = orc.gen_uniform_f32((1 << 16,), 1234)
    add({"name": "gen_uniform_f32_65536_seed1234", "kind": "oracle_regression", "cite": "SURVEY.md 8(d) C1 generator",
         "output": {"dtype": "float32", "sha256": sha(src), "first8_bits": src[:8].view(np.uint32).tolist()}})

    out = os.path.join(ROOT, "tests", "golden", "vectors.json")
    with open(out, "w") as f:
        json.dump(v, f, indent=1)
    print("wrote", out, len(v["vectors"]), "vectors")


if __name__ == "__main__":
    main()
