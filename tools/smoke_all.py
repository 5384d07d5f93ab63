#!/usr/bin/env python
"""Small-shape pass over every hot kernel through the public API with on-device checks only (no oracle): a 2-second
"is the library alive" run for a fresh box.  (compute-sanitizer is closed on this pool; bounds are checked by the parity tests.)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
import b200ka as KA  # noqa: E402
from b200ka import _lib, examples as EX  # noqa: E402

be = KA.B200Backend()
u = lambda *shape, seed=1: KA.rand_uniform_(KA.allocate(be, np.float32, *shape), seed)  # noqa: E731

a, b = KA.zeros(be, np.float32, 1 << 16), u(1 << 16)
EX.mycopy_(a, b)
assert KA.isequal(a, b)
for (m, n) in [(256, 128), (100, 37)]:
    src, dst, back = u(n, m), KA.zeros(be, np.float32, m, n), KA.zeros(be, np.float32, n, m)
    EX.naive_transpose_(dst, src)
    EX.naive_transpose_(back, dst)
    assert KA.isequal(back, src)
inp = KA.rand_bins_(KA.allocate(be, np.int32, 100003), 1236, 256)
bins = KA.zeros(be, np.int32, 256)
EX.histogram_(bins, inp)
assert KA.sum(bins) == 100003 and KA.maximum(inp) == 256
for (M, K, N) in [(256, 64, 224), (130, 40, 70), (384, 32, 112)]:
    A, B = u(M, K, seed=2), u(K, N, seed=3)
    Cd = KA.zeros(be, np.float32, M, N)
    EX.performant_matmul_(Cd, A, B)
    ref = KA.matmul(A, B)
    assert KA.isequal(Cd, ref)
    if K % 8 == 0 and M % 4 == 0:
        T = KA.zeros(be, np.float32, M, N)
        EX.matmul_tf32_(T, A, B)
        assert KA.isapprox(T, ref, rtol=1e-2)
        Bf = KA.zeros(be, np.float32, M, N)
        EX.matmul_bf16_(Bf, A, B)
        assert KA.isapprox(Bf, ref, rtol=1e-2)
KA.synchronize(be)
print("smoke_all ok")
