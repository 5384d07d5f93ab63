#!/usr/bin/env python
"""b200_fill variants: 128-bit stores (fill.impl=0) against the TMA bulk-store form (fill.impl=1) over chunk size and CTAs per SM;
CUDA-event time of 20 back-to-back fills, result checked."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
import b200ka as KA  # noqa: E402
from b200ka import _lib  # noqa: E402

be = KA.B200Backend()


def tune(k, v):
    _lib.call("b200_tune_set", k.encode(), int(v))


def ev():
    e = C.c_void_p()
    _lib.call("b200_event_create", C.byref(e))
    return e


def time_ms(fn, iters=20):
    for _ in range(3):
        fn()
    KA.synchronize(be)
    a, b = ev(), ev()
    _lib.call("b200_event_record", a)
    for _ in range(iters):
        fn()
    _lib.call("b200_event_record", b)
    KA.synchronize(be)
    ms = C.c_float()
    _lib.call("b200_event_elapsed_ms", a, b, C.byref(ms))
    return ms.value / iters


for logn in (26, 28):
    n = 1 << logn
    d = [KA.allocate(be, np.float32, n) for _ in range(2)]
    val = np.array([1.5], dtype=np.float32)
    k = [0]

    def f():
        k[0] ^= 1
        _lib.call("b200_fill", C.c_void_p(d[k[0]].ptr), n, 4, val.ctypes.data_as(C.c_void_p))

    cfgs = [(0, 0, 0, 0)] + [(1, kb, cps, sp) for sp in (0, 1) for kb in (32, 64, 128) for cps in (1, 2, 3, 4, 6, 8, 12, 16, 32)]
    for impl, kb, cps, sp in cfgs:
        tune("fill.impl", impl)
        if impl:
            tune("fill.chunk_kb", kb)
            tune("fill.ctas_per_sm", cps)
            tune("fill.span", sp)
        d[0].fill_(0.0)
        d[1].fill_(0.0)
        tune("fill.impl", impl)
        ms = time_ms(f)
        h = KA.Array(d[0])
        ok = bool(np.all(h == 1.5))
        print(f"fill 2^{logn} f32 impl={impl} span={sp} chunk_kb={kb} ctas_per_sm={cps}: {ms * 1e3:.1f} us {4.0 * n / ms / 1e6:.0f} GB/s ok={ok}", flush=True)
