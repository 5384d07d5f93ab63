#!/usr/bin/env python
"""Histogram counting-phase A/B on one GPU: contiguous slices vs interleaved 16 KiB units, at the 1-GPU and 8-GPU shard sizes,
local-only and fused (one rank as its own peer) kernels, with the fused kernel's timeline.  python tools/hist_probe.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
import b200ka as KA  # noqa: E402
from b200ka import _lib, examples as EX, multigpu as mg  # noqa: E402

be = KA.B200Backend()
ev0, ev1 = C.c_void_p(), C.c_void_p()
_lib.call("b200_event_create", C.byref(ev0))
_lib.call("b200_event_create", C.byref(ev1))


def timed(fn, steps=100, warm=10):
    for _ in range(warm):
        fn()
    KA.synchronize(be)
    _lib.call("b200_event_record", ev0)
    for _ in range(steps):
        fn()
    _lib.call("b200_event_record", ev1)
    KA.synchronize(be)
    ms = C.c_float()
    _lib.call("b200_event_elapsed_ms", ev0, ev1, C.byref(ms))
    return ms.value / steps * 1e3


def tune(k, v):
    _lib.call("b200_tune_set", k.encode(), v)


peers = mg.PeerGroup(1, 0, lambda blob: [blob])
for G in (1, 8):
    ne = (1 << 30) // G
    d_in = KA.rand_bins_(KA.allocate(be, np.int32, ne), 1236, 256)
    bins = KA.zeros(be, np.int32, 256)
    for inter, ll, npf, dyn in ((0, 0, 16, 0), (1, 0, 0, 0), (1, 0, 16, 0), (1, 0, 16, 4), (1, 0, 16, 8), (1, 0, 16, 16), (1, 0, 16, 32), (1, 0, 16, 64)):
        tune("hist.interleave", inter)
        tune("hist.fused_ll", ll)
        tune("hist.l2_prefetch_rounds", npf)
        tune("hist.dyn_rounds", dyn)
        us_l = timed(lambda: EX.histogram_(bins, d_in))
        us_f = timed(lambda: peers.histogram_allreduce(bins, d_in))
        tl = []
        for rep in range(3):
            peers.timeline()
            peers.histogram_allreduce(bins, d_in)
            tl.append(peers.timeline())
        spread = [x["slowest_cta_done"] for x in tl]
        tail = [x["all_summed"] for x in tl]
        print(f"histogram 2^30/{G} interleave={inter} fused_ll={ll} l2_prefetch_rounds={npf} dyn_rounds={dyn}: local-only {us_l:8.2f} us ({4 * ne / us_l / 1e3:6.0f} GB/s), fused 1 rank {us_f:8.2f} us; "
              f"stream time at 7.25 TB/s {4 * ne / 7.25e12 * 1e6:7.2f} us; fastest->slowest CTA {spread} ns, slowest CTA -> summed {tail} ns", flush=True)
    tune("hist.interleave", -1)
    for stages in (2, 4, 6):
        tune("hist.impl", 2)
        tune("hist.tma_stages", stages)
        us_t = timed(lambda: EX.histogram_(bins, d_in))
        tune("hist.impl", 1)
        print(f"histogram 2^30/{G} TMA feed (hist.impl=2, {stages} x 16 KiB ring): local-only {us_t:8.2f} us ({4 * ne / us_t / 1e3:6.0f} GB/s)", flush=True)
    del d_in
tune("hist.interleave", -1)
tune("hist.fused_ll", 0)
tune("hist.l2_prefetch_rounds", 16)
tune("hist.dyn_rounds", 0)
peers.free()
