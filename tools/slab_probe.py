#!/usr/bin/env python
"""Per-launch fixed cost of the slab kernels at the 8-GPU shard sizes, measured on ONE GPU (what limits t(1)/t(8)):
  * transpose of the (2048 x 16384) rows-of-b slab -> 16384 x 2048, with the residency cap and PDL knobs;
  * copy of 2^23 Float32; histogram of 2^27 Int32 with the fused kernel's timeline (one rank, itself as only peer).
    python tools/slab_probe.py > gpurun_out/slab_probe.txt
"""
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


def timed(fn, steps=200, warm=10, graph=True):
    if graph:
        fn()
        KA.synchronize(be)
        g = KA.Graph(be)
        with g:
            fn()
        run = g.launch
    else:
        run = fn
    for _ in range(warm):
        run()
    KA.synchronize(be)
    _lib.call("b200_event_record", ev0)
    for _ in range(steps):
        run()
    _lib.call("b200_event_record", ev1)
    KA.synchronize(be)
    ms = C.c_float()
    _lib.call("b200_event_elapsed_ms", ev0, ev1, C.byref(ms))
    return ms.value / steps * 1e3   # us


def tune(k, v):
    _lib.call("b200_tune_set", k.encode(), v)


n = 16384
for G in (1, 2, 4, 8):
    rb = n // G
    dB = KA.rand_uniform_(KA.allocate(be, np.float32, rb, n), 1235)
    dA = KA.zeros(be, np.float32, n, rb)
    ideal = 2 * 4 * n * rb / 6545.9e9 * 1e6
    for pdl, pf in ((1, 0), (1, 4), (1, 8), (1, 16), (1, 100000)):
        tune("transpose.pdl", pdl)
        tune("transpose.l2_prefetch", pf)
        us = timed(lambda: EX.naive_transpose_(dA, dB))
        print(f"transpose slab 1/{G}: pdl={pdl} l2_prefetch={pf}: {us:8.2f} us  ({2 * 4 * n * rb / us / 1e3:7.1f} GB/s; at the measured peak {ideal:6.2f} us)", flush=True)
    tune("transpose.l2_prefetch", 1)
    tune("transpose.pdl", 1)
    tune("transpose.ctas_per_sm", 0)
    tune("transpose.impl", 1)
    for stages, cps in [(3, 2)]:
        tune("transpose.tma_stages", stages)
        tune("transpose.tma_ctas_per_sm", cps)
        us = timed(lambda: EX.naive_transpose_(dA, dB))
        print(f"transpose slab 1/{G}: TMA variant stages={stages} ctas_per_sm={cps}: {us:8.2f} us  ({2 * 4 * n * rb / us / 1e3:7.1f} GB/s)", flush=True)
    tune("transpose.impl", 0)
    del dA, dB

for G in (1, 2, 4, 8):
    ne = (1 << 26) // G
    npairs = max(2, (512 << 20) // (8 * ne))
    src = [KA.rand_uniform_(KA.allocate(be, np.float32, ne), 1 + i) for i in range(npairs)]
    dst = [KA.zeros(be, np.float32, ne) for _ in range(npairs)]

    def cycle():
        for i in range(npairs):
            EX.mycopy_(dst[i], src[i])
    tune("copy.l2_prefetch", 0)
    us0 = timed(cycle, steps=20) / npairs
    tune("copy.l2_prefetch", 1)
    print(f"copy slab 1/{G} without the L2 prefetch: {us0:8.2f} us", flush=True)
    us = timed(cycle, steps=20) / npairs
    print(f"copy slab 1/{G} ({npairs} pairs rotated): {us:8.2f} us ({8 * ne / us / 1e3:7.1f} GB/s; at the measured peak {8 * ne / 6545.9e9 * 1e6:6.2f} us)", flush=True)
    if True:
        for impl, kb, stages, cps in [(1, 16, 4, 1), (1, 16, 6, 1), (1, 16, 8, 1), (1, 32, 4, 1), (1, 32, 6, 1), (1, 8, 8, 1), (0, 0, 0, 0)]:
            tune("copy.impl", impl)
            if impl:
                tune("copy.bulk_stage_kb", kb)
                tune("copy.bulk_stages", stages)
                tune("copy.bulk_ctas_per_sm", cps)
            us = timed(cycle, steps=20) / npairs
            print(f"   copy.impl={impl} stage_kb={kb} stages={stages} ctas_per_sm={cps}: {us:8.2f} us ({8 * ne / us / 1e3:7.1f} GB/s)", flush=True)
        tune("copy.impl", 1)
        tune("copy.bulk_stage_kb", 16)
        tune("copy.bulk_stages", 0)
        tune("copy.bulk_ctas_per_sm", 1)
    del src, dst

peers = mg.PeerGroup(1, 0, lambda blob: [blob])
for G in (1, 8):
    ne = (1 << 30) // G
    d_in = KA.rand_bins_(KA.allocate(be, np.int32, ne), 1236, 256)
    bins = KA.zeros(be, np.int32, 256)
    for inter in (0, 1):
        tune("hist.interleave", inter)
        us_l = timed(lambda: EX.histogram_(bins, d_in), steps=100)
        us_f = timed(lambda: peers.histogram_allreduce(bins, d_in), steps=100, graph=False)
        print(f"histogram slab 1/{G} interleave={inter}: local-only {us_l:8.2f} us, fused (1 rank) {us_f:8.2f} us; stream time at 7.25 TB/s {4 * ne / 7.25e12 * 1e6:7.2f} us", flush=True)
        KA.synchronize(be)
        for rep in range(3):
            peers.timeline()
            peers.histogram_allreduce(bins, d_in)
            print("   fused timeline (ns, relative to the fastest CTA leaving the counting loop):", peers.timeline(), flush=True)
    tune("hist.interleave", -1)
    del d_in
peers.free()
