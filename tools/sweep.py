#!/usr/bin/env python
"""Tuning sweep over the kernel variants behind the C ABI (run on the B200 box; writes gpurun_out/sweep.json).

Each configuration is timed with CUDA events on the library stream (3 warm-ups, median of N launches, inputs larger
than L2 or alternating buffers) and reported as algorithmic GB/s or TFLOP/s.  Results decide the tuning defaults
compiled into csrc/ (see DESIGN.md)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200ka as KA  # noqa: E402
import ka_oracle as orc  # noqa: E402
from b200ka import _lib  # noqa: E402

be = KA.B200Backend()
only = set(sys.argv[1:])
results = []


def tune(**kv):
    for k, v in kv.items():
        _lib.call("b200_tune_set", k.replace("__", ".").encode(), int(v))


def ev():
    e = C.c_void_p()
    _lib.call("b200_event_create", C.byref(e))
    return e


def time_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    KA.synchronize(be)
    evs = [(ev(), ev()) for _ in range(iters)]
    for a, b in evs:
        _lib.call("b200_event_record", a)
        fn()
        _lib.call("b200_event_record", b)
    KA.synchronize(be)
    out = []
    for a, b in evs:
        ms = C.c_float()
        _lib.call("b200_event_elapsed_ms", a, b, C.byref(ms))
        out.append(ms.value)
        _lib.call("b200_event_destroy", a)
        _lib.call("b200_event_destroy", b)
    return float(np.median(out)), float(np.min(out))


def record(kind, cfg, med, mn, work, unit):
    scale = 1e9 if unit == "GB/s" else 1e12
    r = dict(kind=kind, cfg=cfg, ms_median=med, ms_min=mn, value=work / (med * 1e-3) / scale, best=work / (mn * 1e-3) / scale, unit=unit)
    results.append(r)
    print(f"{kind:10s} {json.dumps(cfg):70s} median {med:9.4f} ms  {r['value']:9.1f} {unit}  (best {r['best']:.1f})", flush=True)


def want(name):
    return not only or name in only


# ------------------------------------------------------------------------------------------------ copy
if want("copy"):
    for logn in (26, 28):
        n = 1 << logn
        src = [KA.to_device(orc.gen_uniform_f32((n,), 1 + i)) for i in range(2)]
        dst = [KA.zeros(be, np.float32, n) for _ in range(2)]
        st = {"i": 0}

        def step():
            i = st["i"] & 1
            _lib.call("b200_copy", C.c_void_p(dst[i].ptr), C.c_void_p(src[i].ptr), 4 * n)
            st["i"] += 1

        def step_memcpy():
            i = st["i"] & 1
            _lib.call("b200_memcpy_d2d_async", C.c_void_p(dst[i].ptr), C.c_void_p(src[i].ptr), 4 * n)
            st["i"] += 1

        work = 2 * 4 * n
        med, mn = time_ms(step_memcpy)
        record("copy", dict(n=f"2^{logn}", impl="cudaMemcpyAsync"), med, mn, work, "GB/s")
        for unroll in (2, 4, 8):
            for cps in (4, 8, 16, 32):
                tune(copy__impl=0, copy__unroll=unroll, copy__ctas_per_sm=cps)
                med, mn = time_ms(step)
                record("copy", dict(n=f"2^{logn}", impl="ldg", unroll=unroll, ctas_per_sm=cps), med, mn, work, "GB/s")
        for stage_kb, stages, cps in ((16, 4, 1), (16, 8, 1), (32, 4, 1), (32, 6, 1), (64, 3, 1), (16, 4, 2), (32, 3, 2), (16, 6, 2), (16, 3, 4), (8, 6, 4)):
            tune(copy__impl=1, copy__bulk_stage_kb=stage_kb, copy__bulk_stages=stages, copy__bulk_ctas_per_sm=cps)
            try:
                med, mn = time_ms(step)
                record("copy", dict(n=f"2^{logn}", impl="bulk", stage_kb=stage_kb, stages=stages, ctas_per_sm=cps), med, mn, work, "GB/s")
            except Exception as e:
                print("copy bulk cfg failed:", stage_kb, stages, cps, e)
        tune(copy__impl=0, copy__unroll=4, copy__ctas_per_sm=8, copy__bulk_stage_kb=32, copy__bulk_stages=4, copy__bulk_ctas_per_sm=1)
        # fill (store only)
        val = np.array([1.0], dtype=np.float32)
        med, mn = time_ms(lambda: _lib.call("b200_fill", C.c_void_p(dst[0].ptr), n, 4, val.ctypes.data_as(C.c_void_p)))
        record("fill", dict(n=f"2^{logn}"), med, mn, 4 * n, "GB/s")
        del src, dst

# ------------------------------------------------------------------------------------------- transpose
if want("transpose"):
    for n in (16384, 8192):
        b = KA.to_device(orc.gen_uniform_f32((n, n), 1235))
        a = KA.zeros(be, np.float32, n, n)
        work = 2 * 4 * n * n
        for order in (0, 1):
            tune(transpose__order=order, transpose__force_generic=0)
            med, mn = time_ms(lambda: _lib.call("b200_transpose", C.c_void_p(a.ptr), C.c_void_p(b.ptr), n, n, n, n, 4))
            record("transpose", dict(n=n, impl="tile64", order=order), med, mn, work, "GB/s")
        tune(transpose__force_generic=1)
        med, mn = time_ms(lambda: _lib.call("b200_transpose", C.c_void_p(a.ptr), C.c_void_p(b.ptr), n, n, n, n, 4))
        record("transpose", dict(n=n, impl="generic32"), med, mn, work, "GB/s")
        tune(transpose__force_generic=0, transpose__order=0, registry__force_twins=1)
        from b200ka import examples as EX
        med, mn = time_ms(lambda: EX.naive_transpose_(a, b), iters=5)
        record("transpose", dict(n=n, impl="literal naive twin (reference access pattern)"), med, mn, work, "GB/s")
        tune(registry__force_twins=0)
        del a, b

# ------------------------------------------------------------------------------------------- histogram
if want("hist"):
    n = 1 << 28
    work = 4 * n
    for dist_name in ("uniform", "all_equal"):
        inp = orc.gen_bins_i32(n, 1236, 256) if dist_name == "uniform" else np.full(n, 2, dtype=np.int32)
        d_in = KA.to_device(inp)
        want_bins = np.bincount(inp[: 1 << 22], minlength=257)[1:]
        bins = KA.zeros(be, np.int32, 256)

        def step():
            _lib.call("b200_histogram_i32", C.c_void_p(bins.ptr), 256, C.c_void_p(d_in.ptr), n)

        cfgs = [dict(hist__impl=0, hist__threads=t, hist__batch=p, hist__ilp=i, hist__ctas_per_sm=c)
                for (t, p, i, c) in [(512, 8, 2, 1), (512, 8, 1, 1), (512, 4, 2, 1), (512, 4, 1, 1), (768, 4, 2, 1), (768, 4, 1, 1),
                                     (768, 8, 2, 1), (256, 8, 2, 1), (256, 8, 2, 2), (256, 8, 2, 3), (256, 8, 1, 3)]]
        cfgs += [dict(hist__impl=1, hist__atomic_threads=t, hist__atomic_batch=p, hist__atomic_ctas_per_sm=c)
                 for (t, p, c) in [(256, 4, 4), (256, 4, 8), (256, 8, 8), (512, 4, 2), (512, 4, 4), (512, 8, 4), (1024, 4, 1), (1024, 4, 2)]]
        for cfg in cfgs:
            tune(**cfg)
            try:
                med, mn = time_ms(step, iters=10)
                record("hist", dict(dist=dist_name, **{k.split("__")[1]: v for k, v in cfg.items()}), med, mn, work, "GB/s")
            except Exception as e:
                print("hist cfg failed:", cfg, e)
        tune(hist__impl=0, hist__threads=512, hist__batch=8, hist__ilp=2, hist__ctas_per_sm=1)
        del d_in, bins

# ----------------------------------------------------------------------------------------------- saxpy
if want("saxpy"):
    for n in (1 << 26, 1 << 28):
        X, Y, Z = (KA.to_device(orc.gen_uniform_f32((n,), 5 + i)) for i in range(3))
        for cps in (2, 4, 8, 16, 32):
            tune(saxpy__ctas_per_sm=cps)
            med, mn = time_ms(lambda: _lib.call("b200_saxpy", C.c_void_p(Z.ptr), C.c_double(2.0), C.c_void_p(X.ptr), C.c_void_p(Y.ptr), n, 0))
            record("saxpy", dict(n=n, ctas_per_sm=cps), med, mn, 12.0 * n, "GB/s")
        del X, Y, Z
    tune(saxpy__ctas_per_sm=32)

# ----------------------------------------------------------------------------------------------- host pipeline
if want("hostpipe"):
    from b200ka import examples as EX
    n = 16384
    hp = [C.c_void_p(), C.c_void_p()]
    for h in hp:
        _lib.call("b200_host_alloc", C.byref(h), 4 * n * n)
    h_in = np.ctypeslib.as_array((C.c_float * (n * n)).from_address(hp[0].value)).reshape((n, n), order="F")
    h_out = np.ctypeslib.as_array((C.c_float * (n * n)).from_address(hp[1].value)).reshape((n, n), order="F")
    h_in[...] = orc.gen_uniform_f32((n, n), 1235)
    for mib in (4, 8, 16, 32, 64, 128):
        tune(hostpipe__slab_mib=mib)
        med, mn = time_ms(lambda: EX.naive_transpose_host_(be, h_out, h_in), iters=5, warm=2)
        record("hostpipe_transpose", dict(n=n, slab_mib=mib), med, mn, 8.0 * n * n, "GB/s")
    assert np.array_equal(h_out[7, :100], h_in[:100, 7])
    tune(hostpipe__slab_mib=32)
    nh = 1 << 28
    hh = C.c_void_p()
    _lib.call("b200_host_alloc", C.byref(hh), 4 * nh)
    hin = np.ctypeslib.as_array((C.c_int32 * nh).from_address(hh.value))
    hin[...] = orc.gen_bins_i32(nh, 1236, 256)
    bins = np.zeros(256, dtype=np.int32)
    for mib in (8, 32, 128):
        tune(hostpipe__slab_mib=mib)
        med, mn = time_ms(lambda: EX.histogram_host_(be, bins, hin), iters=5, warm=2)
        record("hostpipe_hist", dict(n=nh, slab_mib=mib), med, mn, 4.0 * nh, "GB/s")
    tune(hostpipe__slab_mib=32)
    for h in hp + [hh]:
        _lib.call("b200_host_free", h)

# ----------------------------------------------------------------------------------------------- sgemm
if want("sgemm"):
    n = 8192
    A = KA.to_device(orc.gen_uniform_f32((n, n), 1237))
    B = KA.to_device(orc.gen_uniform_f32((n, n), 1238))
    Cd = KA.zeros(be, np.float32, n, n)
    work = 2.0 * n ** 3
    for mode in (0, 1):
        for variant in (2, 4):       # 0-3: register-staged engine variants (bit0 FFMA2, bit1 lane map); 4: TMA engine
            for gm in (8,) if variant != 4 else (4, 8, 16, 32):
                tune(sgemm__group_m=gm, sgemm__variant=variant & 3, sgemm__engine=1 if variant == 4 else 0)
                med, mn = time_ms(lambda: _lib.call("b200_matmul_f32", C.c_void_p(Cd.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), n, n, n, n, n, n, mode),
                                  iters=4, warm=1)
                record("sgemm", dict(n=n, mode=["tile16", "running"][mode], group_m=gm, variant=variant), med, mn, work, "TFLOP/s")
    tune(sgemm__group_m=8, sgemm__variant=2, sgemm__engine=1)
    # ------------------------------------------------------------------------------------------ tcgen05
    if want("tc"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_tc_check.py")], capture_output=True, text=True, timeout=300)
        print(r.stdout[-3000:], r.stderr[-2000:])
        if r.returncode == 0:
            ws = KA.allocate(be, np.uint16, 2 * n * n)
            _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr), C.c_void_p(A.ptr), n, n, n, 1)
            _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr + 2 * n * n), C.c_void_p(B.ptr), n, n, n, 0)
            for gm in (4, 8, 16, 32, 64):
                tune(tcgemm__group_m=gm)
                med, mn = time_ms(lambda: _lib.call("b200_matmul_bf16", C.c_void_p(Cd.ptr), C.c_void_p(ws.ptr), C.c_void_p(ws.ptr + 2 * n * n), n, n, n, n), iters=10)
                record("tcgemm", dict(n=n, group_m=gm), med, mn, work, "TFLOP/s")
            med, mn = time_ms(lambda: _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr), C.c_void_p(A.ptr), n, n, n, 1), iters=5)
            record("convert", dict(n=n, transpose=1), med, mn, 6.0 * n * n, "GB/s")
            med, mn = time_ms(lambda: _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr), C.c_void_p(B.ptr), n, n, n, 0), iters=5)
            record("convert", dict(n=n, transpose=0), med, mn, 6.0 * n * n, "GB/s")
        else:
            print("tcgen05 check FAILED; skipping its sweep")

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = "_".join(sorted(only)) if only else "all"
json.dump(results, open(os.path.join(ROOT, "gpurun_out", f"sweep_{tag}.json"), "w"), indent=1)
print("sweep done", len(results), "configs")
