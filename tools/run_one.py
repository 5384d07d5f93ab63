#!/usr/bin/env python
"""Run ONE kernel configuration a few times (the short command line handed to ncu; see profiles/README.md).

usage: run_one.py {copy|transpose|hist|sgemm|tcgemm|cublas_sgemm} [size] [mode] [iters]
Prints the CUDA-event time of the launches and the SM clock / power sampled through NVML while they run."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200ka as KA  # noqa: E402
import ka_oracle as orc  # noqa: E402
from b200ka import _lib  # noqa: E402

kind = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
be = KA.B200Backend()
for kv in filter(None, os.environ.get("B200_TUNE", "").split(",")):     # e.g. B200_TUNE=tcgemm.group_m=4,tcgemm.cta2=0
    k, v = kv.rsplit("=", 1)
    _lib.call("b200_tune_set", k.encode(), int(v))


def ev():
    e = C.c_void_p()
    _lib.call("b200_event_create", C.byref(e))
    return e


class Clocks:
    def __init__(self):
        self.s, self.p, self.stop = [], [], threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop.is_set():
            try:
                self.s.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.p.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)


def run(fn, work, unit):
    fn()
    KA.synchronize(be)
    clk = Clocks()
    th = None
    if clk.nv:
        th = threading.Thread(target=clk.run, daemon=True)
        th.start()
    a, b = ev(), ev()
    _lib.call("b200_event_record", a)
    for _ in range(iters):
        fn()
    _lib.call("b200_event_record", b)
    KA.synchronize(be)
    if th:
        clk.stop.set()
        th.join()
    ms = C.c_float()
    _lib.call("b200_event_elapsed_ms", a, b, C.byref(ms))
    per = ms.value / iters
    scale = 1e9 if unit == "GB/s" else 1e12
    print(f"{kind} size={size} mode={mode}: {per:.4f} ms/launch, {work / (per * 1e-3) / scale:.1f} {unit}; "
          f"sm_mhz median {np.median(clk.s) if clk.s else None} min {min(clk.s) if clk.s else None} "
          f"power_w max {max(clk.p) if clk.p else None} samples {len(clk.s)}", flush=True)


if kind == "copy":
    n = size or (1 << 26)
    s, d = KA.to_device(orc.gen_uniform_f32((n,), 1)), KA.zeros(be, np.float32, n)
    run(lambda: _lib.call("b200_copy", C.c_void_p(d.ptr), C.c_void_p(s.ptr), 4 * n), 8.0 * n, "GB/s")
elif kind == "copy_twin":
    # gpu_copy_kernel! through the generic NDRange lowering (registry.force_twins = 1): mode 0 = 128-bit vectorised twin,
    # 1 = one work-item per thread (native-grid / multiply-high index arithmetic), 2 = one work-item per thread, Int64 div/mod
    from b200ka import examples as EXc
    n = size or (1 << 26)
    s, d = KA.to_device(orc.gen_uniform_f32((n,), 1)), KA.zeros(be, np.float32, n)
    _lib.call("b200_tune_set", b"registry.force_twins", 1)
    _lib.call("b200_tune_set", b"registry.twin_vec", 1 if mode == 0 else 0)
    _lib.call("b200_tune_set", b"registry.ka_mode", 2 if mode == 2 else -1)
    run(lambda: EXc.mycopy_(d, s), 8.0 * n, "GB/s")
elif kind == "transpose":
    n = size or 16384
    b_, a_ = KA.to_device(orc.gen_uniform_f32((n, n), 1235)), KA.zeros(be, np.float32, n, n)
    run(lambda: _lib.call("b200_transpose", C.c_void_p(a_.ptr), C.c_void_p(b_.ptr), n, n, n, n, 4), 8.0 * n * n, "GB/s")
elif kind == "saxpy":
    n = size or (1 << 26)
    x, y = KA.to_device(orc.gen_uniform_f32((n,), 2)), KA.to_device(orc.gen_uniform_f32((n,), 3))
    z = KA.zeros(be, np.float32, n)
    run(lambda: _lib.call("b200_saxpy", C.c_void_p(z.ptr), C.c_double(3.1415), C.c_void_p(x.ptr), C.c_void_p(y.ptr), C.c_size_t(n), 0),
        12.0 * n, "GB/s")
elif kind in ("reduce_max", "reduce_sum", "mismatch", "norms", "gen_uniform", "gen_bins"):
    n = size or (1 << 28)
    if kind.startswith("reduce"):
        d = KA.rand_bins_(KA.allocate(be, np.int32, n), 1236, 256) if mode == 0 else KA.rand_uniform_(KA.allocate(be, np.float32, n), 7)
        f = KA.maximum if kind == "reduce_max" else KA.sum
        run(lambda: f(d), 4.0 * n, "GB/s")
    elif kind == "mismatch":
        a, b = KA.rand_uniform_(KA.allocate(be, np.float32, n), 7), KA.rand_uniform_(KA.allocate(be, np.float32, n), 7)
        run(lambda: KA.isequal(a, b), 8.0 * n, "GB/s")
    elif kind == "norms":
        a, b = KA.rand_uniform_(KA.allocate(be, np.float32, n), 7), KA.rand_uniform_(KA.allocate(be, np.float32, n), 8)
        run(lambda: KA.isapprox(a, b), 8.0 * n, "GB/s")
    elif kind == "gen_uniform":
        a = KA.allocate(be, np.float32, n)
        run(lambda: KA.rand_uniform_(a, 7), 4.0 * n, "GB/s")
    else:
        a = KA.allocate(be, np.int32, n)
        run(lambda: KA.rand_bins_(a, 7, 256), 4.0 * n, "GB/s")
elif kind == "hist":
    n = size or (1 << 30)
    d_in = KA.to_device(orc.gen_bins_i32(n, 1236, 256) if mode == 0 else np.full(n, 2, dtype=np.int32))
    bins = KA.zeros(be, np.int32, 256)
    run(lambda: _lib.call("b200_histogram_i32", C.c_void_p(bins.ptr), 256, C.c_void_p(d_in.ptr), n), 4.0 * n, "GB/s")
elif kind == "hist64":
    n = size or (1 << 29)
    d32 = KA.rand_bins_(KA.allocate(be, np.int32, n), 1236, 256)
    d_in = KA.to_device(KA.Array(d32).astype(np.int64))
    bins = KA.zeros(be, np.int64, 256)
    run(lambda: _lib.call("b200_histogram_i64", C.c_void_p(bins.ptr), 256, C.c_void_p(d_in.ptr), n), 8.0 * n, "GB/s")
elif kind == "hist_fused":
    # the fused histogram + peer all-reduce kernel with a world of one rank (the exchange degenerates to local reductions)
    from b200ka import multigpu as mg
    n = size or (1 << 30)
    d_in = KA.to_device(orc.gen_bins_i32(n, 1236, 256))
    bins = KA.zeros(be, np.int32, 256)
    peers = mg.PeerGroup(1, 0, lambda blob: [blob])
    run(lambda: peers.histogram_allreduce(bins, d_in), 4.0 * n, "GB/s")
    print("timeline ns after the counting loop [counts_merged, slots_sent, all_summed]:", peers.debug_stamps())
elif kind == "tf32gemm":
    n = size or 8192
    m = int(os.environ.get("RUN_M", n))      # rows of A / C (a row panel of the sharded problem)
    A, B = KA.to_device(orc.gen_uniform_f32((m, n), 1237)), KA.to_device(orc.gen_uniform_f32((n, n), 1238))
    Cd = KA.zeros(be, np.float32, m, n)
    run(lambda: _lib.call("b200_matmul_tf32", C.c_void_p(Cd.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), m, n, n, m, n, m),
        2.0 * m * n * n, "TFLOP/s")
elif kind == "tf32x3gemm":
    n = size or 8192
    m = int(os.environ.get("RUN_M", n))
    A, B = KA.to_device(orc.gen_uniform_f32((m, n), 1237)), KA.to_device(orc.gen_uniform_f32((n, n), 1238))
    Cd = KA.zeros(be, np.float32, m, n)
    ws = KA.allocate(be, np.float32, m * n + n * n)
    run(lambda: _lib.call("b200_matmul_f32_tf32x3", C.c_void_p(Cd.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), m, n, n, m, n, m, C.c_void_p(ws.ptr)),
        2.0 * m * n * n, "TFLOP/s")
elif kind in ("sgemm", "tcgemm"):
    n = size or 8192
    m = int(os.environ.get("RUN_M", n))      # rows of A / C (a row panel of the sharded problem)
    A, B = KA.to_device(orc.gen_uniform_f32((m, n), 1237)), KA.to_device(orc.gen_uniform_f32((n, n), 1238))
    Cd = KA.zeros(be, np.float32, m, n)
    for knob in ("sgemm.nj", "tcgemm.bn"):
        if os.environ.get(knob.upper().replace(".", "_")):
            _lib.call("b200_tune_set", knob.encode(), int(os.environ[knob.upper().replace(".", "_")]))
    if kind == "sgemm":
        if os.environ.get("SGEMM_VARIANT"):
            _lib.call("b200_tune_set", b"sgemm.variant", int(os.environ["SGEMM_VARIANT"]))
        if os.environ.get("SGEMM_ENGINE"):
            _lib.call("b200_tune_set", b"sgemm.engine", int(os.environ["SGEMM_ENGINE"]))
        run(lambda: _lib.call("b200_matmul_f32", C.c_void_p(Cd.ptr), C.c_void_p(A.ptr), C.c_void_p(B.ptr), m, n, n, m, n, m, mode),
            2.0 * m * n * n, "TFLOP/s")
    else:
        ws = KA.allocate(be, np.uint16, m * n + n * n)
        _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr), C.c_void_p(A.ptr), m, n, m, 1)
        _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr + 2 * m * n), C.c_void_p(B.ptr), n, n, n, 0)
        run(lambda: _lib.call("b200_matmul_bf16", C.c_void_p(Cd.ptr), C.c_void_p(ws.ptr), C.c_void_p(ws.ptr + 2 * m * n), m, n, n, m),
            2.0 * m * n * n, "TFLOP/s")
elif kind == "matmul_naive":
    # examples/matmul.jl through the registry: tiled UNFUSED route (mode 0) or the literal twin (mode 1)
    from b200ka import examples as EXm
    n = size or 4096
    A, B = KA.rand_uniform_(KA.allocate(be, np.float32, n, n), 1237), KA.rand_uniform_(KA.allocate(be, np.float32, n, n), 1238)
    Cd = KA.zeros(be, np.float32, n, n)
    _lib.call("b200_tune_set", b"registry.force_twins", mode)
    run(lambda: EXm.matmul_(Cd, A, B), 2.0 * n ** 3, "TFLOP/s")
elif kind == "cublas_sgemm":
    # library reference point only (never on the product path): cuBLAS FP32 SIMT GEMM through torch
    import torch
    n = size or 8192
    torch.backends.cuda.matmul.allow_tf32 = False
    a = torch.rand(n, n, device="cuda")
    b = torch.rand(n, n, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / iters
    print(f"cublas_sgemm n={n}: {per:.4f} ms, {2.0 * n ** 3 / (per * 1e-3) / 1e12:.1f} TFLOP/s")
