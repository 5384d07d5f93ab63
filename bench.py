#!/usr/bin/env python
"""bench.py -- the north-star measurement (BASELINE.json): HBM GB/s of the KA hot path on B200.

Headline workload (configs[1]): naive_transpose 16384 x 16384 Float32 through B200Backend, one transpose per step.
  value     whole-job algorithmic GB/s (2 * 4 * n^2 bytes per step per GPU), inputs resident in HBM, CUDA-event timed
  e2e       the same metric through the public API with HOST buffers (pinned h2d + transpose + d2h inside the timer)
  roofline  achieved / measured HBM peak (MEASURED_PEAKS.json) for the transpose kernel
  cpu_baseline  the oracle's restatement of the reference CPU() algorithm on this host's cores (reported, not a target)
  suite     the other north-star kernels (copy 2^26, histogram 2^30 -> 256 bins [+ NCCL all-reduce when sharded],
            FP32 SIMT matmul 8192^3 [row-panel sharded], opt-in BF16 tcgen05 matmul), each with its own roofline fraction

`--impl reference` times the reference's own algorithm on the host CPU (the oracle port: the Julia/PoCL reference
cannot run here -- no Julia toolchain), same config/metric/unit/steps, on every host core it may use.
N>1 (torchrun): one rank per GPU; the headline is STRONG scaling of the ONE 16384^2 matrix of configs[1] in the
partition of SURVEY.md 8(e): rank g owns the output columns a[:, g n/G : (g+1) n/G) and holds rows g n/G ... of b as an
(n/G) x n column-major slab -- an ordinary (n/G) x n -> n x (n/G) transpose per GPU, no collective; barrier +
max-over-ranks timing, value = 2*4*n^2 bytes / that time.  A compact per-kernel summary of the suite is repeated under
roofline.suite.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))

N_T = 16384                      # transpose edge (configs[1])
BYTES_T = 2 * 4 * N_T * N_T      # algorithmic bytes per transpose
N_COPY = 1 << 26
N_HIST = 1 << 30
N_MM = 8192


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                        sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
        except Exception:
            pass
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, source="fallback")


def traffic_from_profiles(kernel_key, scale=1.0):
    """dram bytes per launch from the committed `ncu --set full` capture of the same kernel at the 1-GPU size
    (profiles/traffic.json, written by tools/summarize_profiles.py from the .ncu-rep), scaled to the launch timed here
    when that is a slab of it; None when the kernel has no capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            v = json.load(open(p)).get(kernel_key)
            return None if v is None else float(v) * scale
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.index = index
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            for bit, nm in names.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(0.004)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        if self.nv:
            self._once()
            self._stop.set()
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


class NvlinkCounter:
    """Hardware NVLink byte counters of one GPU through NVML (NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX, summed over all links,
    KiB): read before and after the timed steps, the difference is what actually crossed the links -- the proof that a fused
    compute + collective kernel moved its bytes over NVLink and how many.  None when the driver does not expose the counters."""

    def __init__(self, index):
        self.h = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.read()
        except Exception:
            self.h = None

    def read(self):
        nv = self.nv
        vals = nv.nvmlDeviceGetFieldValues(self.h, [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                    (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                raise RuntimeError("NVLink counter not available")
            out.append(int(v.value.ullVal) * 1024)
        return tuple(out)

    def delta_per_step(self, before, steps):
        if self.h is None or before is None:
            return None
        try:
            now = self.read()
        except Exception:
            return None
        return {"tx_bytes": (now[0] - before[0]) / steps, "rx_bytes": (now[1] - before[1]) / steps}

    def start(self):
        if self.h is None:
            return None
        try:
            return self.read()
        except Exception:
            return None


# =================================================================================================
# reference arm / cpu baseline: the oracle port of the reference CPU() algorithm
# =================================================================================================
METRIC = "naive_transpose HBM GB/s (16384x16384 Float32)"


def bench_config(world):
    """The `config` object of the JSON line -- the SAME for the B200 arm and the reference arm (the driver compares them)."""
    return {"workload": "naive_transpose 16384x16384 Float32 (examples/naive_transpose.jl), one matrix per step",
            "l2": "inputs larger than L2 (1 GiB in, 1 GiB out per step; 128 MiB + 128 MiB per GPU at 8 GPUs)",
            "parallelism": f"row-of-b slabs over {world} GPU(s), no collective (SURVEY.md 8(e))" if world > 1 else "1 GPU"}


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_oracle(all_cores=True):
    """The CPU oracle (test infrastructure: imported here only for input generation, the cpu_baseline leg and the reference arm).
    torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU arm must use every core it may, so the OpenMP thread count is set
    before libgomp initialises AND through omp_set_num_threads afterwards."""
    if all_cores:
        os.environ["OMP_NUM_THREADS"] = str(host_cores())
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ka_oracle as orc
    if all_cores:
        orc.set_threads(host_cores())
    return orc


def cpu_transpose_baseline(steps=3, warmup=1, budget_s=12.0):
    """`warmup` untimed + `steps` timed passes of the reference CPU() algorithm (oracle restatement: OpenMP over work-groups,
    work-items as inner loops; workgroup (256,1), ndrange (n,n) -- the launch shape of examples/naive_transpose.jl:18-19).
    A pass is the full 16384^2 matrix unless (warmup + steps) full passes would exceed `budget_s`: then every pass is the
    first `cols` output columns (a bounded sample of the same workload; bytes counted for the sample only)."""
    orc = load_oracle()
    n = N_T
    b = orc.gen_uniform_f32((n, n), 1235)
    a = np.zeros((n, n), dtype=np.float32, order="F")
    t0 = time.perf_counter()
    orc.naive_transpose_wg(a, b)                       # calibration pass (untimed; also first-touches `a`)
    t_full = time.perf_counter() - t0
    cols = n
    while cols > 1024 and (steps + warmup) * t_full * cols / n > budget_s:
        cols //= 2
    a_s, b_s = (a, b) if cols == n else (a[:, :cols], b[:cols, :])
    if cols != n:   # contiguous column-major slabs of the same matrices
        a_s, b_s = np.asfortranarray(a_s), np.asfortranarray(b_s)
    for _ in range(warmup):
        orc.naive_transpose_wg(a_s, b_s)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.naive_transpose_wg(a_s, b_s)
        times.append(time.perf_counter() - t0)
    ok = bool(np.array_equal(a_s[5, :64], b_s[:64, 5]))
    mean = float(np.mean(times))
    nbytes = 2 * 4 * n * cols
    sample = (f"full {n}x{n} Float32 naive_transpose" if cols == n else
              f"first {cols} of {n} output columns of the {n}x{n} Float32 naive_transpose ({nbytes >> 20} MiB of traffic per pass)")
    return dict(value=nbytes / mean / 1e9, unit="GB/s", cores=orc.threads(), kind="port",
                sample=sample + f" (oracle restatement of KA CPU(): OpenMP over work-groups, work-items as inner loops), "
                                f"mean of {steps} timed passes after {warmup} warm-up", seconds=mean, best_seconds=float(min(times)),
                result_checked=ok)


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    base = cpu_transpose_baseline(steps=args.steps, warmup=args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.gpus),
        "impl_detail": "reference CPU() algorithm restated in C (oracle port; the Julia/PoCL reference is not runnable here), "
                       f"{base['cores']} OpenMP threads on rank 0's host",
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "result_checked": base["result_checked"], "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))
    return 0


# =================================================================================================
# B200 arm
# =================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-suite", action="store_true", help="only the headline transpose line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` typed by hand: become the launch the contract describes (one rank per GPU, one node)
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                  "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    warmup = max(args.warmup, 3)

    import b200ka as KA
    from b200ka import _lib, examples as EX, multigpu as mg

    be = KA.B200Backend()
    if not KA.functional(be):
        raise SystemExit("bench.py: no CUDA device -- B200Backend has no CPU fallback")
    _lib.call("b200_set_device", local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        KA.synchronize(be)
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(flag):
        return bool(max_over_ranks(0.0 if flag else 1.0) == 0.0)

    def new_event():
        e = C.c_void_p()
        _lib.call("b200_event_create", C.byref(e))
        return e

    ev0, ev1 = new_event(), new_event()

    def timed(fn, steps, warm):
        """warm untimed steps, then `steps` steps between two CUDA events on the launch stream; max over ranks (ms/step)."""
        for _ in range(warm):
            fn()
        barrier()
        _lib.call("b200_launch_counter", None, 1)
        _lib.call("b200_event_record", ev0)
        for _ in range(steps):
            fn()
        _lib.call("b200_event_record", ev1)
        KA.synchronize(be)
        ms = C.c_float()
        _lib.call("b200_event_elapsed_ms", ev0, ev1, C.byref(ms))
        cnt = C.c_int64()
        _lib.call("b200_launch_counter", C.byref(cnt), 0)
        barrier()
        return max_over_ranks(ms.value) / steps, int(cnt.value)

    def graphed(fn):
        """One step recorded into a CUDA graph through the backend's own record/replay API (b200_graph_*): at 8 GPUs a step is
        ~45 us of device work, less than what the Python mirror needs to issue a launch; the replay is one submission."""
        fn()
        KA.synchronize(be)
        g = KA.Graph(be)
        with g:
            fn()
        return g.launch, g

    peaks = measured_peaks()
    orc = load_oracle(all_cores=(world == 1))  # input generation + cpu_baseline leg only

    # ---------------- headline: ONE transpose 16384^2 F32, rows-of-b slab per GPU (strong scaling, no collective) ----------------
    n = N_T
    j0, j1 = mg.transpose_shard(n, world, rank)
    rb = j1 - j0
    hb_full = orc.gen_uniform_f32((n, n), 1235)
    hb = np.asfortranarray(hb_full[j0:j1, :]) if world > 1 else hb_full        # (rb x n, ld = rb): rows j0..j1 of b
    del hb_full
    dB = KA.to_device(hb)
    dA = KA.zeros(be, np.float32, n, rb)                                        # a[:, j0:j1]

    def step_transpose():
        EX.naive_transpose_(dA, dB)

    step_transpose()
    KA.synchronize(be)
    slab_ok = all_ranks(bool(np.array_equal(KA.Array(dA), hb.T)))              # the whole slab, every rank, bit for bit
    use_graph = world > 1 and os.environ.get("B200_BENCH_GRAPH", "1") != "0"
    step_fn, graph_keep = (graphed(step_transpose) if use_graph else (step_transpose, None))
    with ClockSampler(local_rank) as clk:
        ms_step, launches = timed(step_fn, args.steps, warmup)
    value = BYTES_T / (ms_step * 1e-3) / 1e9                 # whole job: the one matrix / max-over-ranks time
    per_gpu = value / world
    plain = None
    if use_graph:                                            # the same step issued launch by launch from the Python mirror
        ms_plain, _ = timed(step_transpose, args.steps, warmup)
        plain = {"value": BYTES_T / (ms_plain * 1e-3) / 1e9, "ms_per_step": ms_plain}

    # ---------------- e2e: host buffers, copies inside the timed region ----------------
    e2e = None
    try:
        hp_in, hp_out = C.c_void_p(), C.c_void_p()
        nbytes = 4 * n * rb
        _lib.call("b200_host_alloc", C.byref(hp_in), nbytes)
        _lib.call("b200_host_alloc", C.byref(hp_out), nbytes)
        h_in = np.ctypeslib.as_array((C.c_float * (n * rb)).from_address(hp_in.value)).reshape((rb, n), order="F")
        h_out = np.ctypeslib.as_array((C.c_float * (n * rb)).from_address(hp_out.value)).reshape((n, rb), order="F")
        h_in[...] = hb

        def step_pipe():
            # the public host-array call: slabs of b go host -> device, are transposed, and come back device -> host,
            # with the three legs of consecutive slabs overlapped (csrc/hostpipe.cu)
            EX.naive_transpose_host_(be, h_out, h_in)

        def step_seq():
            # copyto! in, kernel, copyto! out on one stream
            KA.copyto_(be, dB, h_in)
            EX.naive_transpose_(dA, dB)
            KA.copyto_(be, h_out, dA)

        def time_e2e(fn, k):
            for _ in range(2):
                fn()
            barrier()
            _lib.call("b200_launch_counter", None, 1)
            _lib.call("b200_event_record", ev0)
            t0 = time.perf_counter()
            for _ in range(k):
                fn()
            _lib.call("b200_event_record", ev1)
            KA.synchronize(be)
            wall = time.perf_counter() - t0
            ms = C.c_float()
            _lib.call("b200_event_elapsed_ms", ev0, ev1, C.byref(ms))
            barrier()
            return max_over_ranks(max(ms.value * 1e-3, wall)) / k   # device events and host clock agree; keep the larger

        # ceiling of any host-array route on this host: the rank's slab h2d on one stream and d2h on another at the same time,
        # no kernel (PCIe is full duplex; with several ranks on one host the host side is shared)
        s2 = C.c_void_p()
        _lib.call("b200_stream_create", C.byref(s2), 0)

        def copy_only():
            KA.copyto_(be, dB, h_in)
            _lib.call("b200_set_stream", s2)
            KA.copyto_(be, h_out, dA)
            _lib.call("b200_set_stream", None)

        def copy_only_wait():
            KA.synchronize(be)
            _lib.call("b200_set_stream", s2)
            KA.synchronize(be)
            _lib.call("b200_set_stream", None)

        copy_only()
        copy_only_wait()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            copy_only()
        copy_only_wait()
        dt_copy = max_over_ranks(time.perf_counter() - t0) / 3
        _lib.call("b200_stream_destroy", s2)

        e2e_steps = max(1, min(args.e2e_steps, args.steps))
        dt_pipe = time_e2e(step_pipe, e2e_steps)
        h_out[...] = 0
        dt_seq = time_e2e(step_seq, e2e_steps)
        ok_seq = bool(np.array_equal(h_out[5, :64], hb[:64, 5]) and np.array_equal(h_out[n - 3, rb - 64:], hb[rb - 64:, n - 3]))
        # the route a user gets is the faster of the two on this host (with 8 ranks sharing one host the 3-stream slab
        # pipeline can lose to the plain in-order copies); both are reported
        h_out[...] = 0
        best_fn, best_dt, best_name = (step_pipe, dt_pipe, "pipelined") if dt_pipe <= dt_seq else (step_seq, dt_seq, "in-order")
        best_fn()
        KA.synchronize(be)
        ok = all_ranks(bool(np.array_equal(h_out, hb.T)))
        e2e = {"value": BYTES_T / best_dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
               "ms_per_step": best_dt * 1e3, "steps": e2e_steps, "result_checked": ok and ok_seq, "route": best_name,
               "pipelined_value": BYTES_T / dt_pipe / 1e9, "unpipelined_value": BYTES_T / dt_seq / 1e9,
               "copy_only_duplex_value": BYTES_T / dt_copy / 1e9,
               "limiter": "host <-> device copies: copy_only_duplex_value is the same slabs moved h2d and d2h concurrently with no kernel "
                          "at all (the ceiling of any host-array route on this host; one NUMA node and no PCIe topology are exposed to the VM)",
               "path": "pipelined = EX.naive_transpose_host_ -> b200_transpose_host (pinned host in/out, 32 MiB slabs, h2d | kernel | d2h "
                       "overlapped); in-order = KA.copyto! in, naive_transpose!, KA.copyto! out on one stream; every rank moves its own "
                       "slab of the one matrix; bytes are whole-job totals"}
        _lib.call("b200_host_free", hp_in)
        _lib.call("b200_host_free", hp_out)
    except Exception as ex:  # keep the headline line even if pinned allocation fails
        e2e = {"value": None, "unit": "GB/s", "error": str(ex)[:200]}

    slab_bytes = 2 * 4 * n * rb
    roofline = {"bound": "hbm", "achieved": per_gpu, "peak": peaks["hbm"], "unit": "GB/s", "frac": per_gpu / peaks["hbm"],
                "traffic": traffic_from_profiles("transpose64_f32_kernel", slab_bytes / BYTES_T), "peak_source": peaks["source"],
                "frac_of_nominal_8000": per_gpu / 8000.0, "kernel": "transpose64_f32_kernel",
                "algorithmic_bytes_per_launch": slab_bytes,
                "note": "per GPU: slab bytes / max-over-ranks step time.  peak = driver-measured copy; independent read + write streams "
                        "reach 6.7 TB/s of total traffic on this part, one direction alone 7.2-7.5 (tools/micro/hbm_stream.cu, "
                        "profiles/r04_hbm_stream.txt)"}

    # ---------------- suite: the other north-star kernels ----------------
    suite = {}
    if not args.no_suite:
        del dA, dB, graph_keep
        suite = run_suite(KA, _lib, EX, mg, orc, be, rank, world, dist, timed, graphed, all_ranks, peaks, min(args.steps, 50), warmup, barrier,
                          cpu=(rank == 0 and world == 1 and not args.no_cpu_baseline))
        roofline["suite"] = {k: {"value": v.get("value"), "unit": v.get("unit"), "ms": v.get("ms_per_step"),
                                 "frac": (v.get("roofline") or {}).get("frac"), "scaling": v.get("scaling")}
                             for k, v in suite.items() if isinstance(v, dict) and "value" in v}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_transpose_baseline(steps=3, warmup=1, budget_s=12.0)
        cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(world),
            "impl_detail": "B200Backend: EX.naive_transpose_ -> b200_launch -> transpose64_f32_kernel on the rank's slab"
                           + ("; one step = one CUDA-graph replay (b200_graph_launch) of that launch" if use_graph else ""),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
            "result_checked": slab_ok, "launch_by_launch": plain,
            "clocks": clk.summary(), "suite": suite,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_suite(KA, _lib, EX, mg, orc, be, rank, world, dist, timed, graphed, all_ranks, peaks, steps, warmup, barrier, cpu=False):
    out = {}
    # Sub-millisecond kernels are timed over at least 100 steps: the ranks leave the host-side barrier tens of microseconds apart,
    # a collective's first timed step absorbs that skew, and over 20 steps it would add 1-2 us to every step of an 85 us kernel
    # (bench 84.8 us vs 83.2 us for the same fused histogram timed over 50 steps by tests/gpu_multi_check.py, r05).
    steps_short = max(steps, 100)

    def cpu_time(fn, work, unit, scale, sample):
        """best of up to 3 passes of the oracle's work-group-vectorised restatement on the host cores (reported, not a target)"""
        best = float("inf")
        t_all = time.perf_counter()
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
            if time.perf_counter() - t_all > 6.0:
                break
        return {"value": work / best / scale, "unit": unit, "cores": orc.threads(), "kind": "port", "sample": sample}

    # ---- copy 2^26 F32 (configs[0] size), contiguous element slabs per rank (strong scaling, no collective).  Buffer pairs are
    # ---- rotated so that nothing is L2 resident: at 8 GPUs a slab pair is 64 MiB, so 8 pairs (512 MiB) go round ----
    try:
        n = N_COPY
        c0, c1 = mg.copy_shard(n, world, rank)
        ns = c1 - c0
        npairs = max(2, min(8, (512 << 20) // (8 * ns)))
        src = [KA.rand_uniform_(KA.allocate(be, np.float32, ns), 1234 + i) for i in range(npairs)]
        dst = [KA.zeros(be, np.float32, ns) for _ in range(npairs)]
        state = {"i": 0}

        def step_copy():
            i = state["i"] % npairs
            EX.mycopy_(dst[i], src[i])
            state["i"] += 1

        for _ in range(npairs):
            step_copy()
        KA.synchronize(be)
        ok = all_ranks(all(KA.isequal(d, s_) for d, s_ in zip(dst, src)))
        fn, keep = step_copy, None
        if world > 1:      # npairs consecutive steps in one graph (keeps the buffer rotation); a "step" stays one copy
            def cycle():
                for _ in range(npairs):
                    step_copy()
            fn, keep = graphed(cycle)
        k = max(1, steps_short // npairs) if world > 1 else steps_short
        ms, launches = timed(fn, k, warmup)
        if world > 1:
            ms /= npairs
        gbs = 2 * 4 * n / (ms * 1e-3) / 1e9          # whole job: 2^26 elements / max-over-ranks time
        out["copy_2^26_f32"] = {"value": gbs, "unit": "GB/s", "ms_per_step": ms, "gpu_launches": launches, "result_checked": ok, "steps": k * (npairs if world > 1 else 1),
                                "roofline": {"bound": "hbm", "achieved": gbs / world, "peak": peaks["hbm"], "frac": gbs / world / peaks["hbm"],
                                             "traffic": traffic_from_profiles("copy_bulk_kernel", 1.0 / world),
                                             "algorithmic_bytes_per_launch": 8 * ns},
                                "scaling": "strong", "buffer_pairs_rotated": npairs,
                                "launch": "CUDA-graph replay of one rotation" if world > 1 else "launch by launch"}
        if world == 1:
            # the literal twin (generic NDRange -> grid lowering of csrc/ka_device.cuh, no routing to b200_copy) in its three forms:
            # 128-bit vectorised, one work-item per thread with the strength-reduced index arithmetic, and one work-item per
            # thread with the reference's 64-bit div/mod per dimension (src/nditeration.jl:115-126)
            for key, knobs, note in (
                    ("copy_2^26_f32_literal_twin", {"registry.twin_vec": 1, "registry.ka_mode": -1},
                     "registry.force_twins=1: gpu_copy_kernel! through the generic lowering, 16 bytes (4 work-items) per thread"),
                    ("copy_2^26_f32_literal_twin_scalar", {"registry.twin_vec": 0, "registry.ka_mode": -1},
                     "registry.force_twins=1, registry.twin_vec=0: one work-item per thread, native-grid / multiply-high index arithmetic"),
                    ("copy_2^26_f32_literal_twin_div64", {"registry.twin_vec": 0, "registry.ka_mode": 2},
                     "registry.force_twins=1, registry.twin_vec=0, registry.ka_mode=2: one work-item per thread, the reference's Int64 div/mod")):
                _lib.call("b200_tune_set", b"registry.force_twins", 1)
                for k_, v_ in knobs.items():
                    _lib.call("b200_tune_set", k_.encode(), v_)
                try:
                    ms_t, _ = timed(step_copy, steps_short, warmup)
                finally:
                    _lib.call("b200_tune_set", b"registry.force_twins", 0)
                    _lib.call("b200_tune_set", b"registry.twin_vec", 1)
                    _lib.call("b200_tune_set", b"registry.ka_mode", -1)
                gbs_t = 2 * 4 * n / (ms_t * 1e-3) / 1e9
                out[key] = {"value": gbs_t, "unit": "GB/s", "ms_per_step": ms_t, "scaling": "strong", "of_tuned": gbs_t / gbs, "note": note,
                            "roofline": {"bound": "hbm", "achieved": gbs_t, "peak": peaks["hbm"], "frac": gbs_t / peaks["hbm"]}}
        del src, dst, keep
        if cpu:
            hB, hA = orc.gen_uniform_f32((n,), 1234), np.zeros(n, dtype=np.float32)
            out["copy_2^26_f32"]["cpu_baseline"] = cpu_time(lambda: orc.copy_wg(hA, hB, 1024), 8.0 * n, "GB/s", 1e9,
                                                            "full 2^26 Float32 copy_kernel!, workgroup 1024, OpenMP over work-groups")
    except Exception as ex:
        out["copy_2^26_f32"] = {"error": str(ex)[:300]}

    # ---- the N-d index arithmetic itself: naive_transpose_kernel! 8192^2 through its LITERAL twin (one work-item per thread,
    # ---- `i, j = @index(Global, NTuple)`, reads strided by a column as in the reference) under the three lowerings ----
    if world == 1:
        try:
            nt = 8192
            tb = KA.rand_uniform_(KA.allocate(be, np.float32, nt, nt), 99)
            ta = KA.zeros(be, np.float32, nt, nt)
            res = {}
            for name, mode in (("native_grid", 0), ("multiply_high", 1), ("div64_reference", 2)):
                _lib.call("b200_tune_set", b"registry.force_twins", 1)
                _lib.call("b200_tune_set", b"registry.ka_mode", mode)
                try:
                    ms_t, _ = timed(lambda: EX.naive_transpose_(ta, tb), max(5, steps // 4), 3)
                finally:
                    _lib.call("b200_tune_set", b"registry.force_twins", 0)
                    _lib.call("b200_tune_set", b"registry.ka_mode", -1)
                res[name] = {"value": 2.0 * 4 * nt * nt / (ms_t * 1e-3) / 1e9, "ms_per_step": ms_t}
            ok_t = bool(np.array_equal(KA.Array(ta)[:64, :64], KA.Array(tb)[:64, :64].T))
            out["naive_transpose_8192_literal_twin"] = {
                "value": res["native_grid"]["value"], "unit": "GB/s", "ms_per_step": res["native_grid"]["ms_per_step"], "scaling": "strong",
                "lowerings": res, "result_checked": ok_t,
                "note": "registry.force_twins=1: the reference's access pattern (32-byte sector per 4-byte element read), so far from the "
                        "tiled kernel by construction; the three entries differ only in how (i, j) is recovered per work-item",
                "roofline": {"bound": "hbm", "achieved": res["native_grid"]["value"], "peak": peaks["hbm"], "frac": res["native_grid"]["value"] / peaks["hbm"]}}
            del ta, tb
        except Exception as ex:
            out["naive_transpose_8192_literal_twin"] = {"error": str(ex)[:300]}

    # ---- SAXPY 2^26 F32 (benchmark/benchmarks.jl:29-70 kernel at a size beyond L2): 12 bytes per element ----
    try:
        n = N_COPY
        from b200ka import kernels as Kc
        X = [KA.to_device(orc.gen_uniform_f32((n,), 1300 + i)) for i in range(2)]
        Y = [KA.to_device(orc.gen_uniform_f32((n,), 1310 + i)) for i in range(2)]
        Z = [KA.zeros(be, np.float32, n) for _ in range(2)]
        kern = Kc.saxpy_kernel_(be, 1024)
        state = {"i": 0}

        def step_saxpy():
            i = state["i"] & 1
            kern(Z[i], np.float32(2.0), X[i], Y[i], ndrange=Z[i].shape)
            state["i"] += 1

        ms, launches = timed(step_saxpy, steps_short, warmup)
        gbs = 3 * 4 * n / (ms * 1e-3) / 1e9
        out["saxpy_2^26_f32"] = {"value": world * gbs, "unit": "GB/s", "ms_per_step": ms, "gpu_launches": launches,
                                 "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "frac": gbs / peaks["hbm"],
                                              "traffic": traffic_from_profiles("saxpy_vec_kernel"), "algorithmic_bytes_per_launch": 12 * n},
                                 "scaling": "weak"}
        del X, Y, Z
        if cpu:
            hX, hY, hZ = orc.gen_uniform_f32((n,), 1300), orc.gen_uniform_f32((n,), 1310), np.zeros(n, dtype=np.float32)
            out["saxpy_2^26_f32"]["cpu_baseline"] = cpu_time(lambda: orc.saxpy_wg(hZ, np.float32(2.0), hX, hY, 1024), 12.0 * n, "GB/s", 1e9,
                                                             "full 2^26 Float32 saxpy_kernel!, workgroup 1024, OpenMP over work-groups")
    except Exception as ex:
        out["saxpy_2^26_f32"] = {"error": str(ex)[:300]}

    # ---- distributed transpose of ONE 16384^2 Float32 matrix, column-sharded in and out (strong scaling, N > 1 only):
    # ---- the tile transpose and its all-to-all in one data kernel whose stores go to the peers' slabs over NVLink ----
    if world > 1:
        try:
            n = N_T
            rb = n // world
            b_loc = KA.rand_uniform_(KA.allocate(be, np.float32, n, rb), 1235 + rank)
            a_sym = mg.SymBuffer(n * rb * 4, world, rank, mg.torch_gather(dist))

            def step_dt():
                mg.transpose_alltoall_(a_sym, b_loc, n, n)

            step_dt()
            KA.synchronize(be)
            barrier()
            # every block of this rank's slab against its source, regenerated here from the peer's seed:
            # a_slab[s*rb:(s+1)*rb, :] == transpose(b_loc of rank s [rank*rb:(rank+1)*rb, :]) -- for s != rank those bytes crossed NVLink
            a_host = KA.Array(a_sym.array(np.float32, (n, rb)))
            tmp = KA.allocate(be, np.float32, n, rb)
            ok, peer_ok = True, True
            for s_ in range(world):
                blk = KA.Array(KA.rand_uniform_(tmp, 1235 + s_))[rank * rb:(rank + 1) * rb, :].T
                same = bool(np.array_equal(a_host[s_ * rb:(s_ + 1) * rb, :], blk))
                ok = ok and same
                if s_ != rank:
                    peer_ok = peer_ok and same
            del tmp, a_host
            ok, peer_ok = all_ranks(ok), all_ranks(peer_ok)
            nvl = NvlinkCounter(env_int("LOCAL_RANK", 0))
            nv0 = nvl.start()
            ms, launches = timed(step_dt, steps_short, warmup)
            nv_meas = nvl.delta_per_step(nv0, steps_short + warmup)
            gbs = 2.0 * 4 * n * n / (ms * 1e-3) / 1e9
            out["distributed_transpose_16384_f32"] = {
                "value": gbs, "unit": "GB/s", "ms_per_step": ms, "gpu_launches": launches, "scaling": "strong", "slab_checked": ok, "peer_block_checked": peer_ok,
                "nvlink_egress_GBps_per_gpu": 4.0 * n * rb * (world - 1) / world / (ms * 1e-3) / 1e9,
                "nvlink_expected_tx_bytes_per_step_per_gpu": 4.0 * n * rb * (world - 1) / world,
                "nvlink_counters_per_step_rank0": nv_meas,   # NVML hardware counters (payload bytes, all links) around the same steps
                "collective": "all-to-all fused into the transpose kernel's store phase (b200_transpose_alltoall: peer stores over NVLink + flag barrier), no NCCL",
                "roofline": {"bound": "nvlink", "achieved": 4.0 * n * rb * (world - 1) / world / (ms * 1e-3) / 1e9, "peak": 900.0, "unit": "GB/s",
                             "frac": 4.0 * n * rb * (world - 1) / world / (ms * 1e-3) / 1e9 / 900.0, "peak_source": "NVLink 5 nominal, per direction"}}
            barrier()
            a_sym.free()
            del b_loc
        except Exception as ex:
            out["distributed_transpose_16384_f32"] = {"error": str(ex)[:300]}

    # ---- histogram 2^30 Int32 -> 256 bins, input slab per rank, one all-reduce of the bins (strong scaling) ----
    try:
        n = N_HIST
        s0, s1 = mg.histogram_shard(n, world, rank)
        inp = orc.gen_bins_i32(s1 - s0, 1236 + rank, 256)
        d_in = EX.move(be, inp)
        KA.synchronize(be)
        bins = KA.zeros(be, np.int32, 256)
        comm, peers = None, None
        if world > 1:
            import torch
            uid = torch.zeros(_lib.B200_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid = torch.frombuffer(bytearray(mg.Comm.unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(uid, 0)
            comm = mg.Comm(world, rank, bytes(uid.cpu().numpy().tobytes()))
            peers = mg.PeerGroup(world, rank, mg.torch_gather(dist))

        def step_hist():
            if peers is not None:
                peers.histogram_allreduce(bins, d_in)      # one kernel: counting + all-reduce over NVLink peer memory
            else:
                EX.histogram_(bins, d_in)

        def step_hist_nccl():
            EX.histogram_(bins, d_in)
            comm.allreduce_sum_i32(bins.ptr, 256)

        bins.fill_(0)
        step_hist()
        KA.synchronize(be)
        got = KA.Array(bins)
        check = bool(int(got.astype(np.int64).sum()) == n)  # every input lies in 1..256: counts must add up to n
        nvl = NvlinkCounter(env_int("LOCAL_RANK", 0)) if world > 1 else None
        nv0 = nvl.start() if nvl else None
        ms, launches = timed(step_hist, steps_short, warmup)
        nv_meas = nvl.delta_per_step(nv0, steps_short + warmup) if nvl else None
        gbs = 4.0 * n / (ms * 1e-3) / 1e9  # whole job: all ranks' inputs / max time
        out["histogram_2^30_i32_256bins"] = {"value": gbs, "unit": "GB/s", "ms_per_step": ms, "gpu_launches": launches,
                                             "sum_conserved": check, "scaling": "strong", "steps": steps_short,
                                             "nvlink_counters_per_step_rank0": nv_meas,
                                             "nvlink_expected_tx_bytes_per_step_per_gpu": (256 * 8 * (world - 1)) if world > 1 else None,
                                             "roofline": {"bound": "hbm", "achieved": gbs / world, "peak": peaks["hbm"],
                                                          "frac": gbs / world / peaks["hbm"],
                                                          "traffic": traffic_from_profiles("hist_atomic_kernel", 1.0 / world),
                                                          "algorithmic_bytes_per_launch": 4 * (s1 - s0) + 1024,
                                                          "note": "read-only stream: the measured peak is a copy (mixed traffic); read-only ceiling "
                                                                  "7.2-7.5 TB/s (profiles/r04_hbm_stream.txt), so frac can exceed 1"},
                                             "collective": ("fused in the counting kernel: {count,tag} slot stores over NVLink peer memory "
                                                            "(b200_histogram_i32_allreduce), no NCCL launch") if world > 1 else None}
        if comm is not None:
            bins.fill_(0)
            step_hist_nccl()
            KA.synchronize(be)
            check2 = bool(np.array_equal(KA.Array(bins), got))
            ms2, _ = timed(step_hist_nccl, steps_short, warmup)
            out["histogram_2^30_i32_256bins"]["nccl_route"] = {"value": 4.0 * n / (ms2 * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms2,
                                                                "same_bits_as_fused": check2,
                                                                "collective": "b200_histogram_i32 + NCCL all-reduce of 256 x Int32 (2 launches)"}
        if peers is not None:
            peers.free()
        if comm is not None:
            comm.destroy()
        if world == 1:
            try:   # e2e: Int32 input in pinned HOST memory, bins back on the host; h2d | counting overlapped by slabs (b200_histogram_i32_host)
                hp = C.c_void_p()
                _lib.call("b200_host_alloc", C.byref(hp), 4 * n)
                h_in = np.ctypeslib.as_array((C.c_int32 * n).from_address(hp.value))
                h_in[...] = inp
                h_bins = np.zeros(256, dtype=np.int32)
                EX.histogram_host_(be, h_bins, h_in)
                KA.synchronize(be)
                ok_e = bool(np.array_equal(h_bins, got))
                t0 = time.perf_counter()
                for _ in range(3):
                    h_bins[:] = 0
                    EX.histogram_host_(be, h_bins, h_in)
                    KA.synchronize(be)
                dt = (time.perf_counter() - t0) / 3
                out["histogram_2^30_i32_256bins"]["e2e"] = {"value": 4.0 * n / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * n,
                                                            "d2h_bytes_per_step": 1024, "ms_per_step": dt * 1e3, "steps": 3, "result_checked": ok_e,
                                                            "path": "EX.histogram_host_ -> b200_histogram_i32_host: pinned host input, slab-wise h2d | counting"}
                _lib.call("b200_host_free", hp)
            except Exception as ex:
                out["histogram_2^30_i32_256bins"]["e2e"] = {"value": None, "error": str(ex)[:200]}
        del d_in, bins
        if cpu:
            hbins = np.zeros(256, dtype=np.int32)
            out["histogram_2^30_i32_256bins"]["cpu_baseline"] = cpu_time(
                lambda: orc.histogram(hbins, inp, 256), 4.0 * n, "GB/s", 1e9,
                "full 2^30 Int32 histogram_kernel!, workgroup 256 (per-group local histogram + merge), OpenMP over work-groups")
    except Exception as ex:
        out["histogram_2^30_i32_256bins"] = {"error": str(ex)[:300]}

    # ---- FP32 SIMT matmul 8192^3 (reference summation order), A/C row panels per rank, B replicated (strong) ----
    try:
        n = N_MM
        m0, m1 = mg.matmul_shard(n, world, rank)
        A = orc.gen_uniform_f32((n, n), 1237)
        Ap = KA.to_device(np.asfortranarray(A[m0:m1, :]))
        del A
        dBm = KA.to_device(orc.gen_uniform_f32((n, n), 1238))
        dC = KA.zeros(be, np.float32, m1 - m0, n)

        def step_mm():
            EX.performant_matmul_(dC, Ap, dBm)

        mm_steps = max(2, min(steps, 5))
        ms, launches = timed(step_mm, mm_steps, 3)
        tf = 2.0 * n * n * n / (ms * 1e-3) / 1e12
        fma_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
        out["matmul_f32_8192"] = {"value": tf, "unit": "TFLOP/s", "ms_per_step": ms, "gpu_launches": launches, "scaling": "strong",
                                  "roofline": {"bound": "fma", "achieved": tf / world, "peak": fma_peak, "frac": tf / world / fma_peak,
                                               "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz"}}
        if world == 1:
            try:   # e2e: A, B in pinned HOST memory, C back to the host: copyto! in, performant_matmul!, copyto! out, synchronize
                hp = [C.c_void_p() for _ in range(3)]
                for q in hp:
                    _lib.call("b200_host_alloc", C.byref(q), 4 * n * n)
                hA_, hB_, hC_e = (np.ctypeslib.as_array((C.c_float * (n * n)).from_address(q.value)).reshape((n, n), order="F") for q in hp)
                hA_[...] = KA.Array(Ap)
                hB_[...] = KA.Array(dBm)

                def step_mm_seq():
                    KA.copyto_(be, Ap, hA_)
                    KA.copyto_(be, dBm, hB_)
                    EX.performant_matmul_(dC, Ap, dBm)
                    KA.copyto_(be, hC_e, dC)
                    KA.synchronize(be)

                def step_mm_e2e():   # the public host-array call: B uploaded once, A / C row panels streamed around the GEMMs
                    EX.performant_matmul_host_(be, hC_e, hA_, hB_)
                    KA.synchronize(be)

                step_mm_seq()
                t0 = time.perf_counter()
                step_mm_seq()
                dt_seq = time.perf_counter() - t0
                step_mm_e2e()
                t0 = time.perf_counter()
                for _ in range(2):
                    step_mm_e2e()
                dt = (time.perf_counter() - t0) / 2
                truth_row = hA_[7, :].astype(np.float64) @ hB_[:, :64].astype(np.float64)
                out["matmul_f32_8192"]["e2e"] = {"value": 2.0 * n ** 3 / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": 8 * n * n,
                                                 "d2h_bytes_per_step": 4 * n * n, "ms_per_step": dt * 1e3, "steps": 2,
                                                 "result_checked": bool(np.allclose(hC_e[7, :64], truth_row, rtol=1e-5)),
                                                 "path": "EX.performant_matmul_host_ -> b200_matmul_f32_host: pinned host A, B, C; B h2d once, A panels h2d | GEMM | C panels d2h overlapped",
                                                 "unpipelined_value": 2.0 * n ** 3 / dt_seq / 1e12}
                for q in hp:
                    _lib.call("b200_host_free", q)
            except Exception as ex:
                out["matmul_f32_8192"]["e2e"] = {"value": None, "error": str(ex)[:200]}
        if cpu:
            ms_ = 2048
            hA_, hB_ = orc.gen_uniform_f32((ms_, ms_), 1237), orc.gen_uniform_f32((ms_, ms_), 1238)
            hC_ = np.zeros((ms_, ms_), dtype=np.float32, order="F")
            out["matmul_f32_8192"]["cpu_baseline"] = cpu_time(
                lambda: orc.coalesced_matmul(hC_, hA_, hB_), 2.0 * ms_ ** 3, "TFLOP/s", 1e12,
                "2048^3 sample of coalesced_matmul_kernel! (16x16 local-memory tiles, reference summation order), OpenMP over work-groups")
        # ---- opt-in tensor-core paths, same A/C row-panel sharding (B replicated); reported only when a sampled parity
        # ---- check of THIS run's result passes (64 entries of C against float64 dot products, rtol 1e-2) ----
        hA, hB = KA.Array(Ap), KA.Array(dBm)
        rs = np.random.default_rng(7)
        ii, jj = rs.integers(0, m1 - m0, 64), rs.integers(0, n, 64)
        truth = np.array([np.dot(hA[i, :].astype(np.float64), hB[:, j].astype(np.float64)) for i, j in zip(ii, jj)])

        def parity_ok(rtol=1e-2):
            got = KA.Array(dC)[ii, jj].astype(np.float64)
            return bool(np.allclose(got, truth, rtol=rtol, atol=0.0)), float(np.max(np.abs(got - truth) / np.abs(truth)))

        mp = m1 - m0
        ws = KA.allocate(be, np.uint16, mp * n + n * n)
        _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr), C.c_void_p(Ap.ptr), mp, n, mp, 1)
        _lib.call("b200_convert_f32_bf16", C.c_void_p(ws.ptr + 2 * mp * n), C.c_void_p(dBm.ptr), n, n, n, 0)

        def step_tc():
            _lib.call("b200_matmul_bf16", C.c_void_p(dC.ptr), C.c_void_p(ws.ptr), C.c_void_p(ws.ptr + 2 * mp * n), mp, n, n, mp)

        def step_tc_both():  # both FP32 operands converted on every call
            EX.matmul_bf16_(dC, Ap, dBm, ws)

        Bbf_res = EX.prepare_bf16_b(dBm)     # the sharded policy: B is replicated AND converted once at set-up

        def step_tc_e2e():   # per call: the rank's FP32 A row panel is converted, then the GEMM
            EX.matmul_bf16_resident_b_(dC, Ap, Bbf_res, n, ws)

        def step_tf32():
            EX.matmul_tf32_(dC, Ap, dBm)

        ws3 = KA.allocate(be, np.float32, mp * n + n * n)

        def step_tf32x3():
            EX.matmul_tf32x3_(dC, Ap, dBm, ws3)

        tkey = {"matmul_bf16_tcgen05_8192": "gemm_tcgen05_kernel<bf16>", "matmul_tf32_tcgen05_8192": "gemm_tcgen05_kernel<tf32>",
                "matmul_f32_tf32x3_tcgen05_8192": "gemm_tcgen05_kernel<tf32x3>"}
        for key, fn, peak, note, rtol in (
                ("matmul_bf16_tcgen05_8192", step_tc, peaks["bf16"], "BF16 operands pre-converted (conversion pass excluded), FP32 accumulate/output", 1e-2),
                ("matmul_bf16_tcgen05_8192_from_f32", step_tc_e2e, peaks["bf16"],
                 "FP32 operands in HBM; conversion policy: B (replicated) converted to BF16 once at set-up and kept resident, the rank's A row "
                 "panel converted on every step, then the GEMM", 1e-2),
                ("matmul_bf16_tcgen05_8192_from_f32_both_converted", step_tc_both, peaks["bf16"],
                 "FP32 operands in HBM; A panel AND the whole B converted on every step, then the GEMM", 1e-2),
                ("matmul_tf32_tcgen05_8192", step_tf32, peaks["bf16"] / 2, "FP32 operands read directly by TMA (no conversion pass); peak = half the measured BF16 peak", 1e-2),
                ("matmul_f32_tf32x3_tcgen05_8192", step_tf32x3, peaks["bf16"] / 6,
                 "FP32-grade product on the tensor cores: operands split into two TF32 terms, 3 MMAs per k step, K accumulated in chunks of 256 "
                 "(split pass included); sampled parity at rtol 1e-5; peak = measured BF16 peak / 6", 1e-5)):
            dC.fill_(0)
            fn()
            KA.synchronize(be)
            ok, err = parity_ok(rtol)
            if not ok:
                out[key] = {"error": "sampled parity check failed", "max_rel_err": err}
                continue
            ms, launches = timed(fn, 10, 3)
            tf = 2.0 * n * n * n / (ms * 1e-3) / 1e12
            out[key] = {"value": tf, "unit": "TFLOP/s", "ms_per_step": ms, "gpu_launches": launches, "scaling": "strong", "note": note,
                        "max_rel_err_sampled_vs_f64": err,
                        "roofline": {"bound": "tensor", "achieved": tf / world, "peak": peak, "frac": tf / world / peak,
                                     "traffic": traffic_from_profiles(tkey[key], 1.0 / world) if key in tkey else None,
                                     "peak_note": ("measured cuBLAS BF16 burst peak" if peak == peaks["bf16"] else
                                                   "builder-defined: measured BF16 peak / 2 (TF32 runs at half the BF16 rate)" if peak == peaks["bf16"] / 2
                                                   else "builder-defined: measured BF16 peak / 6 (three TF32 MMAs per k step at half the BF16 rate)")}}
    except Exception as ex:
        out["matmul_f32_8192"] = out.get("matmul_f32_8192", {"error": str(ex)[:300]})
    return out


if __name__ == "__main__":
    sys.exit(main())
