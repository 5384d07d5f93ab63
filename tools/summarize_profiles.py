#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (named per round).

usage: summarize_profiles.py r01
  gpurun_out/prof_<kernel>.ncu-rep  ->  profiles/<round>_<kernel>_ncu.csv   (selected raw metrics, one capture)
  gpurun_out/launches.csv           ->  profiles/<round>_launches.csv + per-kernel share table
  profiles/traffic.json             <-  dram read+write bytes per launch (bench.py's roofline.traffic)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

traffic_path = os.path.join(OUT, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
summary_lines = []
for f in sorted(os.listdir(GO)):
    if not (f.startswith("prof_") and f.endswith(".ncu-rep")):
        continue
    name = f[len("prof_"):-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", os.path.join(GO, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    kname = d.get("Kernel Name", ("?", ""))[0]
    with open(os.path.join(OUT, f"{rnd}_{name}_ncu.csv"), "w", newline="") as fo:
        w = csv.writer(fo)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", kname, ""])
        w.writerow(["command", "ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1 python tools/run_one.py ...", ""])
        for k in KEEP:
            if k in d:
                w.writerow([k, d[k][0], d[k][1]])
    rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
    if rd and wr:
        tot = float(rd[0].replace(",", "")) * UNIT_SCALE.get(rd[1], 1.0) + float(wr[0].replace(",", "")) * UNIT_SCALE.get(wr[1], 1.0)
        short = kname.split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        # one kernel template serves three tensor-core products: key those by variant, never by the bare template name
        variant = {"tcgemm": "<bf16>", "tf32gemm": "<tf32>", "tf32x3gemm": "<tf32x3>"}.get(name)
        traffic[short + variant if variant else short] = tot
        traffic[f"capture:{name}"] = tot
        dur = d.get("gpu__time_duration.sum", ("?", ""))
        summary_lines.append(f"| {name} | `{short}` | {dur[0]} {dur[1]} | {tot / 1e9:.3f} GB | "
                             f"{d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', ('?',))[0]} | "
                             f"{d.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', ('?',))[0]} | "
                             f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', d.get('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed', ('?',)))[0]} |")
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)

lp = os.path.join(GO, "launches.csv")
share_lines = []
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    with open(os.path.join(OUT, f"{rnd}_launches.csv"), "w", newline="") as fo:
        w = csv.writer(fo)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration_us"])
        for r in rows[1:]:
            if r[ix["Metric Name"]] != "gpu__time_duration.sum":
                continue
            v = float(r[ix["Metric Value"]].replace(",", ""))
            u = r[ix["Metric Unit"]]
            v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
            k = r[ix["Kernel Name"]]
            w.writerow([r[ix["ID"]], k, r[ix["Grid Size"]], r[ix["Block Size"]], f"{v:.3f}"])
            a = agg.setdefault(k.split("(")[0], [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values())
    for k, (n, t) in agg.items():
        share_lines.append(f"| `{k[:80]}` | {n} | {t / n:.1f} | {t:.1f} | {100 * t / tot:.1f}% |")

with open(os.path.join(OUT, f"{rnd}_summary.md"), "w") as fo:
    fo.write(f"# ncu summaries, round {rnd}\n\n")
    fo.write("One `ncu --set full --clock-control none --import-source on` capture per hot kernel (second launch of "
             "`python tools/run_one.py <kernel> ...`, taken only after the identical plain command exited 0). "
             "Durations are cold-cache/serialised: compare shares, not absolutes. Full metric subsets: `" + rnd + "_<kernel>_ncu.csv`.\n\n")
    fo.write("| capture | kernel | duration | DRAM read+write per launch | DRAM throughput % | FMA pipe % | tensor pipe % |\n|---|---|---|---|---|---|---|\n")
    fo.write("\n".join(summary_lines) + "\n\n")
    if share_lines:
        fo.write("## Launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`\n\n"
                 "`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv` (raw: `" + rnd + "_launches.csv`).\n\n")
        fo.write("| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|\n" + "\n".join(share_lines) + "\n")
print(open(os.path.join(OUT, f"{rnd}_summary.md")).read())
