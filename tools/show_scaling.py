#!/usr/bin/env python
"""Strong-scaling table from the bench lines of one measurement pass: python tools/show_scaling.py [dir] [prefix]"""
import json
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "profiles"
pre = sys.argv[2] if len(sys.argv) > 2 else "r05"
lines = {}
for n in (1, 2, 4, 8):
    p = os.path.join(d, f"{pre}_bench_n{n}.json")
    if os.path.exists(p):
        js = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if js:
            lines[n] = json.loads(js[-1])
if 1 not in lines:
    raise SystemExit("no N=1 line")
rows = [("naive_transpose 16384^2 (headline)", {n: (l["value"], l["ms_per_step"]) for n, l in lines.items()})]
for key in lines[1]["suite"]:
    r = {}
    for n, l in lines.items():
        v = l["suite"].get(key)
        if isinstance(v, dict) and "value" in v:
            r[n] = (v["value"], v["ms_per_step"])
    rows.append((key + (" [weak]" if lines[1]["suite"][key].get("scaling") == "weak" else ""), r))
print("| kernel | " + " | ".join(f"N={n}: value, ms/step, x N=1" for n in sorted(lines)) + " |")
print("|---|" + "---|" * len(lines))
for name, r in rows:
    if 1 not in r:
        continue
    cells = []
    for n in sorted(lines):
        if n in r:
            cells.append(f"{r[n][0]:.1f}, {r[n][1]:.4f}, {r[n][0] / r[1][0]:.2f}x")
        else:
            cells.append("-")
    print(f"| {name} | " + " | ".join(cells) + " |")
for n, l in sorted(lines.items()):
    e = l.get("e2e") or {}
    print(f"e2e N={n}: {e.get('value')} GB/s ({e.get('route')}); pipelined {e.get('pipelined_value')}, in-order {e.get('unpipelined_value')}, copy-only ceiling {e.get('copy_only_duplex_value')}")
    dt = l["suite"].get("distributed_transpose_16384_f32")
    if dt:
        print(f"   distributed transpose N={n}: {dt.get('value')} GB/s, {dt.get('ms_per_step')} ms, egress {dt.get('nvlink_egress_GBps_per_gpu')} GB/s/GPU, slab {dt.get('slab_checked')} peer {dt.get('peer_block_checked')}")
    h = l["suite"].get("histogram_2^30_i32_256bins", {})
    if "nccl_route" in h:
        print(f"   histogram NCCL route N={n}: {h['nccl_route']['ms_per_step']} ms, same bits {h['nccl_route']['same_bits_as_fused']}")
