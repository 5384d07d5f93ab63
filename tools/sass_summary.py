#!/usr/bin/env python
"""Blackwell evidence from the shipped binary: per cubin of libb200ka.so, the count of the SASS opcodes that prove tcgen05 / TMEM /
TMA / bulk-copy use (the mnemonics of /opt/skills/guides/B200_PROFILING.md), plus the kernels each cubin holds.

    python tools/sass_summary.py [round]     ->  profiles/<round>_sass_summary.txt   (run by __graft_entry__.build())

Needs only cuobjdump (no GPU): `cuobjdump -sass` of the in-tree .so, which is compiled for sm_100a alone.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "kernelabstractions.jl_b200", "libb200ka.so")
PATTERNS = [
    ("UTCHMMA (tcgen05.mma, 1 CTA)", r"\bUTCHMMA(?!\.2CTA)"),
    ("UTCHMMA.2CTA (tcgen05.mma cta_group::2)", r"\bUTCHMMA\.2CTA"),
    ("UTCQMMA/UTCOMMA/UTCIMMA (other tcgen05 kinds)", r"\bUTC[QOI]MMA"),
    ("LDTM (tcgen05.ld, TMEM -> registers)", r"\bLDTM"),
    ("STTM (tcgen05.st)", r"\bSTTM"),
    ("UTCBAR (tcgen05.commit -> mbarrier)", r"\bUTCBAR"),
    ("UTCATOMSWS (tcgen05.alloc/dealloc)", r"\bUTCATOMSWS"),
    ("UTMALDG (TMA tensor load)", r"\bUTMALDG"),
    ("UTMASTG (TMA tensor store)", r"\bUTMASTG"),
    ("UBLKCP (cp.async.bulk, 1-d)", r"\bUBLKCP"),
    ("SYNCS (mbarrier ops)", r"\bSYNCS"),
    ("FFMA2 (packed FP32 FMA)", r"\bFFMA2"),
    ("ATOMS/REDS (shared atomics)", r"\b(ATOMS|REDS)\b"),
    ("ACQBULK / PREEXIT (griddepcontrol: programmatic dependent launch)", r"\bACQBULK|\bPREEXIT"),
]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r05"
    if not os.path.exists(SO):
        raise SystemExit(f"{SO} not built")
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    per_fn = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_fn[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                per_fn[cur][name] += 1
    demangle = {}
    try:
        names = list(per_fn)
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except Exception:
        pass
    total = collections.Counter()
    for c in per_fn.values():
        total.update(c)
    lines = [f"# SASS evidence, {rnd}: cuobjdump -sass kernelabstractions.jl_b200/libb200ka.so",
             f"architectures in the fat binary: {', '.join(arch)}   kernels: {len(per_fn)}", "",
             "## totals over all kernels"]
    for name, _ in PATTERNS:
        lines.append(f"{total[name]:8d}  {name}")
    lines += ["", "## per kernel (only kernels with at least one of the opcodes above; FFMA2 / SYNCS alone are not listed)"]
    for fn, c in per_fn.items():
        key = {k: v for k, v in c.items() if not k.startswith(("FFMA2", "SYNCS", "ATOMS"))}
        if not key:
            continue
        short = re.sub(r"\s+", " ", demangle.get(fn, fn))
        short = short.replace("void b200::", "").split("(")[0][:110]
        lines.append(f"{short}")
        lines.append("      " + ", ".join(f"{k.split(' ')[0]}={v}" for k, v in c.items()))
    out_path = os.path.join(ROOT, "profiles", f"{rnd}_sass_summary.txt")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{out_path}: " + ", ".join(f"{k.split(' ')[0]}={total[k]}" for k, _ in PATTERNS if total[k]))


if __name__ == "__main__":
    main()
