#!/usr/bin/env python
"""Freeze the host-side API surface the reference's own callers use on a backend.

Scans the testsuite files and examples on the hot path (test/{test,intrinsics,localmem,private,unroll,copyto,devices}.jl
and examples/*.jl) for every `KernelAbstractions.<name>` / `KA.<name>` / `KI.<name>` reference and for calls of the names KA exports,
and writes them to tests/golden/api_calls.json with the file:line of the first use.  /root/reference does not travel to the GPU box,
so tests/test_julia_binding.py reads the frozen list, not the reference.

    python tests/golden/make_api_calls.py /root/reference
"""
import json
import os
import re
import sys

FILES = ["test/test.jl", "test/intrinsics.jl", "test/localmem.jl", "test/private.jl", "test/unroll.jl", "test/copyto.jl", "test/devices.jl",
         "examples/memcopy.jl", "examples/memcopy_static.jl", "examples/naive_transpose.jl", "examples/histogram.jl", "examples/matmul.jl",
         "examples/performant_matmul.jl", "examples/mpi.jl", "examples/numa_aware.jl", "examples/performance.jl"]
QUAL = re.compile(r"\b(KernelAbstractions|KA|KI)\.(@?[A-Za-z_][A-Za-z0-9_!]*)")
# names exported by KernelAbstractions (src/KernelAbstractions.jl:3-10) that are host-side functions a backend must make work
EXPORTED = ["allocate", "get_backend", "synchronize", "copyto!", "zeros", "ones", "adapt", "pagelock!", "unsafe_free!"]
BARE = re.compile(r"(?<![\w.!])(" + "|".join(re.escape(n) for n in EXPORTED) + r")\(")


def main(ref):
    found = {}
    for rel in FILES:
        path = os.path.join(ref, rel)
        for ln, line in enumerate(open(path, encoding="utf-8"), 1):
            code = line.split("#", 1)[0]
            for m in QUAL.finditer(code):
                mod = "KI" if m.group(1) == "KI" else "KA"
                found.setdefault(f"{mod}.{m.group(2)}", f"{rel}:{ln}")
            for m in BARE.finditer(code):
                found.setdefault(f"KA.{m.group(1)}", f"{rel}:{ln}")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_calls.json")
    json.dump({"source": "reference files scanned: " + ", ".join(FILES), "calls": dict(sorted(found.items()))}, open(out, "w"), indent=1)
    print(len(found), "names ->", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
