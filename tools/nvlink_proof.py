#!/usr/bin/env python
"""NVLink byte counters of the fused compute + collective kernels, measured by ncu.

ncu serialises and replays kernels, so two ranks that wait for each other inside their kernels would deadlock under it.  This tool
runs ONE rank of a 2-rank job, in ONE process that owns both GPUs (b200_sym_connect_ptrs / b200_p2p_connect_ptrs), with the
waiting switched off for that rank (`a2a.debug_no_sync`; the peer's histogram slots pre-filled with b200_p2p_debug_prefill): the
rank's kernel is exactly the production kernel, and what it stores into the peer's buffer crosses NVLink where ncu counts it.

    python tools/nvlink_proof.py                      # plain run: checks the peer's memory, prints the expected byte counts
    ncu --metrics nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,nvltx__bytes.sum,nvlrx__bytes.sum --clock-control none \
        -k regex:'transpose_a2a_kernel|hist_allreduce_kernel' --csv --log-file gpurun_out/nvlink_ncu.csv python tools/nvlink_proof.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
import b200ka as KA  # noqa: E402
from b200ka import _lib  # noqa: E402

be = KA.B200Backend()
if KA.ndevices(be) < 2:
    raise SystemExit("needs two GPUs")
N = int(os.environ.get("NVLINK_PROOF_N", 8192))
world = 2


def on(dev):
    _lib.call("b200_set_device", dev)


# ---------------- distributed transpose: rank 0 of 2, N x N Float32 ----------------
rb = N // world
syms = []
for r in range(world):
    on(r)
    h = C.c_void_p()
    _lib.call("b200_sym_alloc", C.byref(h), C.c_size_t(N * rb * 4))
    syms.append(h)
ptrs = (C.c_void_p * world)()
for r in range(world):
    on(r)
    p = C.c_void_p()
    _lib.call("b200_sym_ptr", syms[r], -1, C.byref(p))
    ptrs[r] = p.value
for r in range(world):
    on(r)
    _lib.call("b200_sym_connect_ptrs", syms[r], world, r, ptrs)
on(1)
_lib.call("b200_memset_async", C.c_void_p(ptrs[1]), 0, C.c_size_t(N * rb * 4))
_lib.call("b200_synchronize")
on(0)
b_loc = KA.rand_uniform_(KA.allocate(be, np.float32, N, rb), 1235)        # rank 0's column block of the global b
_lib.call("b200_tune_set", b"a2a.debug_no_sync", 1)
for _ in range(2):
    _lib.call("b200_transpose_alltoall", syms[0], C.c_void_p(b_loc.ptr), N, N, 4)
_lib.call("b200_synchronize")
_lib.call("b200_tune_set", b"a2a.debug_no_sync", 0)
hb = KA.Array(b_loc)
on(1)
peer = np.empty((N, rb), dtype=np.float32, order="F")
_lib.call("b200_memcpy_d2h_async", peer.ctypes.data_as(C.c_void_p), C.c_void_p(ptrs[1]), C.c_size_t(peer.nbytes))
_lib.call("b200_synchronize")
# rank 0 delivers transpose(b_loc[rows of rank 1, :]) into rows [0, rb) of rank 1's slab
ok_t = bool(np.array_equal(peer[:rb, :], hb[rb:2 * rb, :].T))
print(f"transpose_a2a_kernel, rank 0 of 2, {N}x{N} Float32: peer block landed in GPU 1's memory: {ok_t}; "
      f"expected NVLink user bytes per launch, GPU 0 tx: {4 * rb * rb} ({4 * rb * rb / 2**20:.0f} MiB = half of the rank's {4 * N * rb / 2**20:.0f} MiB slab)")

# ---------------- fused histogram + all-reduce: rank 0 of 2 ----------------
hs = []
for r in range(world):
    on(r)
    h = C.c_void_p()
    _lib.call("b200_p2p_alloc", C.byref(h))
    hs.append(h)
harr = (C.c_void_p * world)(*[h.value for h in hs])
for r in range(world):
    on(r)
    _lib.call("b200_p2p_connect_ptrs", hs[r], world, r, harr)
on(0)
n = 1 << 24
d_in = KA.rand_bins_(KA.allocate(be, np.int32, n), 1236, 256)
bins = KA.zeros(be, np.int32, 256)
fake = np.arange(256, dtype=np.uint32)                                      # what "rank 1" contributes
for step in range(2):
    _lib.call("b200_p2p_debug_prefill", hs[0], 1, fake.ctypes.data_as(C.POINTER(C.c_uint32)), 256)
    _lib.call("b200_histogram_i32_allreduce", hs[0], C.c_void_p(bins.ptr), 256, C.c_void_p(d_in.ptr), n)
    _lib.call("b200_synchronize")
got = KA.Array(bins).astype(np.int64)
want = 2 * (np.bincount(KA.Array(d_in), minlength=257)[1:].astype(np.int64) + fake)
ok_h = bool(np.array_equal(got, want))
print(f"hist_allreduce_kernel, rank 0 of 2, 2^24 Int32: bins == local counts + the pre-filled peer counts: {ok_h}; "
      f"expected NVLink user bytes per launch, GPU 0 tx: {256 * 8} (256 x {{count, tag}} words to the one peer)")
for h in hs:
    _lib.call("b200_p2p_free", h)
for h in syms:
    _lib.call("b200_sym_free", h)
if not (ok_t and ok_h):
    raise SystemExit("nvlink_proof: result check failed")
