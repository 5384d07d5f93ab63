"""GPU: the multi-GPU entry points, driver visible.

* world = 1 (any box): the exchange! pattern of examples/mpi.jl:24-39 with ONE rank (the example itself runs with one rank:
  dst_rank == src_rank == 0), the fused histogram + all-reduce kernel and the fused transpose + all-to-all kernel with the rank
  as its own only peer -- the same kernels, slot / flag protocol included, that run on 2-8 GPUs;
* world >= 2 (skipped unless the box has at least two GPUs): tests/gpu_multi_check.py under torch.distributed.run, one rank per
  GPU -- sharded histogram (NCCL and fused routes) against the oracle, ring exchange, row-panel matmul, slab copy / transpose,
  distributed transpose against its definition and the NCCL route.
"""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import b200ka as KA
from b200ka import _lib, examples as EX, multigpu as mg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _self_gather(blob):
    return [blob]


def test_exchange_single_rank():
    """examples/mpi.jl:41-75 with MPI.Comm_size == 1: the payload comes back from the rank itself, stream ordered."""
    be = KA.B200Backend()
    comm = mg.Comm(1, 0, mg.Comm.unique_id())
    rank, world = 0, 1
    dst_rank, src_rank = (rank + 1) % world, (rank - 1) % world
    T, M = np.int64, 10
    d_recv_buf = KA.allocate(be, T, M).fill_(-1)
    d_send_buf = KA.to_device(np.full(M, rank + 41, dtype=T))
    KA.priority_(be, "high")                       # examples/mpi.jl:27
    mg.exchange_(comm, be, d_send_buf, d_recv_buf, src_rank, dst_rank)
    KA.synchronize(be)                             # "guaranteed to be cooperative" (:30)
    KA.priority_(be, "normal")
    assert np.all(KA.Array(d_recv_buf) == src_rank + 41)
    # all-reduce and broadcast of a one-rank communicator are identities
    bins = KA.to_device(np.arange(256, dtype=np.int32))
    comm.allreduce_sum_i32(bins.ptr, 256)
    comm.broadcast(bins.ptr, bins.nbytes, 0)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(bins), np.arange(256, dtype=np.int32))
    comm.destroy()


@pytest.mark.parametrize("ll,dyn", [(0, 16), (0, 0), (0, 1), (0, 127), (1, 16)])
@pytest.mark.parametrize("n,nbins,offset", [(100003, 256, 0), (17, 256, 0), (0, 256, 0), (5 * (1 << 20) + 3, 1000, 0), (999999, 64, 1),
                                            (1 << 22, 256, 3), (2, 256, 2), (3 * (1 << 20), 1, 0), (3 * (1 << 20), 1024, 0), (1 << 21, 1025, 0),
                                            (1 << 21, 5000, 1), (1 << 22, 7, 0), ((1 << 25) + 12345, 256, 1), (148 * 4096 * 20 + 5, 256, 0)])
def test_fused_histogram_allreduce_single_rank(orc, n, nbins, offset, ll, dyn):
    """b200_histogram_i32_allreduce with the rank as its only peer == the oracle; accumulates like the reference's +=; several
    steps on one PeerGroup (slot parity / tags); unaligned inputs (offset elements off a 16-byte boundary); both intra-GPU merges
    (hist.fused_ll = 0, the default: ticket + partial; 1: per-CTA {count, tag} rows polled by CTA 0, for nbins <= 1024)."""
    be = KA.B200Backend()
    _lib.call("b200_tune_set", b"hist.fused_ll", ll)
    _lib.call("b200_tune_set", b"hist.dyn_rounds", dyn)     # rounds of the work-stealing tail (0: every unit dealt out statically)
    peers = mg.PeerGroup(1, 0, _self_gather)
    try:
        for step in range(3):
            whole = orc.gen_bins_i32(max(n + offset, 1), 4000 + step, nbins)[:n + offset]
            d_all = EX.move(be, whole) if whole.size else KA.allocate(be, np.int32, 0)
            d_in = d_all.view(offset, (n,)) if n else KA.allocate(be, np.int32, 0)
            want = np.zeros(nbins, dtype=np.int32)
            if n:
                orc.histogram(want, np.ascontiguousarray(whole[offset:]), 256)
            bins = KA.to_device(np.arange(nbins, dtype=np.int32))
            peers.histogram_allreduce(bins, d_in)
            KA.synchronize(be)
            assert np.array_equal(KA.Array(bins), want + np.arange(nbins, dtype=np.int32)), f"step {step}"
        t = peers.timeline()
        assert t["all_summed"] >= t["slots_sent"] >= t["counts_merged"]
    finally:
        _lib.call("b200_tune_set", b"hist.fused_ll", 0)
        _lib.call("b200_tune_set", b"hist.dyn_rounds", 0)
        peers.free()


@pytest.mark.parametrize("N0,N1,dt", [(64, 128, np.float32), (256, 64, np.float64), (1024, 2048, np.float32)])
def test_transpose_alltoall_single_rank(N0, N1, dt):
    """b200_transpose_alltoall on one rank == the definition; three calls back to back with a consumer of the result in between
    and NO barrier from the caller (the write-after-read guard of the kernel)."""
    be = KA.B200Backend()
    rng = np.random.default_rng(N0 + N1)
    es = np.dtype(dt).itemsize
    a_sym = mg.SymBuffer(N1 * N0 * es, 1, 0, _self_gather)
    try:
        a_view = a_sym.array(dt, (N1, N0))
        keep = KA.allocate(be, dt, N1, N0)
        for tma, fused in ((0, 0), (1, 0), (0, 1), (0, 0)):
            _lib.call("b200_tune_set", b"a2a.tma_store", tma)
            _lib.call("b200_tune_set", b"a2a.fused_signal", fused)     # completion inside the data kernel (opt-in A/B)
            whole = np.asfortranarray((rng.standard_normal((N0, N1)) * 100).astype(dt))
            b_loc = KA.to_device(whole)
            mg.transpose_alltoall_(a_sym, b_loc, N0, N1)
            KA.copyto_(be, keep, a_view)           # a consumer of the slab, stream ordered behind the call
            KA.synchronize(be)
            assert np.array_equal(KA.Array(keep), whole.T)
    finally:
        _lib.call("b200_tune_set", b"a2a.tma_store", 0)
        _lib.call("b200_tune_set", b"a2a.fused_signal", 0)
        a_sym.free()


def test_missing_peer_times_out_instead_of_hanging():
    """Rank 0 of a 2-rank job whose peer never makes the matching call: every wait inside the kernels is bounded
    (`a2a.timeout_ms`, `hist.fused_timeout_ms`), the call returns, and the NEXT call on the handle reports the lost step."""
    import ctypes as C
    import time

    be = KA.B200Backend()
    # two symmetric buffers on the one GPU, connected as ranks 0 and 1 of a single-process job; only rank 0 ever calls
    N0 = N1 = 256
    hs = []
    for _ in range(2):
        h = C.c_void_p()
        _lib.call("b200_sym_alloc", C.byref(h), C.c_size_t(N1 * (N0 // 2) * 4))
        hs.append(h)
    ptrs = (C.c_void_p * 2)()
    for r in range(2):
        p = C.c_void_p()
        _lib.call("b200_sym_ptr", hs[r], -1, C.byref(p))
        ptrs[r] = p.value
    for r in range(2):
        _lib.call("b200_sym_connect_ptrs", hs[r], 2, r, ptrs)
    b_loc = KA.rand_uniform_(KA.allocate(be, np.float32, N0, N1 // 2), 5)
    _lib.call("b200_tune_set", b"a2a.timeout_ms", 30)
    try:
        t0 = time.perf_counter()
        _lib.call("b200_transpose_alltoall", hs[0], C.c_void_p(b_loc.ptr), N0, N1, 4)
        KA.synchronize(be)
        assert time.perf_counter() - t0 < 5.0
        with pytest.raises(_lib.B200Error, match="timed out waiting for a peer"):
            _lib.call("b200_transpose_alltoall", hs[0], C.c_void_p(b_loc.ptr), N0, N1, 4)
    finally:
        _lib.call("b200_tune_set", b"a2a.timeout_ms", 10000)
        for h in hs:
            _lib.call("b200_sym_free", h)
    # the fused histogram: rank 0 of 2, the peer's slots never arrive
    ps = []
    for _ in range(2):
        h = C.c_void_p()
        _lib.call("b200_p2p_alloc", C.byref(h))
        ps.append(h)
    harr = (C.c_void_p * 2)(*[h.value for h in ps])
    for r in range(2):
        _lib.call("b200_p2p_connect_ptrs", ps[r], 2, r, harr)
    d_in = KA.rand_bins_(KA.allocate(be, np.int32, 1 << 16), 1, 256)
    bins = KA.zeros(be, np.int32, 256)
    _lib.call("b200_tune_set", b"hist.fused_timeout_ms", 30)
    try:
        t0 = time.perf_counter()
        _lib.call("b200_histogram_i32_allreduce", ps[0], C.c_void_p(bins.ptr), 256, C.c_void_p(d_in.ptr), d_in.size)
        KA.synchronize(be)
        assert time.perf_counter() - t0 < 5.0
        with pytest.raises(_lib.B200Error, match="timed out waiting for a peer"):
            _lib.call("b200_histogram_i32_allreduce", ps[0], C.c_void_p(bins.ptr), 256, C.c_void_p(d_in.ptr), d_in.size)
    finally:
        _lib.call("b200_tune_set", b"hist.fused_timeout_ms", 10000)
        for h in ps:
            _lib.call("b200_p2p_free", h)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_multi_gpu_parity_all_ranks():
    """tests/gpu_multi_check.py, one rank per GPU (2..8): skipped on a one-GPU box."""
    ngpu = KA.ndevices(KA.B200Backend())
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gpu_multi_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world
    for key in ["histogram_routes_match_oracle", "ring_exchange", "matmul_row_panels_bit_exact", "slab_copy_and_transpose_bit_exact",
                "distributed_transpose_fused_eq_nccl_eq_definition", "distributed_transpose_back_to_back"]:
        assert out[key] is True, key
