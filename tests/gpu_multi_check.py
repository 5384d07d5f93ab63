#!/usr/bin/env python
"""Multi-GPU parity + timing check, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tests/gpu_multi_check.py [--time]

Checks (every rank asserts; rank 0 prints one JSON line):
  * sharded histogram, NCCL route  (b200_histogram_i32 + b200_allreduce_sum_i32)          == oracle of the whole input
  * sharded histogram, fused route (b200_histogram_i32_allreduce: one kernel, NVLink peer reductions) == the same bits,
    over several consecutive steps with different sizes (step parity / flag epochs), ragged and empty slabs
  * ring exchange (examples/mpi.jl:24-39) through b200_sendrecv
  * row-panel matmul: each rank's C panel is bit-identical to the same rows of the 1-GPU result
With --time: per-step device time (max over ranks) of both histogram routes at 2^30 elements total.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import b200ka as KA  # noqa: E402
import ka_oracle as orc  # noqa: E402
from b200ka import _lib, examples as EX, multigpu as mg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.call("b200_set_device", local)
    be = KA.B200Backend()
    for kv in filter(None, os.environ.get("B200_TUNE", "").split(",")):     # e.g. B200_TUNE=a2a.tma_store=0
        k, v = kv.rsplit("=", 1)
        _lib.call("b200_tune_set", k.encode(), int(v))
    out = {"world": world}

    uid = torch.zeros(_lib.B200_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(mg.Comm.unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    comm = mg.Comm(world, rank, bytes(uid.cpu().numpy().tobytes()))
    peers = mg.PeerGroup(world, rank, mg.torch_gather(dist))

    # ---- histogram: both routes against the oracle, several steps back to back ----
    sizes = [100003, 4 * 1000 * world, 17, 0, 5 * (1 << 20) + 4 * world, 999999, 3 * (1 << 20) + 1, 1 << 21]
    nbins_list = [256, 256, 256, 256, 1000, 64, 2000, 1024]
    for step, (n, nbins) in enumerate(zip(sizes, nbins_list)):
        s0, s1 = mg.histogram_shard(n, world, rank)
        whole = orc.gen_bins_i32(max(n, 1), 4000 + step, nbins)[:n]
        mine = np.ascontiguousarray(whole[s0:s1])
        d_in = EX.move(be, mine) if mine.size else KA.allocate(be, np.int32, 0)
        want = np.zeros(nbins, dtype=np.int32)
        if n:
            orc.histogram(want, whole, 256)
        # NCCL route
        bins = KA.zeros(be, np.int32, nbins)
        if mine.size:
            EX.histogram_(bins, d_in)
        comm.allreduce_sum_i32(bins.ptr, nbins)
        KA.synchronize(be)
        assert np.array_equal(KA.Array(bins), want), f"NCCL route mismatch at step {step} on rank {rank}"
        # fused route (accumulates into existing counts)
        bins2 = KA.to_device(np.arange(nbins, dtype=np.int32))
        peers.histogram_allreduce(bins2, d_in)
        KA.synchronize(be)
        assert np.array_equal(KA.Array(bins2), want + np.arange(nbins, dtype=np.int32)), f"fused route mismatch at step {step} on rank {rank}"
    out["histogram_routes_match_oracle"] = True

    # ---- ring exchange (examples/mpi.jl:41-75): M = 10 Int64, payload = sender's rank ----
    dst_rank, src_rank = (rank + 1) % world, (rank - 1) % world
    d_recv = KA.allocate(be, np.int64, 10).fill_(-1)
    d_send = KA.to_device(np.full(10, rank, dtype=np.int64))
    mg.exchange_(comm, be, d_send, d_recv, src_rank, dst_rank)
    KA.synchronize(be)
    assert np.all(KA.Array(d_recv) == src_rank)
    out["ring_exchange"] = True

    # ---- matmul row panels: bit-identical to the same rows of the full product ----
    n = 512
    A, B = orc.gen_uniform_f32((n, n), 1237), orc.gen_uniform_f32((n, n), 1238)
    m0, m1 = mg.matmul_shard(n, world, rank)
    dB = KA.allocate(be, np.float32, n, n)
    if rank == 0:
        KA.copyto_(be, dB, B)
        KA.synchronize(be)
    comm.broadcast(dB.ptr, dB.nbytes, 0)           # B replicated with one broadcast
    Cp = KA.zeros(be, np.float32, m1 - m0, n)
    EX.performant_matmul_(Cp, KA.to_device(np.asfortranarray(A[m0:m1, :])), dB)
    KA.synchronize(be)
    Cref = np.zeros((n, n), dtype=np.float32, order="F")
    orc.coalesced_matmul(Cref, A, B)
    assert np.array_equal(KA.Array(Cp), Cref[m0:m1, :])
    out["matmul_row_panels_bit_exact"] = True

    # ---- slab copy and slab transpose (contiguous element slabs; rows-of-b slabs, no collective) ----
    nc = 1 << 20
    whole_c = orc.gen_uniform_f32((nc,), 1234)
    c0, c1 = mg.copy_shard(nc, world, rank)
    d_dst = KA.zeros(be, np.float32, c1 - c0)
    EX.mycopy_(d_dst, KA.to_device(np.ascontiguousarray(whole_c[c0:c1])))
    nt = 512 * world
    whole_b = orc.gen_uniform_f32((nt, nt), 1235)
    j0, j1 = mg.transpose_shard(nt, world, rank)
    d_a = KA.zeros(be, np.float32, nt, j1 - j0)
    EX.naive_transpose_(d_a, KA.to_device(np.asfortranarray(whole_b[j0:j1, :])))     # (n/G) x n slab -> n x (n/G) slab
    KA.synchronize(be)
    # reassemble on every rank and compare with the single-device definition
    parts_c, parts_t = [None] * world, [None] * world
    dist.all_gather_object(parts_c, KA.Array(d_dst))
    dist.all_gather_object(parts_t, KA.Array(d_a))
    assert np.array_equal(np.concatenate(parts_c), whole_c)
    assert np.array_equal(np.concatenate(parts_t, axis=1), whole_b.T)
    out["slab_copy_and_transpose_bit_exact"] = True

    # ---- distributed transpose: fused tile transpose + all-to-all over peer memory == definition == NCCL route ----
    gather = mg.torch_gather(dist)
    for (N0, N1, dt) in [(64 * world, 128 * world, np.float32), (256 * world, 64 * world, np.float64), (1024, 2048, np.float32),
                         (64 * world, 256 * world, np.float64), (192 * world, 512 * world, np.float32)]:
        if N0 % (64 * world) or N1 % (64 * world):
            continue
        rb, cb = N0 // world, N1 // world
        rng = np.random.default_rng(N0 + N1)
        whole = np.asfortranarray((rng.standard_normal((N0, N1)) * 100).astype(dt))      # every rank builds the same global b
        b_loc = KA.to_device(np.asfortranarray(whole[:, rank * cb:(rank + 1) * cb]))
        es = np.dtype(dt).itemsize
        a_sym = mg.SymBuffer(N1 * rb * es, world, rank, gather)
        a_view = a_sym.array(dt, (N1, rb))
        a_view.fill_(0)
        KA.synchronize(be)
        dist.barrier()                                   # nobody stores into a slab that is still being zeroed
        want = np.asfortranarray(whole.T[:, rank * rb:(rank + 1) * rb])                   # a[:, c0(rank)]
        for tma in (0, 1, 0):                            # plain-store and TMA-store variants; repeated: the flag epochs must keep working
            _lib.call("b200_tune_set", b"a2a.tma_store", tma)
            a_view.fill_(0)
            KA.synchronize(be)
            dist.barrier()
            mg.transpose_alltoall_(a_sym, b_loc, N0, N1)
            KA.synchronize(be)
            assert np.array_equal(KA.Array(a_view), want), f"fused distributed transpose (tma_store={tma}) mismatch on rank {rank} for {N0}x{N1}"
            dist.barrier()
        a_nccl = KA.zeros(be, dt, N1, rb)
        st_s, st_r = KA.allocate(be, dt, rb * cb * world), KA.allocate(be, dt, rb * cb * world)
        mg.transpose_alltoall_nccl_(comm, be, a_nccl, b_loc, N0, N1, st_s, st_r)
        KA.synchronize(be)
        assert np.array_equal(KA.Array(a_nccl), want), f"NCCL distributed transpose mismatch on rank {rank} for {N0}x{N1}"
        dist.barrier()
        a_sym.free()
    out["distributed_transpose_fused_eq_nccl_eq_definition"] = True

    # ---- back-to-back calls with a consumer of the slab in between and NO barrier from the caller: the kernel's own
    # ---- write-after-read guard ("slab may be overwritten" flags) must keep call k+1 out of a slab that call k's consumer reads ----
    N0 = N1 = 2048 if 2048 % (64 * world) == 0 else 64 * world * 4
    rb, cb = N0 // world, N1 // world
    a_sym = mg.SymBuffer(N1 * rb * 4, world, rank, gather)
    a_view = a_sym.array(np.float32, (N1, rb))
    keep = [KA.allocate(be, np.float32, N1, rb) for _ in range(6)]
    wholes, b_locs = [], []
    for k in range(6):
        rng = np.random.default_rng(900 + k)
        whole = np.asfortranarray(rng.standard_normal((N0, N1)).astype(np.float32))
        wholes.append(whole)
        b_locs.append(KA.to_device(np.asfortranarray(whole[:, rank * cb:(rank + 1) * cb])))   # kept alive: freeing would synchronise
    KA.synchronize(be)
    dist.barrier()
    for k in range(6):
        mg.transpose_alltoall_(a_sym, b_locs[k], N0, N1)
        for _ in range(1 + 7 * (rank == 0)):       # rank 0 is a slow consumer: its peers reach call k+1 first
            KA.copyto_(be, keep[k], a_view)
    KA.synchronize(be)
    for k in range(6):
        assert np.array_equal(KA.Array(keep[k]), wholes[k].T[:, rank * rb:(rank + 1) * rb]), f"back-to-back call {k} on rank {rank}"
    dist.barrier()
    a_sym.free()
    out["distributed_transpose_back_to_back"] = True

    # ---- timing: NCCL route vs fused route at the C3 size ----
    if "--time" in sys.argv:
        # distributed transpose of ONE 16384 x 16384 Float32 matrix sharded over the ranks (strong scaling)
        N = 16384
        rb = N // world
        b_loc = KA.rand_uniform_(KA.allocate(be, np.float32, N, rb), 1235 + rank)
        a_sym = mg.SymBuffer(N * rb * 4, world, rank, gather)
        a_nccl = KA.zeros(be, np.float32, N, rb)
        st_s, st_r = KA.allocate(be, np.float32, rb * rb * world), KA.allocate(be, np.float32, rb * rb * world)
        ev0, ev1 = C.c_void_p(), C.c_void_p()
        _lib.call("b200_event_create", C.byref(ev0))
        _lib.call("b200_event_create", C.byref(ev1))

        def timed_t(fn, steps=20, warm=3):
            for _ in range(warm):
                fn()
            KA.synchronize(be)
            dist.barrier()
            _lib.call("b200_event_record", ev0)
            for _ in range(steps):
                fn()
            _lib.call("b200_event_record", ev1)
            KA.synchronize(be)
            ms = C.c_float()
            _lib.call("b200_event_elapsed_ms", ev0, ev1, C.byref(ms))
            t = torch.tensor([ms.value / steps], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        ms_f = timed_t(lambda: mg.transpose_alltoall_(a_sym, b_loc, N, N))
        ms_n = timed_t(lambda: mg.transpose_alltoall_nccl_(comm, be, a_nccl, b_loc, N, N, st_s, st_r))
        out["dist_transpose_16384_fused_ms"] = ms_f
        out["dist_transpose_16384_fused_GBps"] = 2.0 * 4 * N * N / (ms_f * 1e-3) / 1e9
        out["dist_transpose_16384_nccl_ms"] = ms_n
        out["dist_transpose_16384_nccl_GBps"] = 2.0 * 4 * N * N / (ms_n * 1e-3) / 1e9
        KA.synchronize(be)
        dist.barrier()
        a_sym.free()
        del b_loc, a_nccl, st_s, st_r

        n = 1 << 30
        s0, s1 = mg.histogram_shard(n, world, rank)
        d_in = EX.move(be, orc.gen_bins_i32(s1 - s0, 1236 + rank, 256))
        bins = KA.zeros(be, np.int32, 256)
        e0, e1 = C.c_void_p(), C.c_void_p()
        _lib.call("b200_event_create", C.byref(e0))
        _lib.call("b200_event_create", C.byref(e1))

        def timed(fn, steps=50, warm=5):
            for _ in range(warm):
                fn()
            KA.synchronize(be)
            dist.barrier()
            _lib.call("b200_event_record", e0)
            for _ in range(steps):
                fn()
            _lib.call("b200_event_record", e1)
            KA.synchronize(be)
            ms = C.c_float()
            _lib.call("b200_event_elapsed_ms", e0, e1, C.byref(ms))
            t = torch.tensor([ms.value / steps], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def step_nccl():
            EX.histogram_(bins, d_in)
            comm.allreduce_sum_i32(bins.ptr, 256)

        def step_fused():
            peers.histogram_allreduce(bins, d_in)

        def step_local():
            EX.histogram_(bins, d_in)

        for name, fn in (("local_only", step_local), ("nccl", step_nccl), ("fused", step_fused)):
            ms = timed(fn)
            out[f"hist_{name}_ms"] = ms
            out[f"hist_{name}_GBps"] = 4.0 * n / (ms * 1e-3) / 1e9
            if name == "fused":
                peers.timeline()                      # reset the min / max stamps, then one more step
                dist.barrier()
                step_fused()
                stamps = peers.timeline()
                gathered = [None] * world
                dist.all_gather_object(gathered, stamps)
                out["fused_timeline_ns_per_rank"] = gathered

    KA.synchronize(be)
    dist.barrier()
    peers.free()
    comm.destroy()
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
