"""CPU-only, world_size 2 over gloo: the sharding plan of SURVEY.md 8e (slabs for copy/transpose, partial bins +
all-reduce for histogram, A/C row panels with replicated B for matmul) reassembles to the single-device result.
The per-rank compute is done by the oracle here (no GPU in this container); on the GPU box bench.py runs the same
plan with the CUDA kernels and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "kernelabstractions.jl_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import ka_oracle as orc
    from b200ka import multigpu as mg

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    # histogram: contiguous input slabs -> partial bins -> all-reduce(sum)
    n = 100003
    inp = orc.gen_bins_i32(n, 1236, 256)
    s0, s1 = mg.histogram_shard(n, world, rank)
    part = np.zeros(256, dtype=np.int32)
    orc.histogram(part, np.ascontiguousarray(inp[s0:s1]), 256)
    t = torch.from_numpy(part)
    mg.allreduce_bins_torch(t)
    ok &= np.array_equal(t.numpy(), orc.create_histogram(inp))
    # transpose: rank owns output columns [j0, j1) of a; needs rows [j0, j1) of b
    nn = 256
    b = orc.gen_uniform_f32((nn, nn), 1235)
    j0, j1 = mg.transpose_shard(nn, world, rank)
    b_slab = np.asfortranarray(b[j0:j1, :])                 # (j1-j0) x n, ldb = j1-j0
    a_slab = np.zeros((nn, j1 - j0), dtype=np.float32, order="F")
    orc.naive_transpose(a_slab, b_slab)
    gathered = [torch.zeros(nn * (mg.transpose_shard(nn, world, r)[1] - mg.transpose_shard(nn, world, r)[0])) for r in range(world)]
    dist.all_gather(gathered, torch.from_numpy(a_slab.ravel(order="F").copy()))
    a_full = np.concatenate([g.numpy() for g in gathered]).reshape((nn, nn), order="F")
    ok &= np.array_equal(a_full, b.T)
    # matmul: A/C row panels, B replicated; k is not split so the FP32 bits equal the 1-device result
    M, Kd, N = 256, 96, 64
    A = orc.gen_uniform_f32((M, Kd), 1237)
    B = orc.gen_uniform_f32((Kd, N), 1238)
    m0, m1 = mg.matmul_shard(M, world, rank)
    Cp = np.zeros((m1 - m0, N), dtype=np.float32, order="F")
    orc.coalesced_matmul(Cp, np.asfortranarray(A[m0:m1, :]), B)
    Cfull = np.zeros((M, N), dtype=np.float32, order="F")
    orc.coalesced_matmul(Cfull, A, B)
    ok &= np.array_equal(Cp, Cfull[m0:m1, :])
    # timing reduction used by bench.py: max over ranks
    tmax = torch.tensor([float(rank + 1)])
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ok &= tmax.item() == float(world)
    open(os.path.join(out_dir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_plan_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"
