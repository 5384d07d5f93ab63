"""Timing of the distributed transpose's data kernel with and without its completion barrier (torchrun, one rank per GPU)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
import torch, torch.distributed as dist
import b200ka as KA
from b200ka import _lib, multigpu as mg
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.call("b200_set_device", local)
be = KA.B200Backend()
N = 16384
rb = N // world
b_loc = KA.rand_uniform_(KA.allocate(be, np.float32, N, rb), 1235 + rank)
a_sym = mg.SymBuffer(N * rb * 4, world, rank, mg.torch_gather(dist))
e0, e1 = C.c_void_p(), C.c_void_p()
_lib.call("b200_event_create", C.byref(e0)); _lib.call("b200_event_create", C.byref(e1))
def timed(fn, steps=20):
    for _ in range(3): fn()
    KA.synchronize(be); dist.barrier()
    _lib.call("b200_event_record", e0)
    for _ in range(steps): fn()
    _lib.call("b200_event_record", e1); KA.synchronize(be)
    ms = C.c_float(); _lib.call("b200_event_elapsed_ms", e0, e1, C.byref(ms))
    t = torch.tensor([ms.value / steps], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
for order in (0, 1):
    for tma in (0, 1):
        _lib.call("b200_tune_set", b"a2a.tma_store", tma)
        _lib.call("b200_tune_set", b"a2a.order", order)
        ms = timed(lambda: mg.transpose_alltoall_(a_sym, b_loc, N, N))
        remote = N * rb * 4 * (world - 1) / world
        if rank == 0:
            print(f"world {world} order={order} tma_store={tma}: {ms:.4f} ms/step, NVLink egress {remote / ms / 1e6:.0f} GB/s per GPU", flush=True)
_lib.call("b200_tune_set", b"a2a.tma_store", 0)
_lib.call("b200_tune_set", b"a2a.order", 1)
KA.synchronize(be); dist.barrier(); a_sym.free(); dist.destroy_process_group()
