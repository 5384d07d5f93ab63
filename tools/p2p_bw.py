import ctypes as C, os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
import torch, torch.distributed as dist
import b200ka as KA
from b200ka import _lib, multigpu as mg
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.call("b200_set_device", local)
be = KA.B200Backend()
nbytes = 256 << 20
sym = mg.SymBuffer(nbytes, world, rank, mg.torch_gather(dist))
src = KA.allocate(be, np.uint8, nbytes).fill_(1)
e0, e1 = C.c_void_p(), C.c_void_p()
_lib.call("b200_event_create", C.byref(e0)); _lib.call("b200_event_create", C.byref(e1))
peer = sym.ptr((rank + 1) % world)
def timed(fn, steps=20):
    for _ in range(3): fn()
    KA.synchronize(be); dist.barrier()
    _lib.call("b200_event_record", e0)
    for _ in range(steps): fn()
    _lib.call("b200_event_record", e1); KA.synchronize(be)
    ms = C.c_float(); _lib.call("b200_event_elapsed_ms", e0, e1, C.byref(ms))
    return ms.value / steps
ms = timed(lambda: _lib.call("b200_memcpy_d2d_async", C.c_void_p(peer), C.c_void_p(src.ptr), C.c_size_t(nbytes)))
print(rank, "cudaMemcpyAsync to peer (all ranks at once, ring):", round(nbytes / ms / 1e6, 1), "GB/s", flush=True)
for knobs in ({"copy.impl": 1, "copy.bulk_stage_kb": 16}, {"copy.impl": 1, "copy.bulk_stage_kb": 4, "copy.bulk_stages": 8},
              {"copy.impl": 1, "copy.bulk_stage_kb": 1, "copy.bulk_stages": 8}, {"copy.impl": 1, "copy.bulk_stage_kb": 32, "copy.bulk_stages": 4},
              {"copy.impl": 1, "copy.bulk_stage_kb": 16, "copy.bulk_ctas_per_sm": 2}, {"copy.impl": 0}):
    for k, v in knobs.items():
        _lib.call("b200_tune_set", k.encode(), v)
    ms = timed(lambda: _lib.call("b200_copy", C.c_void_p(peer), C.c_void_p(src.ptr), C.c_size_t(nbytes)))
    if rank == 0:
        print(rank, "b200_copy to peer", knobs, ":", round(nbytes / ms / 1e6, 1), "GB/s", flush=True)
    for k in knobs:
        _lib.call("b200_tune_set", k.encode(), {"copy.impl": 1, "copy.bulk_stage_kb": 16, "copy.bulk_stages": 4, "copy.bulk_ctas_per_sm": 1}[k])
KA.synchronize(be); dist.barrier(); sym.free(); dist.destroy_process_group()
