import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, "kernelabstractions.jl_b200")
import b200ka as KA
from b200ka import _lib, examples as EX
be = KA.B200Backend()
ev0, ev1 = C.c_void_p(), C.c_void_p()
_lib.call("b200_event_create", C.byref(ev0)); _lib.call("b200_event_create", C.byref(ev1))
def timed(fn, steps=30, warm=5):
    for _ in range(warm): fn()
    KA.synchronize(be)
    _lib.call("b200_event_record", ev0)
    for _ in range(steps): fn()
    _lib.call("b200_event_record", ev1)
    KA.synchronize(be)
    ms = C.c_float(); _lib.call("b200_event_elapsed_ms", ev0, ev1, C.byref(ms))
    return ms.value / steps * 1e3
for ne in (1 << 30, 1 << 27):
    d_in = KA.rand_bins_(KA.allocate(be, np.int32, ne), 1236, 256)
    bins = KA.zeros(be, np.int32, 256)
    for T, P in ((1024, 4), (1024, 2), (1024, 8), (512, 4), (512, 8), (768, 4), (256, 8)):
        for cps in (1, 2):
            if T == 1024 and cps == 2: continue
            _lib.call("b200_tune_set", b"hist.atomic_threads", T); _lib.call("b200_tune_set", b"hist.atomic_batch", P); _lib.call("b200_tune_set", b"hist.atomic_ctas_per_sm", cps)
            try:
                us = timed(lambda: EX.histogram_(bins, d_in))
                print(f"n=2^{ne.bit_length()-1} T={T} P={P} ctas/sm={cps}: {us:8.2f} us {4*ne/us/1e3:7.0f} GB/s", flush=True)
            except Exception as ex:
                print(T, P, cps, "error", str(ex)[:80])
    del d_in
