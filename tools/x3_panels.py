#!/usr/bin/env python
"""3xTF32 (b200_matmul_f32_tf32x3) on the row-panel shapes of the multi-GPU shard (M x 8192 x 8192): wide (256 x 256 pair tiles),
narrow (256 x 128) and the automatic choice; CUDA-event time (split pass included) and the error of 64 columns against Float64."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200ka as KA, ka_oracle as orc
from b200ka import _lib
be = KA.B200Backend()
K = N = 8192
B = orc.gen_uniform_f32((K, N), 1238); dB = KA.to_device(B)
def ev():
    e = C.c_void_p(); _lib.call("b200_event_create", C.byref(e)); return e
for M in (8192, 4096, 2048, 1024, 512, 256, 128):
    A = orc.gen_uniform_f32((M, K), 1237); dA = KA.to_device(A)
    dC = KA.zeros(be, np.float32, M, N); ws = KA.allocate(be, np.float32, M*K + K*N)
    f = lambda: _lib.call("b200_matmul_f32_tf32x3", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, M, K, M, C.c_void_p(ws.ptr))
    for w in (1, 0, -1):
        _lib.call("b200_tune_set", b"tcgemm.x3_wide", w)
        f(); KA.synchronize(be)
        a, b = ev(), ev(); _lib.call("b200_event_record", a)
        for _ in range(5): f()
        _lib.call("b200_event_record", b); KA.synchronize(be)
        ms = C.c_float(); _lib.call("b200_event_elapsed_ms", a, b, C.byref(ms)); ms = ms.value/5
        got = KA.Array(dC)[:, :64]
        truth = A.astype(np.float64) @ B[:, :64].astype(np.float64)
        err = (np.abs(got-truth)/np.abs(truth)).max()
        print(f"M={M} wide={w}: {ms:.3f} ms {2.0*M*K*N/ms/1e9:.1f} TFLOP/s err {err:.2e}", flush=True)
