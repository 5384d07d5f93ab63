"""CUDA-graph recording of launch-bound sequences (b200_graph_*): the replay gives the bits of the eager sequence, a
synchronising call inside a recording is refused, and the replay is measurably cheaper than re-issuing the launches."""
import ctypes as C
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))

pytestmark = pytest.mark.gpu

import b200ka as KA  # noqa: E402
from b200ka import _lib, examples as EX, kernels as K  # noqa: E402

be = KA.B200Backend()


def sequence(A, T, bins, vals, X, Y, Z, P):
    """a launch-bound mix of the testsuite / example kernels (test/localmem.jl, test/private.jl, examples/*)"""
    K.localmem(be, 16)(A, ndrange=(64,))
    K.private(be, 16)(P, ndrange=(64,))
    EX.naive_transpose_(T, X)
    bins.fill_(0)
    EX.histogram_(bins, vals)
    K.saxpy_kernel_(be, 256)(Z, np.float32(2.0), Y, Y, ndrange=Z.shape)
    EX.mycopy_(Y, Z)                                   # Y <- 3 Y: every replay changes the data


def make():
    rng = np.random.default_rng(0)
    return dict(A=KA.to_device(np.full(64, -1, dtype=np.int64)), P=KA.to_device(np.full(64, -1, dtype=np.int64)),
                X=KA.to_device(np.asfortranarray(rng.standard_normal((64, 128)).astype(np.float32))), T=KA.zeros(be, np.float32, 128, 64),
                vals=KA.to_device(rng.integers(1, 257, 5000).astype(np.int32)), bins=KA.zeros(be, np.int32, 256),
                Y=KA.to_device(rng.standard_normal(4096).astype(np.float32)), Z=KA.zeros(be, np.float32, 4096))


def test_replay_equals_eager_sequence():
    eager, rec = make(), make()
    for _ in range(3):
        sequence(**eager)
    KA.synchronize(be)
    with KA.Graph(be) as g:
        sequence(**rec)                                # recorded, not executed
    KA.synchronize(be)
    assert np.all(KA.Array(rec["A"]) == -1) and np.all(KA.Array(rec["T"]) == 0)
    for _ in range(3):
        g.launch()
    KA.synchronize(be)
    for k in eager:
        assert np.array_equal(KA.Array(eager[k]), KA.Array(rec[k])), k


def test_synchronising_call_invalidates_the_recording():
    d = make()
    with pytest.raises(_lib.B200Error):
        with KA.Graph(be):
            sequence(**d)
            KA.maximum(d["vals"])                      # returns a value -> must wait for the stream -> not recordable
    KA.synchronize(be)
    sequence(**d)                                      # the thread's stream is usable again
    KA.synchronize(be)
    assert np.array_equal(KA.Array(d["A"])[:16], np.arange(16, 0, -1))


def test_replay_is_cheaper_than_relaunching():
    d = make()
    n_rep = 50
    sequence(**d)
    KA.synchronize(be)
    t0 = time.perf_counter()
    for _ in range(n_rep):
        sequence(**d)
    KA.synchronize(be)
    eager = (time.perf_counter() - t0) / n_rep
    with KA.Graph(be) as g:
        sequence(**d)
    g.launch()
    KA.synchronize(be)
    t0 = time.perf_counter()
    for _ in range(n_rep):
        g.launch()
    KA.synchronize(be)
    replay = (time.perf_counter() - t0) / n_rep
    print(f"7-launch sequence: eager {eager * 1e6:.1f} us, graph replay {replay * 1e6:.1f} us")
    assert replay < eager
