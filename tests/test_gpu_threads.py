"""The C ABI is callable concurrently from several OS threads (SURVEY.md 8(b) "Threading"; the reference keeps its runtime state
per task, src/pocl/pocl.jl:12-41, and examples/mpi.jl:26-36 launches from `Threads.@spawn` tasks): every thread gets its own
stream, per-thread error strings and scratch buffers, and results do not interfere."""
import ctypes as C
import os
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))

pytestmark = pytest.mark.gpu

import b200ka as KA  # noqa: E402
from b200ka import _lib, examples as EX  # noqa: E402

be = KA.B200Backend()


def test_concurrent_host_threads_have_private_streams_and_correct_results():
    nthreads, rounds = 6, 8
    streams, errors = [None] * nthreads, []

    def work(tid):
        try:
            s = C.c_void_p()
            _lib.call("b200_get_stream", C.byref(s))
            streams[tid] = s.value
            rng = np.random.default_rng(tid)
            for r in range(rounds):
                m, n = 64 * (1 + (tid + r) % 5), 64 * (1 + (3 * tid + r) % 7)
                hb = np.asfortranarray(rng.standard_normal((n, m)).astype(np.float32))
                b, a = KA.to_device(hb), KA.zeros(be, np.float32, m, n)
                EX.naive_transpose_(a, b)
                vals = rng.integers(1, 257, 50000 + 1000 * tid).astype(np.int32)
                d = KA.to_device(vals)
                bins = KA.zeros(be, np.int32, 256)
                EX.histogram_(bins, d)
                # value-returning calls wait on THIS thread's stream only
                assert int(KA.maximum(d)) == int(vals.max()) and int(KA.sum(bins)) == vals.size
                assert np.array_equal(KA.Array(bins), np.bincount(vals, minlength=257)[1:].astype(np.int32))
                assert np.array_equal(KA.Array(a), hb.T)
                A, B = KA.to_device(hb[:64, :m].copy(order="F")), KA.to_device(hb[:64, :m].T.copy(order="F"))
                Cd = KA.matmul(A, B)
                ref = KA.Array(A).astype(np.float64) @ KA.Array(B).astype(np.float64)
                assert np.allclose(KA.Array(Cd), ref, rtol=1e-4, atol=1e-4)
            # an error in one thread is reported in that thread only
            with pytest.raises(_lib.B200Error):
                _lib.call("b200_transpose", None, None, 4, 4, 2, 2, 4)
            assert b"bad shape" in _lib.lib().b200_last_error()
        except BaseException as ex:  # noqa: BLE001
            errors.append((tid, repr(ex)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    assert len(set(streams)) == nthreads, "every host thread must get its own stream"
    _lib.call("b200_device_synchronize")
