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


def test_task_style_streams_bind_create_destroy():
    """What the Julia layer does per task (reference keeps queue + device in task_local_storage, src/pocl/pocl.jl:12-41): own
    stream handles (b200_stream_create), bound together with the device before every call (b200_bind), from WHICHEVER OS thread
    the task happens to run on -- here two "tasks" whose calls alternate between two OS threads."""
    dev = C.c_int32()
    _lib.call("b200_get_device", C.byref(dev))
    streams = []
    for prio in (0, 1):
        s = C.c_void_p()
        _lib.call("b200_stream_create", C.byref(s), prio)
        streams.append(s)
    assert streams[0].value != streams[1].value
    n = 1 << 20
    src = [KA.to_device(np.full(n, k + 1, dtype=np.float32)) for k in range(2)]
    dst = [KA.zeros(be, np.float32, n) for _ in range(2)]
    KA.synchronize(be)
    errors = []

    def step(task, thread_tag):
        try:
            _lib.call("b200_bind", dev.value, streams[task])                  # the task's state -> this OS thread
            got = C.c_void_p()
            _lib.call("b200_get_stream", C.byref(got))
            assert got.value == streams[task].value
            _lib.call("b200_copy", C.c_void_p(dst[task].ptr), C.c_void_p(src[task].ptr), C.c_size_t(4 * n))
        except BaseException as ex:  # noqa: BLE001
            errors.append((task, thread_tag, repr(ex)))

    for rnd in range(4):             # task 0 and task 1 swap OS threads every round
        ts = [threading.Thread(target=step, args=((rnd + k) % 2, k)) for k in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    assert not errors, errors
    for task in range(2):            # cooperative wait on the TASK's stream, from the main thread
        _lib.call("b200_bind", dev.value, streams[task])
        while _lib.lib().b200_stream_query() == _lib.B200_NOT_READY:
            pass
        assert np.all(KA.Array(dst[task]) == task + 1)
    _lib.call("b200_bind", dev.value, None)                                    # back to the thread's own stream
    for s in streams:
        _lib.call("b200_stream_destroy", s)
    with pytest.raises(_lib.B200Error):
        bad = C.c_void_p()
        _lib.call("b200_stream_create", C.byref(bad), 7)                       # priority out of range
