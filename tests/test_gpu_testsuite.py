"""GPU: the reference's testsuite files for the hot path, replayed against B200Backend through the C ABI.

Each test is the Python spelling of one reference testset (cited), asserting the same values; where the reference
only checks that a launch completes we additionally compare with the oracle.
"""
import numpy as np
import pytest

import b200ka as KA
from b200ka import kernels as K
from b200ka.kernel import Val, ki_kernel

pytestmark = pytest.mark.gpu
backend = KA.B200Backend


def ArrayT(dtype, *shape):
    return KA.B200Array(dtype, shape)


# ---- test/localmem.jl:67-79 -------------------------------------------------------------------------------
@pytest.mark.parametrize("kname", ["localmem", "localmem2", "localmem_unsafe_indices", "many_localmem"])
def test_localmem(kname):
    kernel_ = getattr(K, kname)(backend(), 16)
    A = ArrayT(np.int64, 64)
    kernel_(A, ndrange=A.shape)
    KA.synchronize(backend())
    B = KA.Array(A)
    for s in range(0, 64, 16):
        assert np.array_equal(B[s:s + 16], np.arange(16, 0, -1))


@pytest.mark.parametrize("variant,kname", list(enumerate(["localmem", "localmem2", "localmem_unsafe_indices", "many_localmem"])))
def test_localmem_ragged_vs_oracle(orc, variant, kname):
    # ndrange not a multiple of the workgroup: the DynamicCheck path
    A = ArrayT(np.int64, 50).fill_(-7)
    getattr(K, kname)(backend(), 16)(A, ndrange=A.shape)
    KA.synchronize(backend())
    ref = np.full(50, -7, dtype=np.int64)
    orc.localmem(variant, ref, (50,), 16)
    got = KA.Array(A)
    if variant == 2:
        assert np.array_equal(got, ref)
    else:
        # lanes of the last group whose mirror lane is inactive read uninitialised local memory in the reference
        assert np.array_equal(got[:48], ref[:48])


# ---- test/private.jl:61-91 ----------------------------------------------------------------------------------
def test_private():
    K.stmt_form(backend(), 16)(ndrange=16)
    KA.synchronize(backend())
    A = ArrayT(np.int64, 64)
    K.private(backend(), 16)(A, ndrange=A.shape)
    KA.synchronize(backend())
    B = KA.Array(A)
    for s in range(0, 64, 16):
        assert np.array_equal(B[s:s + 16], np.arange(16, 0, -1))

    A = ArrayT(np.int64, 64, 64)
    A.fill_(1)  # A .= 1
    K.forloop(backend())(A, Val(A.shape[1]), ndrange=A.shape[0], workgroupsize=A.shape[0])
    KA.synchronize(backend())
    h = KA.Array(A)
    assert np.all(h[:, 0] == 64)
    assert np.all(h[:, 1:] == 1)

    B = ArrayT(np.bool_, *A.shape)
    K.typetest(backend(), 16)(A, B, ndrange=A.shape)
    KA.synchronize(backend())
    assert np.all(KA.Array(B))

    A = KA.to_device(np.ones((64, 3), dtype=np.float32))
    out = ArrayT(np.float32, 64)
    K.reduce_private(backend(), 8)(out, A, ndrange=out.shape)
    KA.synchronize(backend())
    assert np.all(KA.Array(out) == 3.0)


def test_reduce_private_random_vs_oracle(orc):
    rng = np.random.default_rng(3)
    h = np.asfortranarray(rng.standard_normal((100, 37)).astype(np.float32))
    out = ArrayT(np.float32, 100)
    K.reduce_private(backend(), 8)(out, KA.to_device(h), ndrange=out.shape)
    ref = np.zeros(100, dtype=np.float32)
    orc.reduce_private(ref, h, (100,), 8)
    assert np.array_equal(KA.Array(out), ref)  # same k-ascending sum, bit for bit


# ---- test/unroll.jl:39-48 ----------------------------------------------------------------------------------------
def test_unroll(orc):
    a = KA.to_device(np.zeros(5, dtype=np.float32))
    kernel_ = K.kernel_unroll_(backend(), 1, 1)
    kernel_(a)
    KA.synchronize(backend())
    assert np.array_equal(KA.Array(a), [1, 2, 3, 4, 5])
    kernel_(a, Val(5))
    kernel2_ = K.kernel_unroll2_(backend(), 1, 1)
    kernel2_(a)
    KA.synchronize(backend())
    ref = np.zeros(5, dtype=np.float32)
    orc.unroll(0, ref)
    orc.unroll(1, ref, 5)
    orc.unroll(2, ref)
    assert np.array_equal(KA.Array(a), ref)


# ---- test/copyto.jl:5-21 --------------------------------------------------------------------------------------------
def test_copyto():
    M = 1024
    be = backend()
    ET = np.float64 if KA.supports_float64(be) else np.float32
    rng = np.random.default_rng(11)
    A = KA.to_device(rng.random(M).astype(ET))
    hb = rng.random(M).astype(ET)
    B = KA.to_device(hb)
    a = np.empty(M, dtype=ET)
    KA.copyto_(be, a, B)
    KA.synchronize(be)  # the host buffer is only valid after synchronize (async contract)
    KA.copyto_(be, A, a)
    KA.synchronize(be)
    assert np.array_equal(a, KA.Array(A))
    assert np.array_equal(a, KA.Array(B))
    assert np.array_equal(a, hb)


def test_copyto_matrix_layout():
    """Device arrays are column-major: a C-ordered host matrix (numpy's default) must arrive element for element, not byte for
    byte, and a C-ordered host destination is refused instead of being filled with permuted data."""
    be = backend()
    h = np.arange(6 * 4, dtype=np.float32).reshape(6, 4)          # C order
    assert not h.flags.f_contiguous
    dA = ArrayT(np.float32, 6, 4)
    KA.copyto_(be, dA, h)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(dA), h)
    out_f = np.empty((6, 4), dtype=np.float32, order="F")
    KA.copyto_(be, out_f, dA)
    KA.synchronize(be)
    assert np.array_equal(out_f, h)
    with pytest.raises(KA.ArgumentError, match="column-major"):
        KA.copyto_(be, np.empty((6, 4), dtype=np.float32), dA)


def test_pending_host_refs_are_per_stream():
    """synchronize() releases only the references of the stream it waited for (another thread's in-flight copy keeps its buffer)."""
    import threading

    from b200ka import backend as B

    be = backend()
    dA = ArrayT(np.float32, 1 << 16)
    h = np.ones(1 << 16, dtype=np.float32)
    KA.copyto_(be, dA, h)
    main_key = B._stream_key()
    assert main_key in B._pending_host_refs
    seen = {}

    def other():
        d2 = ArrayT(np.float32, 16)
        KA.copyto_(be, d2, np.zeros(16, dtype=np.float32))
        seen["key"] = B._stream_key()
        KA.synchronize(be)
        seen["main_still_held"] = main_key in B._pending_host_refs
        seen["own_released"] = seen["key"] not in B._pending_host_refs

    t = threading.Thread(target=other)
    t.start()
    t.join()
    assert seen["key"] != main_key and seen["main_still_held"] and seen["own_released"]
    KA.synchronize(be)
    assert main_key not in B._pending_host_refs


def test_copyto_device_checks():
    be = backend()
    A = ArrayT(np.float32, 100)
    B = ArrayT(np.float32, 101)
    with pytest.raises(KA.ErrorException, match="match in length"):
        KA.copyto_(be, A, B)
    with pytest.raises(KA.ErrorException, match="alias"):
        KA.copyto_(be, A, A)
    with pytest.raises(KA.ErrorException, match="alias"):
        KA.copyto_(be, A.view(0, (60,)), A.view(40, (60,)))
    C_ = KA.to_device(np.arange(100, dtype=np.float32))
    assert KA.copyto_(be, A, C_) is A
    KA.synchronize(be)
    assert np.array_equal(KA.Array(A), np.arange(100, dtype=np.float32))


# ---- test/test.jl:75-113 get_backend / adapt ----------------------------------------------------------------------------
def test_get_backend_and_adapt():
    be = backend()
    unified = KA.supports_unified(be)
    U = KA.allocate(be, np.float32, 5, unified=unified)
    U.fill_(2.0)
    if unified:
        assert isinstance(U[2], np.float32) and U[2] == 2.0
    x = KA.allocate(be, np.float32, 5)
    with pytest.raises(KA.ErrorException):
        x[2]
    A = KA.allocate(be, np.float32, 5, 5)
    assert isinstance(KA.get_backend(A), KA.B200Backend)
    assert isinstance(KA.get_backend(A.view(5, (5, 2))), KA.B200Backend)
    with pytest.raises(KA.ArgumentError):
        KA.get_backend("unknown")
    assert isinstance(KA.adapt(KA.CPU(), x), np.ndarray)
    y = KA.adapt(be, np.empty(5, dtype=np.float32))
    assert type(y) is type(x)


# ---- test/test.jl:116-160 indextest ---------------------------------------------------------------------------------------
def test_indextest():
    be = backend()
    A = KA.allocate(be, np.int64, 16, 16)
    K.index_linear_global(be, 8)(A, ndrange=A.size)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(A).ravel(order="F"), np.arange(1, 257))

    for shape in [(8,), (16,), (8, 2)]:
        A = KA.allocate(be, np.int64, *shape)
        K.index_linear_local(be, 8)(A, ndrange=A.size)
        KA.synchronize(be)
        flat = KA.Array(A).ravel(order="F")
        for s in range(0, flat.size, 8):
            assert np.array_equal(flat[s:s + 8], np.arange(1, 9))

    A = KA.allocate(be, KA.CartesianIndex(2), 16, 16)
    K.index_cartesian_global(be, 8)(A, ndrange=A.shape)
    KA.synchronize(be)
    h = KA.Array(A)["I"]
    for i in range(16):
        for j in range(16):
            assert tuple(h[i, j]) == (i + 1, j + 1)

    A = KA.allocate(be, KA.CartesianIndex(1), 16, 16)
    K.index_cartesian_global(be, 8)(A, ndrange=A.size)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(A)["I"].ravel(order="F"), np.arange(1, 257))

    # Non-multiples of the workgroupsize
    A = KA.allocate(be, np.int64, 7, 7)
    K.index_linear_global(be, 8)(A, ndrange=A.size)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(A).ravel(order="F"), np.arange(1, 50))

    A = KA.allocate(be, np.int64, 5)
    K.index_linear_local(be, 8)(A, ndrange=A.size)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(A), np.arange(1, 6))


def test_index_group_and_local_cartesian_vs_oracle(orc):
    be = backend()
    A = KA.allocate(be, np.int64, 40).fill_(0)
    K.index_linear_group(be, 8)(A, ndrange=A.size)
    ref = np.zeros(40, dtype=np.int64)
    orc.index_linear(2, ref, 40, 8)
    assert np.array_equal(KA.Array(A), ref)
    for which, kern in [(1, K.index_cartesian_local), (2, K.index_cartesian_group)]:
        A = KA.allocate(be, KA.CartesianIndex(2), 12, 10).fill_(0)
        kern(be, (4, 2))(A, ndrange=A.shape)
        ref = np.zeros((2, 12, 10), dtype=np.int64, order="F")
        orc.index_cartesian(which, ref, (12, 10), (12, 10), (4, 2))
        got = np.moveaxis(KA.Array(A)["I"], -1, 0)
        assert np.array_equal(got, ref)


# ---- test/test.jl:216-237 ------------------------------------------------------------------------------------------------
def test_kernel_val_and_zero_iteration_space():
    be = backend()
    A = KA.zeros(be, np.int64, 1024)
    K.kernel_val_(be)(A, Val(3), ndrange=A.shape)
    KA.synchronize(be)
    assert np.all(KA.Array(A) == 3)

    K.kernel_empty(be, 1)(ndrange=1)
    K.kernel_empty(be, 1)(ndrange=0)
    KA.synchronize(be)
    K.kernel_empty(be, 1)(ndrange=0)
    KA.synchronize(be)


def test_zero_size_allocations():
    # test/test.jl:290-305
    be = backend()
    for dims in [(0,), (0, 5), (3, 0, 2)]:
        A = KA.allocate(be, np.float32, dims)
        assert A.size == 0 and KA.Array(A).shape == dims
        Z = KA.zeros(be, np.float32, dims)
        assert Z.size == 0
    with pytest.raises(KA.MethodError):
        KA.allocate(be, object(), (4,))


# ---- test/test.jl:275-321, 339-385 -----------------------------------------------------------------------------------------
def test_context_kernel_and_gpu_return():
    be = backend()
    A = KA.zeros(be, np.int64, 1024)
    K.context_kernel(be)(A, ndrange=A.shape)          # "No CPU kernel" (test/test.jl:284-289)
    KA.synchronize(be)
    assert np.all(KA.Array(A) == 1)

    A = KA.zeros(be, np.int64, 1024)
    K.gpu_return_kernel_(be)(A, ndrange=A.size)       # "GPU kernel return statement" (test/test.jl:314-321)
    KA.synchronize(be)
    Ah = KA.Array(A)
    assert np.all(Ah[:512] == 1) and np.all(Ah[512:] == 0)


def test_unaliased_accumulate(orc):
    # test/test.jl:358-385: N = 8, M = 5, input[i,k] = i + k
    be = backend()
    N, M = 8, 5
    inp = np.asfortranarray(np.array([[i + k for k in range(1, N + 1)] for i in range(1, M + 1)], dtype=np.float32))
    reference = np.zeros((M, N), dtype=np.float32, order="F")
    for i in range(M):
        for j in range(N):
            for k in range(j, N):
                reference[i, j] = np.float32(reference[i, j] + inp[i, k])
    d_in = KA.adapt(be, inp)
    out = KA.zeros(be, np.float32, M, N)
    K.unaliased_accumulate_(be)(out, d_in, N, ndrange=out.shape)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(out), reference)
    out.fill_(0)
    K.unaliased_accumulate_local_(be)(out, d_in, N, ndrange=out.shape)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(out), reference)
    # random data, ragged ndrange, against the oracle (bit-exact: plain adds, k ascending)
    inp = np.asfortranarray(orc.gen_uniform_f32((37, 50), 9))
    d_in = KA.adapt(be, inp)
    for local, ctor in ((False, K.unaliased_accumulate_), (True, K.unaliased_accumulate_local_)):
        out = KA.zeros(be, np.float32, 37, 50)
        ctor(be, (8, 4))(out, d_in, 50, ndrange=(33, 47))
        KA.synchronize(be)
        ref = np.zeros((37, 50), dtype=np.float32, order="F")
        orc.unaliased_accumulate(local, ref, inp, 50, (33, 47), (8, 4))
        assert np.array_equal(KA.Array(out), ref)


# ---- test/convert.jl:46-73, test/specialfunctions.jl:20-59 -----------------------------------------------------------------
@pytest.mark.parametrize("ET", [np.float64, np.float32])
def test_convert_testsuite(ET):
    be = backend()
    N = 32
    rng = np.random.default_rng(3)
    A = (rng.random(N) * 3).astype(ET)
    A[:4] = [0.5, 1.5, 2.5, 2.0]          # ties: round() is round-half-to-even in Julia
    d_A = KA.to_device(A)
    d_B = KA.zeros(be, ET, N, 30)
    K.convert_kernel_(be, 4)(d_A, d_B, ndrange=N)
    KA.synchronize(be)
    B = KA.Array(d_B)
    for i in range(1, 11):
        if i in (5, 10):
            assert np.all(B[:, i - 1] == 0) and np.all(B[:, i + 9] == 0) and np.all(B[:, i + 19] == 0)   # Int128 lines are commented out
            continue
        assert np.array_equal(B[:, i - 1], np.ceil(A))
        assert np.array_equal(B[:, i + 9], np.floor(A))
        assert np.array_equal(B[:, i + 19], np.rint(A))


def test_specialfunctions_testsuite():
    import math
    be = backend()
    rtol = float(np.sqrt(np.finfo(np.float32).eps))     # Julia's isapprox default for Float32
    cases = [(K.gamma_knl, math.gamma, [1.0, 2.0, 3.0, 5.5]),
             (K.erf_knl, math.erf, [-1.0, -0.5, 0.0, 1.0e-3, 1.0, 2.0, 5.5]),
             (K.erfc_knl, math.erfc, [-1.0, -0.5, 0.0, 1.0e-3, 1.0, 2.0, 5.5])]
    for kern, f, xs in cases:
        x = np.array(xs, dtype=np.float32)
        cx = KA.allocate(be, np.float32, len(x))
        KA.copyto_(be, cx, x)
        cy = KA.similar(cx)
        kern(be)(cy, cx, ndrange=len(x))
        KA.synchronize(be)
        want = np.array([f(float(v)) for v in x], dtype=np.float64)
        assert np.allclose(KA.Array(cy).astype(np.float64), want, rtol=rtol, atol=0.0), kern.fname


# ---- test/intrinsics.jl:27-124 ---------------------------------------------------------------------------------------------
def test_ki_launch_parameters():
    be = backend()
    arr1d = KA.to_device(np.zeros(4, dtype=np.float32))
    ki_kernel(be, "launch_kernel1d", arr1d, numworkgroups=2, workgroupsize=2)
    KA.synchronize(be)
    assert np.all(KA.Array(arr1d) == 1)
    arr1dt = KA.to_device(np.zeros(4, dtype=np.float32))
    ki_kernel(be, "launch_kernel1d", arr1dt, numworkgroups=(2,), workgroupsize=(2,))
    KA.synchronize(be)
    assert np.all(KA.Array(arr1dt) == 1)
    arr2d = KA.to_device(np.zeros((4, 4), dtype=np.float32))
    ki_kernel(be, "launch_kernel2d", arr2d, numworkgroups=(2, 2), workgroupsize=(2, 2))
    KA.synchronize(be)
    assert np.all(KA.Array(arr2d) == 1)
    arr3d = KA.to_device(np.zeros((4, 4, 4), dtype=np.float32))
    ki_kernel(be, "launch_kernel3d", arr3d, numworkgroups=(2, 2, 2), workgroupsize=(2, 2, 2))
    KA.synchronize(be)
    assert np.all(KA.Array(arr3d) == 1)
    with pytest.raises(KA.ArgumentError):
        ki_kernel(be, "launch_kernel3d", arr3d, numworkgroups=(2, 2, 2, 2), workgroupsize=(2, 2, 2))
    with pytest.raises(KA.ArgumentError):
        ki_kernel(be, "launch_kernel3d", arr3d, numworkgroups=(2, 2, 2), workgroupsize=(2, 2, 2, 2))


def test_ki_basic_intrinsics():
    be = backend()
    assert isinstance(KA.max_work_group_size(be), int) and KA.max_work_group_size(be) == 1024
    assert isinstance(KA.multiprocessor_count(be), int) and KA.multiprocessor_count(be) > 0
    workgroupsize, numworkgroups = 4, 4
    N = workgroupsize * numworkgroups
    results = KA.allocate(be, KA.KernelData, N)
    kernel = ki_kernel(be, "test_intrinsics_kernel", results, launch=False)
    assert isinstance(KA.kernel_max_work_group_size(kernel), int)
    assert KA.kernel_max_work_group_size(kernel, max_work_items=1) == 1
    kernel(results, workgroupsize=workgroupsize, numworkgroups=numworkgroups)
    KA.synchronize(be)
    host = KA.Array(results)
    for i, k in enumerate(host, start=1):
        assert k["global_id"] == i
        assert k["global_size"] == N
        assert k["local_size"] == workgroupsize
        assert k["num_groups"] == numworkgroups
        assert k["group_id"] == (i - 1) // numworkgroups + 1
        assert k["local_id"] == (i - 1) % workgroupsize + 1


# ---- test/devices.jl ------------------------------------------------------------------------------------------------------------
def test_devices():
    be = backend()
    n = KA.ndevices(be)
    assert n >= 1
    cur = KA.device(be)
    assert 1 <= cur <= n
    with pytest.raises(KA.ArgumentError):
        KA.device_(be, 0)
    with pytest.raises(KA.ArgumentError):
        KA.device_(be, n + 1)
    KA.device_(be, cur)
    assert KA.device(be) == cur


def test_priority_and_pagelock():
    be = backend()
    for p in ("high", "normal", "low"):
        KA.priority_(be, p)
        A = KA.ones(be, np.float32, 1000)
        KA.synchronize(be)
        assert np.all(KA.Array(A) == 1)
    KA.priority_(be, "normal")
    h = np.arange(4096, dtype=np.float32)
    KA.pagelock_(be, h)
    D = KA.allocate(be, np.float32, 4096)
    KA.copyto_(be, D, h)
    KA.synchronize(be)
    assert np.array_equal(KA.Array(D), h)
    # the registration goes away with the array (a stale one would poison the address range for whatever is allocated there next)
    from b200ka import backend as B
    addr = h.ctypes.data
    assert addr in B._pagelocked
    del h
    import gc
    gc.collect()
    assert addr not in B._pagelocked
    for _ in range(50):          # arrays that reuse the freed pages copy fine
        g = np.arange(4096, dtype=np.float32)
        KA.copyto_(be, D, g)
        KA.synchronize(be)


# ---- no JIT: unknown kernels and bad arguments fail loudly ---------------------------------------------------------------------------
def test_unknown_kernel_and_bad_args():
    be = backend()
    A = KA.allocate(be, np.float32, 16)
    with pytest.raises(KA.ErrorException, match="no precompiled"):
        KA.KernelFunction("my_custom_kernel")(be, 8)(A, ndrange=16)
    with pytest.raises(KA.ErrorException, match="no precompiled"):
        ki_kernel(be, "my_custom_ki_kernel", A, numworkgroups=1, workgroupsize=1)
    with pytest.raises(KA.B200Error):  # ndrange larger than the arrays
        K.copy_kernel_(be, 8)(A, A, ndrange=32)
    with pytest.raises(KA.ErrorException):  # host array passed to a device kernel
        K.copy_kernel_(be, 8)(A, np.zeros(16, dtype=np.float32), ndrange=16)
    with pytest.raises(KA.ArgumentError):  # more than 1024 work-items per group
        K.copy_kernel_(be, 2048)(A, A, ndrange=16)
