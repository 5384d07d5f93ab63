"""CPU-only: the host-side mirror of KA's launch logic against the reference's known answers and the oracle."""
import os

import numpy as np
import pytest

import b200ka as KA
from b200ka.kernel import Kernel, check_launch_args
from b200ka.ndrange import DYNAMIC, StaticSize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- test/test.jl:14-44 "partition" ---------------------------------------------------------------
def test_partition_static_wgs_dynamic_ndrange():
    ws, nd = StaticSize(64), DYNAMIC
    it, dyn = KA.ka_partition(nd, ws, (128,), None)
    assert len(it) == 2 and isinstance(dyn, KA.NoDynamicCheck)
    it, dyn = KA.ka_partition(nd, ws, (129,), None)
    assert len(it) == 3 and isinstance(dyn, KA.DynamicCheck)
    it, dyn = KA.ka_partition(nd, ws, (129,), (64,))
    assert len(it) == 3 and isinstance(dyn, KA.DynamicCheck)
    with pytest.raises(KA.ErrorException):
        KA.ka_partition(nd, ws, (129,), (65,))


def test_partition_static_both():
    ws, nd = StaticSize(64), StaticSize(128)
    it, dyn = KA.ka_partition(nd, ws, (128,), None)
    assert len(it) == 2 and isinstance(dyn, KA.NoDynamicCheck)
    it, dyn = KA.ka_partition(nd, ws, None, None)
    assert len(it) == 2 and isinstance(dyn, KA.NoDynamicCheck)
    with pytest.raises(KA.ErrorException):
        KA.ka_partition(nd, ws, (129,), None)


def test_partition_missing_runtime_sizes():
    with pytest.raises(KA.ErrorException):
        KA.ka_partition(DYNAMIC, DYNAMIC, None, (8,))
    with pytest.raises(KA.ErrorException):
        KA.ka_partition(DYNAMIC, DYNAMIC, (8,), None)


@pytest.mark.parametrize("ndrange,wgs", [((128,), (64,)), ((129,), (64,)), ((16384, 16384), (256,)), ((10, 7), (0, 2)),
                                         ((7, 7, 3), (4, 2)), ((0,), (1,)), ((5, 9), (8, 8))])
def test_nd_partition_matches_oracle(orc, ndrange, wgs):
    blocks, w, dyn = KA.nd_partition(ndrange, wgs)
    ob, ow, od = orc.partition(ndrange, wgs)
    assert blocks == ob and w == ow and isinstance(dyn, KA.DynamicCheck) == od


def test_expand_matches_oracle(orc):
    for blocks, items in [((4, 4), (32, 32)), ((4, 128), (32, 1)), ((128, 4), (1, 32)), ((3, 2, 2), (4, 2, 3))]:
        nd = KA.NDRange(len(blocks), blocks, items)
        ctx = orc.raw_ctx(blocks, items)
        G, L = len(nd), nd.nitems()
        rng = np.random.default_rng(0)
        for g, l in zip(rng.integers(1, G + 1, 200), rng.integers(1, L + 1, 200)):
            assert KA.expand(nd, int(g), int(l)) == orc.expand(ctx, int(g), int(l))


def test_threads_to_workgroupsize(orc):
    for threads, nd in [(1024, (256, 45)), (1024, (4096,)), (1024, (8, 8, 100)), (256, (16384, 16384))]:
        assert KA.threads_to_workgroupsize(threads, nd) == orc.threads_to_workgroupsize(threads, nd)


def test_static_size_indexing():
    s = StaticSize(32, 4)
    assert s[1] == 32 and s[2] == 4 and s[3] == 1 and len(s) == 128 and s.ndims() == 2


def test_launch_config_like_reference():
    be = KA.B200Backend()
    # memcopy_static: copy_kernel!(backend, 32, size(A)) called with ndrange=size(A)
    k = KA.kernels.copy_kernel_(be, 32, (128, 128))
    nd, wgs, it, dyn = k.launch_config((128, 128), None)
    assert nd is None and it.blocks == (4, 128) and it.workitems == (32, 1)
    with pytest.raises(KA.ErrorException):
        k.launch_config((128, 129), None)
    # naive_transpose: static wgs 256, dynamic ndrange
    k = KA.kernels.naive_transpose_kernel_(be, 256)
    nd, wgs, it, dyn = k.launch_config((16384, 16384), None)
    assert it.workitems == (256, 1) and it.blocks == (64, 16384) and len(it) == 2 ** 20
    desc = k.mkcontext(nd, it)
    assert desc.kind == 0 and desc.ndim == 2 and desc.dynamic_check == 1
    assert list(desc.ndrange)[:2] == [16384, 16384] and list(desc.blocks)[:2] == [64, 16384]
    # fully dynamic kernel without ndrange -> error
    with pytest.raises(KA.ErrorException):
        KA.kernels.copy_kernel_(be).launch_config(None, None)


def test_check_launch_args():
    check_launch_args((2, 2, 2), (2, 2, 2))
    with pytest.raises(KA.ArgumentError):
        check_launch_args((2, 2, 2, 2), (2, 2, 2))
    with pytest.raises(KA.ArgumentError):
        check_launch_args((2, 2, 2), (2, 2, 2, 2))


def test_backend_traits():
    be = KA.B200Backend()
    assert be == KA.B200Backend()
    assert KA.supports_float64(be) and KA.supports_atomics(be) and KA.supports_unified(be) in (True, False)
    with pytest.raises(KA.ErrorException):
        KA.priority_(be, "urgent")
    with pytest.raises(KA.MethodError):
        KA.synchronize(object())
    with pytest.raises(KA.ArgumentError):
        KA.get_backend(object())


def test_argconvert_records():
    from b200ka import _lib
    from b200ka.kernel import argconvert

    a = argconvert(KA.Val(16))
    assert a.kind == _lib.B200_ARG_VAL and a.scalar == 16
    a = argconvert(2.5)
    assert a.kind == _lib.B200_ARG_SCALAR and a.fscalar == 2.5 and a.eltype == _lib.B200_T_F64
    a = argconvert(np.float32(2.5))
    assert a.eltype == _lib.B200_T_F32
    with pytest.raises(KA.ErrorException):
        argconvert(np.zeros(3))


# ---- multi-GPU partitioning -------------------------------------------------------------------------------
@pytest.mark.parametrize("ws", [1, 2, 4, 8])
def test_slabs_tile_the_range(ws):
    from b200ka import multigpu as mg

    for n, fn in [(2 ** 26, mg.copy_shard), (16384, mg.transpose_shard), (2 ** 30, mg.histogram_shard),
                  (8192, mg.matmul_shard), (1000003, mg.histogram_shard)]:
        edges = [fn(n, ws, r) for r in range(ws)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
            assert a1 == b0 and a1 > a0


def test_distributed_transpose_block_plan():
    # numpy model of b200_transpose_alltoall: every rank applies its block plan; the slabs then hold a = b^T column-sharded
    from b200ka import multigpu as mg
    for (N0, N1, G) in [(8, 12, 2), (12, 8, 4), (16, 16, 8), (6, 6, 1)]:
        b = np.arange(N0 * N1, dtype=np.float64).reshape(N0, N1)
        rb, cb = N0 // G, N1 // G
        slabs = [np.full((N1, rb), -1.0) for _ in range(G)]                       # a_h = a[:, h-th block of N0/G columns]
        for g in range(G):
            b_loc = b[:, g * cb:(g + 1) * cb]                                      # b_g = b[:, g-th block of N1/G columns]
            for (h, src_row, dst_row) in mg.transpose_alltoall_blocks(N0, N1, G, g):
                slabs[h][dst_row:dst_row + cb, :] = b_loc[src_row:src_row + rb, :].T
        a = np.concatenate(slabs, axis=1)
        assert np.array_equal(a, b.T)
    with pytest.raises(KA.ArgumentError):
        mg.transpose_alltoall_blocks(10, 8, 4, 0)


def test_histogram_slice_cuts_model():
    """Python restatement of `slice_cut` (csrc/kernels/histogram.cu): the int4 units [cut(b), cut(b+1)) of CTA b tile [0, n4),
    and every interior cut falls on a 512-byte line of the input ADDRESS (32 units), whatever the 16-byte alignment of the input."""
    def cut(off, n4, grid, b):
        if b == 0:
            return 0
        if b >= grid:
            return n4
        u = ((n4 * b // grid + off) & ~31) - off
        return max(u, 0)

    rng = np.random.default_rng(7)
    cases = [(0, 2 ** 28, 148), (5, 2 ** 28 - 3, 148), (31, 1000, 148), (17, 33, 148), (0, 0, 148), (9, 1, 1), (3, 10 ** 6 + 7, 7)]
    cases += [(int(rng.integers(0, 32)), int(rng.integers(0, 1 << 20)), int(rng.integers(1, 300))) for _ in range(200)]
    for off, n4, grid in cases:
        cuts = [cut(off, n4, grid, b) for b in range(grid + 1)]
        assert cuts[0] == 0 and cuts[-1] == n4
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        assert all((c + off) % 32 == 0 for c in cuts[1:-1] if c > 0)
        per = n4 / grid                                   # no slice grows by more than one line over its fair share
        assert all(b - a <= per + 32 for a, b in zip(cuts, cuts[1:]))


# ---- strength-reduced NDRange -> grid lowering (csrc/ka_device.cuh, host half) ----------------------------------------------
_MAGIC_HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include "ka_device.cuh"
using namespace b200;
static unsigned umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static unsigned fastdiv(unsigned n, unsigned mul, unsigned shr) { return mul ? (umulhi(n, mul) >> shr) : n; }
int main() {
    // every divisor class: 1, powers of two, primes, 2^k +- 1, the largest legal values; dividends up to 2^31 - 1
    unsigned long long bad = 0, checked = 0;
    unsigned ds[] = {1, 2, 3, 4, 5, 7, 8, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000, 1023, 1024, 1025, 65535, 65536, 65537,
                     1000003, 16777215, 16777216, 16777217, 0x3fffffff, 0x40000000, 0x40000001, 0x7ffffffe, 0x7fffffff};
    unsigned long long x = 88172645463325252ull;
    for (unsigned d : ds) {
        unsigned mul, shr;
        ka_magic(d, &mul, &shr);
        unsigned ns[] = {0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0x7fffffff, 0x7ffffffe, 0x40000000, 0x3fffffff};
        for (unsigned n : ns) {
            if (n > 0x7fffffffu) continue;
            ++checked;
            if (fastdiv(n, mul, shr) != n / d) ++bad;
        }
        for (int k = 0; k < 200000; ++k) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            const unsigned n = (unsigned)(x >> 33);           // < 2^31
            ++checked;
            if (fastdiv(n, mul, shr) != n / d) ++bad;
        }
    }
    // random divisors
    for (int k = 0; k < 300000; ++k) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const unsigned d = (unsigned)((x >> 33) >> (x & 31)) | 1u;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const unsigned n = (unsigned)(x >> 33);
        unsigned mul, shr;
        ka_magic(d, &mul, &shr);
        ++checked;
        if (fastdiv(n, mul, shr) != n / d) ++bad;
    }
    // lowering choice
    KaCtx c = {};
    c.ndim = 2; c.dynamic_check = 1;
    c.ndrange[0] = 16384; c.ndrange[1] = 16384; c.wgs[0] = 256; c.wgs[1] = 1; c.blocks[0] = 64; c.blocks[1] = 16384;
    ka_ctx_finish(&c);
    printf("transpose mode=%d ragged=%d items=%u\n", c.mode, c.ragged, c.items);
    c.ndrange[1] = 70000; c.blocks[1] = 70000;
    ka_ctx_finish(&c);
    printf("wide mode=%d ragged=%d\n", c.mode, c.ragged);
    c.ndim = 4; c.ndrange[0] = 5; c.ndrange[1] = 3; c.ndrange[2] = 2; c.ndrange[3] = 3;
    c.wgs[0] = 2; c.wgs[1] = 2; c.wgs[2] = 1; c.wgs[3] = 2; c.blocks[0] = 3; c.blocks[1] = 2; c.blocks[2] = 2; c.blocks[3] = 2;
    ka_ctx_finish(&c);
    printf("4d mode=%d ragged=%d items=%u\n", c.mode, c.ragged, c.items);
    ka_ctx_finish(&c, KA_MODE_DIV64);
    printf("forced mode=%d\n", c.mode);
    printf("checked=%llu bad=%llu\n", checked, bad);
    return bad != 0;
}
"""


def test_magic_number_division_and_lowering_choice(tmp_path):
    """q = umulhi(n, mul) >> shr equals n / d for every n < 2^31 (the bound ka_ctx_finish guarantees: group and item ids), and
    the host picks the native grid where CUDA's limits allow it -- compiled from the real header with g++ (host half only)."""
    import subprocess

    src = tmp_path / "magic.cpp"
    src.write_text(_MAGIC_HARNESS)
    exe = tmp_path / "magic"
    inc = os.path.join(ROOT, "kernelabstractions.jl_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", inc, "-x", "c++", str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "transpose mode=0 ragged=0 items=256"     # native 2-d grid, nothing ragged: __validindex folds to true
    assert lines[1] == "wide mode=1 ragged=0"                     # 70000 groups in y exceed the grid limit -> multiply-high
    assert lines[2] == "4d mode=1 ragged=11 items=8"              # dims 1, 2 and 4 are ragged (5 % 2, 3 % 2, 3 % 2)
    assert lines[3] == "forced mode=2"
    assert lines[4].endswith("bad=0") and int(lines[4].split()[0].split("=")[1]) > 7_000_000
