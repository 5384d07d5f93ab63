"""Child process of test_gpu_parity.test_matmul_bf16_tensor_core (and of bench.py's opt-in BF16 line).

Checks the tcgen05/TMEM BF16 GEMM against the oracle: FP32-oracle GEMM (reference summation structure) on
BF16-rounded inputs, rtol 1e-2 as north_star states; additionally reports the distance to FP64 truth on the
rounded inputs, which for FP32 accumulation should be ~1e-6.  Exit code 0 == pass.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kernelabstractions.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import b200ka as KA  # noqa: E402
import ka_oracle as orc  # noqa: E402
from b200ka import _lib  # noqa: E402

be = KA.B200Backend()
ok = True
for (M, K, N) in [(128, 64, 256), (128, 256, 256), (256, 512, 512), (1024, 2048, 1024), (384, 4096, 768)]:
    A = orc.gen_uniform_f32((M, K), 1237) - 0.5
    B = orc.gen_uniform_f32((K, N), 1238) - 0.5
    dA, dB = KA.to_device(A), KA.to_device(B)
    dC = KA.zeros(be, np.float32, M, N)
    ws = KA.allocate(be, np.uint16, M * K + K * N)
    _lib.call("b200_matmul_f32_via_bf16", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, M, K, M,
              C.c_void_p(ws.ptr))
    KA.synchronize(be)
    got = KA.Array(dC)
    # operand conversion kernels: bit-exact against round-to-nearest-even
    wsh = KA.Array(ws)
    a_bits = wsh[:M * K].reshape(M, K)            # M rows of K (row-major)
    b_bits = wsh[M * K:].reshape(N, K)            # N rows of K == column-major K x N
    conv_ok = np.array_equal(a_bits, orc.to_bf16_bits(A)) and np.array_equal(b_bits, orc.to_bf16_bits(B).T)
    Ar = np.asfortranarray(orc.round_bf16(A))
    Br = np.asfortranarray(orc.round_bf16(B))
    ref = np.zeros((M, N), dtype=np.float32, order="F")
    orc.coalesced_matmul(ref, Ar, Br)
    truth = orc.matmul_f64_truth(Ar, Br)
    scale = np.abs(truth).max()
    err_oracle = np.abs(got - ref).max() / scale
    err_truth = np.abs(got - truth).max() / scale
    close = np.allclose(got, ref, rtol=1e-2, atol=1e-2 * scale * 1e-2)
    print(f"bf16 gemm {M}x{K}x{N}: convert_bit_exact={conv_ok} max_rel_err_vs_oracle={err_oracle:.3e} "
          f"vs_fp64_truth={err_truth:.3e} allclose_rtol1e-2={close}", flush=True)
    ok = ok and conv_ok and close and err_truth < 1e-4


def tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_rn(x):
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


# ---- TF32 path: FP32 operands straight from HBM (A is MN-major for the tensor core), incl. ragged edge tiles ----
for (M, K, N) in [(128, 32, 256), (128, 256, 256), (256, 512, 512), (1024, 2048, 1024), (384, 4096, 768), (200, 100, 300), (129, 40, 257)]:
    A = orc.gen_uniform_f32((M, K), 1237) - 0.5
    B = orc.gen_uniform_f32((K, N), 1238) - 0.5
    lda, ldb = (M + 3) // 4 * 4, (K + 3) // 4 * 4           # TMA needs leading dimensions that are multiples of 4
    Ap = np.zeros((lda, K), dtype=np.float32, order="F")
    Ap[:M] = A
    Bp = np.zeros((ldb, N), dtype=np.float32, order="F")
    Bp[:K] = B
    dA, dB = KA.to_device(Ap), KA.to_device(Bp)
    dC = KA.allocate(be, np.float32, M, N).fill_(-7.0)
    _lib.call("b200_matmul_tf32", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, lda, ldb, M)
    KA.synchronize(be)
    got = KA.Array(dC)
    errs = {}
    for name, f in (("trunc", tf32_trunc), ("rn", tf32_rn)):
        truth = orc.matmul_f64_truth(np.asfortranarray(f(A)), np.asfortranarray(f(B)))
        errs[name] = np.abs(got - truth).max() / np.abs(truth).max()
    Ar, Br = np.asfortranarray(tf32_rn(A)), np.asfortranarray(tf32_rn(B))
    ref = np.zeros((M, N), dtype=np.float32, order="F")
    orc.coalesced_matmul(ref, Ar, Br)
    scale = np.abs(ref).max()
    close = np.allclose(got, ref, rtol=1e-2, atol=1e-2 * scale)
    print(f"tf32 gemm {M}x{K}x{N}: max_rel_err vs fp64 truth on truncated inputs={errs['trunc']:.3e}, on RN-rounded inputs={errs['rn']:.3e} "
          f"allclose_rtol1e-2_vs_fp32_oracle_on_tf32_inputs={close}", flush=True)
    ok = ok and close and min(errs.values()) < 2e-3

# ---- BF16 path on ragged shapes (edge tiles: TMA zero fill + predicated stores) ----
for (M, K, N) in [(200, 104, 300), (129, 72, 257)]:
    A = orc.gen_uniform_f32((M, K), 1237) - 0.5
    B = orc.gen_uniform_f32((K, N), 1238) - 0.5
    dA, dB = KA.to_device(A), KA.to_device(B)
    dC = KA.allocate(be, np.float32, M, N).fill_(-7.0)
    ws = KA.allocate(be, np.uint16, M * K + K * N)
    _lib.call("b200_matmul_f32_via_bf16", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, M, K, M, C.c_void_p(ws.ptr))
    KA.synchronize(be)
    got = KA.Array(dC)
    truth = orc.matmul_f64_truth(np.asfortranarray(orc.round_bf16(A)), np.asfortranarray(orc.round_bf16(B)))
    err = np.abs(got - truth).max() / np.abs(truth).max()
    print(f"bf16 gemm ragged {M}x{K}x{N}: max_rel_err_vs_fp64_truth={err:.3e}", flush=True)
    ok = ok and err < 1e-4

# ---- 3xTF32: FP32-grade accuracy on the tensor cores (split operands, chunked accumulation) ----
for (M, K, N) in [(128, 32, 256), (256, 512, 512), (200, 100, 300), (129, 40, 257), (512, 8192, 512), (1024, 4100, 768)]:
    A = orc.gen_uniform_f32((M, K), 1237)            # uniform [0,1) like rand(Float32, ...) in examples/performant_matmul.jl:82-83
    B = orc.gen_uniform_f32((K, N), 1238)
    lda, ldb = (M + 3) // 4 * 4, (K + 3) // 4 * 4
    Ap = np.zeros((lda, K), dtype=np.float32, order="F")
    Ap[:M] = A
    Bp = np.zeros((ldb, N), dtype=np.float32, order="F")
    Bp[:K] = B
    dA, dB = KA.to_device(Ap), KA.to_device(Bp)
    dC = KA.allocate(be, np.float32, M, N).fill_(-7.0)
    ws = KA.allocate(be, np.float32, lda * K + ldb * N)
    _lib.call("b200_matmul_f32_tf32x3", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, lda, ldb, M, C.c_void_p(ws.ptr))
    KA.synchronize(be)
    got = KA.Array(dC)
    truth = orc.matmul_f64_truth(np.asfortranarray(A), np.asfortranarray(B))
    rel = np.abs(got - truth) / np.abs(truth)
    As, Bs = A - 0.5, B - 0.5                          # signed inputs: norm-wise error (cancellation makes element-wise rtol meaningless)
    KA.copyto_(be, dA, np.asfortranarray(np.vstack([As, np.zeros((lda - M, K), np.float32)])))
    KA.copyto_(be, dB, np.asfortranarray(np.vstack([Bs, np.zeros((ldb - K, N), np.float32)])))
    _lib.call("b200_matmul_f32_tf32x3", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, lda, ldb, M, C.c_void_p(ws.ptr))
    KA.synchronize(be)
    ts = orc.matmul_f64_truth(np.asfortranarray(As), np.asfortranarray(Bs))
    nrm = np.linalg.norm(KA.Array(dC) - ts) / np.linalg.norm(ts)
    print(f"tf32x3 gemm {M}x{K}x{N}: max element-wise rel err vs fp64 = {rel.max():.3e} (rtol 1e-5 bar), signed inputs norm-wise = {nrm:.3e}", flush=True)
    ok = ok and rel.max() < 1e-5 and nrm < 1e-5

# ---- every tile width the wave-filling heuristic may choose (tcgemm.bn pinned), both kinds, ragged N ----
M, K, N = 256, 256, 1000
A = orc.gen_uniform_f32((M, K), 1237) - 0.5
B = orc.gen_uniform_f32((K, N), 1238) - 0.5
dA, dB = KA.to_device(A), KA.to_device(B)
ws = KA.allocate(be, np.uint16, M * K + K * N)
truth_bf = orc.matmul_f64_truth(np.asfortranarray(orc.round_bf16(A)), np.asfortranarray(orc.round_bf16(B)))
truth_tf = orc.matmul_f64_truth(np.asfortranarray(tf32_trunc(A)), np.asfortranarray(tf32_trunc(B)))
worst = {"bf16": 0.0, "tf32": 0.0}
for bn in range(128, 257, 16):
    _lib.call("b200_tune_set", b"tcgemm.bn", bn)
    dC = KA.allocate(be, np.float32, M, N).fill_(-7.0)
    _lib.call("b200_matmul_f32_via_bf16", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, M, K, M, C.c_void_p(ws.ptr))
    KA.synchronize(be)
    worst["bf16"] = max(worst["bf16"], np.abs(KA.Array(dC) - truth_bf).max() / np.abs(truth_bf).max())
    dC.fill_(-7.0)
    _lib.call("b200_matmul_tf32", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, M, K, M)
    KA.synchronize(be)
    worst["tf32"] = max(worst["tf32"], np.abs(KA.Array(dC) - truth_tf).max() / np.abs(truth_tf).max())
# 3xTF32 with an Inf in A: the affected row of C is non-finite (Inf, or NaN where the Inf meets a zero remainder of B in a
# correction term -- documented deviation from plain FP32), every other row finite
Ainf = A.copy()
Ainf[5, 7] = np.inf
dAi = KA.to_device(np.asfortranarray(Ainf))
dCi = KA.allocate(be, np.float32, M, N).fill_(-7.0)
ws_i = KA.allocate(be, np.float32, M * K + K * N)
Bpos = np.abs(B) + 0.25
dBp = KA.to_device(np.asfortranarray(Bpos))
_lib.call("b200_matmul_f32_tf32x3", C.c_void_p(dCi.ptr), C.c_void_p(dAi.ptr), C.c_void_p(dBp.ptr), M, K, N, M, K, M, C.c_void_p(ws_i.ptr))
KA.synchronize(be)
gi = KA.Array(dCi)
inf_ok = bool(not np.any(np.isfinite(gi[5, :])) and np.all(np.isfinite(np.delete(gi, 5, axis=0))))
print(f"tf32x3 with an Inf operand: its row is non-finite and the rest finite = {inf_ok}", flush=True)
ok = ok and inf_ok
worst["tf32x3"] = 0.0
ws3 = KA.allocate(be, np.float32, M * K + K * N)
truth_x3 = orc.matmul_f64_truth(np.asfortranarray(A), np.asfortranarray(B))
x3_ref = None                                         # the tile width must not change a bit of C (same per-element arithmetic)
x3_same = True
for wide, widths in ((1, (64, 96, 128, 160, 224, 256)), (0, (64, 96, 128))):
    _lib.call("b200_tune_set", b"tcgemm.x3_wide", wide)
    x3_ref = None                                     # (the two layouts differ in how often the correction accumulator is read)
    for bn in widths:
        _lib.call("b200_tune_set", b"tcgemm.bn", bn)
        dC = KA.allocate(be, np.float32, M, N).fill_(-7.0)
        _lib.call("b200_matmul_f32_tf32x3", C.c_void_p(dC.ptr), C.c_void_p(dA.ptr), C.c_void_p(dB.ptr), M, K, N, M, K, M, C.c_void_p(ws3.ptr))
        KA.synchronize(be)
        got = KA.Array(dC)
        worst["tf32x3"] = max(worst["tf32x3"], np.abs(got - truth_x3).max() / np.abs(truth_x3).max())
        if x3_ref is None:
            x3_ref = got
        x3_same = x3_same and np.array_equal(got.view(np.uint32), x3_ref.view(np.uint32))
_lib.call("b200_tune_set", b"tcgemm.bn", 0)
_lib.call("b200_tune_set", b"tcgemm.x3_wide", -1)
print(f"tf32x3 tile widths 64..256 (wide layout) and 64..128 (narrow) on {M}x{K}x{N}: worst max_rel_err {worst['tf32x3']:.3e}, "
      f"bit-identical within a layout = {x3_same}", flush=True)
ok = ok and worst["tf32x3"] < 1e-5 and x3_same
print(f"tile widths 128..256 step 16 on {M}x{K}x{N}: worst max_rel_err bf16={worst['bf16']:.3e} tf32={worst['tf32']:.3e}", flush=True)
ok = ok and worst["bf16"] < 1e-4 and worst["tf32"] < 1e-4
sys.exit(0 if ok else 1)
