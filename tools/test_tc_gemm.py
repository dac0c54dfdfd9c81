"""GPU check of csrc/tc_gemm.cu (tcgen05 + TMA GEMM) against torch on every layout / epilogue the step uses.
Run on the B200 box:  timeout 300 python tools/test_tc_gemm.py [tn|quick]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gd_mae_b200  # noqa: F401,E402
from gd_mae_b200 import fused  # noqa: E402

BF = torch.bfloat16
torch.manual_seed(0)
dev = torch.device("cuda")
fails = 0


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def check(name, got, want, tol):
    global fails
    torch.cuda.synchronize()
    e = rel(got, want)
    ok = e < tol and bool(torch.isfinite(got.float()).all())
    fails += 0 if ok else 1
    print(f"{'ok  ' if ok else 'FAIL'} {name}: rel err {e:.3e} (tol {tol:.0e})", flush=True)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def ncu_target():
    """a few launches of each kernel variant on the step's largest shapes (for `ncu -k regex:tcg_gemm`)"""
    M, N, K, dff = 71000, 256, 256, 512
    a = torch.randn(M, K, device=dev).to(BF)
    w = (torch.randn(N, K, device=dev) * 0.1).to(BF)
    w1 = (torch.randn(dff, K, device=dev) * 0.1).to(BF)
    bias, gamma, beta, b1 = torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(dff, device=dev)
    res = torch.randn(M, N, device=dev)
    y32, y16, raw = torch.empty(M, N, device=dev), torch.empty(M, N, dtype=BF, device=dev), torch.empty(M, N, dtype=BF, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    E = fused.TcEpilogue()
    E.mode, E.bias, E.res, E.gamma, E.beta_ln, E.eps = 2, bias.data_ptr(), res.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5
    E.y32, E.y16, E.mean, E.rstd = y32.data_ptr(), y16.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    g, h = torch.empty(M, dff, dtype=BF, device=dev), torch.empty(M, dff, dtype=BF, device=dev)
    E2 = fused.TcEpilogue()
    E2.mode, E2.bias, E2.c2, E2.ldc2 = 1, b1.data_ptr(), g.data_ptr(), dff
    dz = torch.randn(M, N, device=dev)
    wk = (torch.randn(K, N, device=dev) * 0.1).to(BF)
    qkv = torch.empty(M, 3 * N, dtype=BF, device=dev)
    w3 = (torch.randn(3 * N, K, device=dev) * 0.1).to(BF)
    dw = torch.zeros(dff, N, device=dev)
    for _ in range(2):
        fused.tc_gemm(a, w.t(), out=raw, epilogue=E)            # LN-fused
        fused.tc_gemm(a, w1.t(), out=h, epilogue=E2)            # GELU-fused
        fused.tc_gemm(a, wk, out=dz, beta=1.0)                  # fp32 read-modify-write
        fused.tc_gemm(a, w3.t(), out=qkv)                       # plain bf16
        fused.tc_gemm(h.t(), a, out=dw, beta=1.0, split_k=True) # weight gradient, split-K
    torch.cuda.synchronize()


def main(quick):
    if quick == "ncu":
        return ncu_target()
    # ---- TN (both K-major): forward GEMMs
    for (M, N, K) in [(128, 64, 64), (300, 128, 64), (1000, 256, 128), (5273 * 8, 384, 128), (70001, 768, 256), (4096, 256, 512),
                      (20000, 256, 1152)]:
        a = torch.randn(M, K, device=dev).to(BF)
        w = torch.randn(N, K, device=dev).to(BF)
        ref = a.float() @ w.float().t()
        for od in (torch.float32, BF):
            c = fused.tc_gemm(a, w.t(), out_dtype=od)
            check(f"TN {M}x{N}x{K} -> {od}", c.float(), ref, 1e-5 if od == torch.float32 else 8e-3)
        c0 = torch.randn(M, N, device=dev)
        c = fused.tc_gemm(a, w.t(), out=c0.clone(), beta=1.0)
        check(f"TN {M}x{N}x{K} beta=1", c, ref + c0, 1e-5)
    if quick == "tn":
        return
    # ---- NN (A K-major, B stored (K,N) = MN-major): input-gradient GEMMs
    for (M, N, K) in [(300, 128, 64), (1000, 256, 128), (40000, 128, 384), (70001, 512, 256), (9000, 2304, 256)]:
        a = torch.randn(M, K, device=dev).to(BF)
        b = torch.randn(K, N, device=dev).to(BF)
        c = fused.tc_gemm(a, b, out_dtype=BF)
        check(f"NN {M}x{N}x{K}", c.float(), a.float() @ b.float(), 8e-3)
    # ---- TN with strided views (columns of a wider matrix), as the executor passes them
    big = torch.randn(5000, 768, device=dev).to(BF)
    w = torch.randn(256, 256, device=dev).to(BF)
    c = fused.tc_gemm(big[:, 512:768], w.t())
    check("TN strided A (lda 768)", c, big[:, 512:768].float() @ w.float().t(), 1e-5)
    # ---- wgrad: A stored (K,M), B stored (K,N), K = tokens, split-K with fp32 reductions
    for (M, N, K) in [(128, 64, 1000), (256, 128, 5273 * 8), (512, 256, 70001), (256, 256, 46000), (256, 1152, 30000), (768, 256, 33333)]:
        at = torch.randn(K, M, device=dev).to(BF)
        b = torch.randn(K, N, device=dev).to(BF)
        ref = at.float().t() @ b.float()
        c = fused.tc_gemm(at.t(), b, split_k=True)
        check(f"wgrad {M}x{N}x{K} split-K", c, ref, 2e-5)
        c0 = torch.randn(M, N, device=dev)
        c = fused.tc_gemm(at.t(), b, out=c0.clone(), beta=1.0, split_k=True)
        check(f"wgrad {M}x{N}x{K} split-K accumulate", c, ref + c0, 2e-5)
    # sub-matrix of dqkv (columns 2d..3d of an (N,3d) matrix) as the transposed operand
    dq = torch.randn(20000, 768, device=dev).to(BF)
    xg = torch.randn(20000, 256, device=dev).to(BF)
    c = fused.tc_gemm(dq[:, 512:].t(), xg, split_k=True)
    check("wgrad strided A^T", c, dq[:, 512:].float().t() @ xg.float(), 2e-5)
    # ---- epilogue 1: bias + GELU
    for (M, N, K) in [(1000, 256, 128), (70001, 512, 256)]:
        a = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * 0.1).to(BF)
        bias = torch.randn(N, device=dev)
        g = torch.empty(M, N, dtype=BF, device=dev)
        E = fused.TcEpilogue()
        E.mode, E.bias, E.c2, E.ldc2 = 1, bias.data_ptr(), g.data_ptr(), N
        dgl = fused.tc_gemm(a, w.t(), out_dtype=BF, epilogue=E)          # C = gelu'(acc + bias): what the backward keeps
        ref_h = a.float() @ w.float().t()
        vv = (ref_h + bias).requires_grad_(True)
        torch.nn.functional.gelu(vv).sum().backward()
        check(f"gelu epilogue {M}x{N}x{K}: gelu'", dgl.float(), vv.grad, 8e-3)
        check(f"gelu epilogue {M}x{N}x{K}: g", g.float(), torch.nn.functional.gelu(ref_h + bias), 8e-3)
    # ---- epilogue 3: GELU backward (NN: dz (M,K=d) @ W2 (d, N=dff)), h saved by epilogue 1
    for (M, N, K) in [(1000, 256, 128), (70001, 512, 256)]:
        dz = torch.randn(M, K, device=dev).to(BF)
        w2 = (torch.randn(K, N, device=dev) * 0.1).to(BF)
        h = torch.randn(M, N, device=dev).to(BF)                        # the saved derivative (any bf16 matrix here)
        colsum = torch.randn(N, device=dev)
        c0 = colsum.clone()
        E = fused.TcEpilogue()
        E.mode, E.h16, E.ldh, E.colsum = 3, h.data_ptr(), N, colsum.data_ptr()
        dh = fused.tc_gemm(dz, w2, out_dtype=BF, epilogue=E)
        ref = (dz.float() @ w2.float()) * h.float()
        check(f"gelu-bwd epilogue {M}x{N}x{K}: dh", dh.float(), ref, 8e-3)
        check(f"gelu-bwd epilogue {M}x{N}x{K}: colsum", colsum - c0, ref.sum(0), 2e-3)
    # ---- epilogue 2: residual + bias + LayerNorm
    for (M, N, K) in [(1000, 128, 128), (42184, 128, 256), (70001, 256, 256), (46000, 256, 512)]:
        a = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * 0.1).to(BF)
        bias, gamma, beta = torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        y32 = torch.empty(M, N, device=dev)
        y16 = torch.empty(M, N, dtype=BF, device=dev)
        mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
        E = fused.TcEpilogue()
        E.mode, E.bias, E.res, E.gamma, E.beta_ln, E.eps = 2, bias.data_ptr(), res.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5
        E.y32, E.y16, E.mean, E.rstd = y32.data_ptr(), y16.data_ptr(), mean.data_ptr(), rstd.data_ptr()
        raw = fused.tc_gemm(a, w.t(), out=torch.empty(M, N, dtype=BF, device=dev), epilogue=E)
        check(f"LN epilogue {M}x{N}x{K}: raw output", raw.float(), a.float() @ w.float().t(), 8e-3)
        z = a.float() @ w.float().t() + bias + res
        ref = torch.nn.functional.layer_norm(z, (N,), gamma, beta, 1e-5)
        check(f"LN epilogue {M}x{N}x{K}: y32", y32, ref, 2e-5)
        check(f"LN epilogue {M}x{N}x{K}: y16", y16.float(), ref, 8e-3)
        check(f"LN epilogue {M}x{N}x{K}: mean", mean, z.mean(1), 2e-5)
        check(f"LN epilogue {M}x{N}x{K}: rstd", rstd, torch.rsqrt(z.var(1, unbiased=False) + 1e-5), 2e-5)
    # ---- timing against cuBLASLt (gdmae_gemm) on the step's shapes
    if not quick:
        for (M, N, K) in [(42184, 384, 128), (71000, 768, 256), (71000, 512, 256), (71000, 256, 512), (71000, 256, 2304)]:
            a = torch.randn(M, K, device=dev).to(BF)
            w = torch.randn(N, K, device=dev).to(BF)
            out = torch.empty(M, N, dtype=BF, device=dev)
            t_own = timeit(lambda: fused.tc_gemm(a, w.t(), out=out))
            t_lt = timeit(lambda: fused.gemm(a, w.t(), out=out))
            byts = (M * K + N * K + M * N) * 2
            print(f"time TN {M}x{N}x{K} bf16 out: own {t_own:.1f} us ({byts / t_own / 1e3:.0f} GB/s, {2 * M * N * K / t_own / 1e6:.0f} TF/s)  "
                  f"cuBLASLt {t_lt:.1f} us", flush=True)
        for (M, N, K, dff) in [(71000, 256, 256, 512), (42184, 128, 128, 256)]:
            a = torch.randn(M, K, device=dev).to(BF)
            w = (torch.randn(N, K, device=dev) * 0.1).to(BF)
            w1 = (torch.randn(dff, K, device=dev) * 0.1).to(BF)
            bias, gamma, beta, b1 = torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(dff, device=dev)
            res = torch.randn(M, N, device=dev)
            y32, y16, raw = torch.empty(M, N, device=dev), torch.empty(M, N, dtype=BF, device=dev), torch.empty(M, N, dtype=BF, device=dev)
            mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
            E = fused.TcEpilogue()
            E.mode, E.bias, E.res, E.gamma, E.beta_ln, E.eps = 2, bias.data_ptr(), res.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5
            E.y32, E.y16, E.mean, E.rstd = y32.data_ptr(), y16.data_ptr(), mean.data_ptr(), rstd.data_ptr()
            t = timeit(lambda: fused.tc_gemm(a, w.t(), out=raw, epilogue=E))
            byts = M * K * 2 + M * N * (4 + 4 + 2 + 2)
            print(f"time LN-fused {M}x{N}x{K}: {t:.1f} us ({byts / t / 1e3:.0f} GB/s)", flush=True)
            g, h = torch.empty(M, dff, dtype=BF, device=dev), torch.empty(M, dff, dtype=BF, device=dev)
            E2 = fused.TcEpilogue()
            E2.mode, E2.bias, E2.c2, E2.ldc2 = 1, b1.data_ptr(), g.data_ptr(), dff
            t = timeit(lambda: fused.tc_gemm(a, w1.t(), out=h, epilogue=E2))
            byts = M * K * 2 + M * dff * 4
            print(f"time GELU-fused {M}x{dff}x{K}: {t:.1f} us ({byts / t / 1e3:.0f} GB/s)", flush=True)
            dz = torch.randn(M, N, device=dev)
            wk = (torch.randn(K, N, device=dev) * 0.1).to(BF)
            t = timeit(lambda: fused.tc_gemm(a, wk, out=dz, beta=1.0))
            print(f"time NN beta=1 fp32 RMW {M}x{N}x{K}: {t:.1f} us ({(M * K * 2 + M * N * 8) / t / 1e3:.0f} GB/s)", flush=True)
        for (M, N, K) in [(512, 256, 71000), (256, 256, 71000), (256, 2304, 71000)]:
            at = torch.randn(K, M, device=dev).to(BF)
            b = torch.randn(K, N, device=dev).to(BF)
            out = torch.zeros(M, N, device=dev)
            t_own = timeit(lambda: fused.tc_gemm(at.t(), b, out=out, beta=1.0, split_k=True))
            t_lt = timeit(lambda: fused.gemm(at.t(), b, out=out, beta=1.0))
            print(f"time wgrad {M}x{N}x{K}: own {t_own:.1f} us ({(M + N) * K * 2 / t_own / 1e3:.0f} GB/s)  cuBLASLt {t_lt:.1f} us", flush=True)


if __name__ == "__main__":
    try:
        main(sys.argv[1] if len(sys.argv) > 1 else "")
    finally:
        try:
            n = ctypes.c_int(0)
            gd_mae_b200._lib.lib().gdmae_tc_gemm_timeouts(ctypes.byref(n))
            print("timeouts", n.value, "fails", fails)
        except Exception as e:  # a trapped launch leaves a sticky error
            print("timeouts unreadable:", e, "fails", fails)
    sys.exit(1 if fails else 0)
