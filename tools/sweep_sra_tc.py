"""Tuning sweep of the tensor-core SRA kernels (csrc/sra_attention_tc.cu): builds the file with different -D tunables,
times forward and backward on the three pyramid scales of a synthetic Waymo-shape batch (B=8) with CUDA events (L2
flushed between launches) and checks every variant against the fp32 SIMT kernels of the main library.
  python tools/sweep_sra_tc.py "MM_STAGER_WARPS=5" "MM_STAGER_WARPS=6" ... > gpurun_out/sweep_sra_tc.txt"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import ops, _lib as L  # noqa: E402
from bench_sra import tables, timeit  # noqa: E402

CSRC = os.path.join(ROOT, "gd-mae_b200", "csrc")
OUT = os.path.join(ROOT, "gpurun_out", "variants")


def build_variant(i, defs):
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, f"libsra_tc_{i}.so")
    flags = [f"-D{x}" for x in defs.split() if x]
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-o", so, os.path.join(CSRC, "sra_attention_tc.cu"), os.path.join(CSRC, "tc_gemm.cu"), os.path.join(CSRC, "api.cu")] + flags
    subprocess.run(cmd, check=True, cwd=CSRC)
    lib = ctypes.CDLL(so)
    lib.gdmae_last_error.restype = ctypes.c_char_p
    return lib


def main():
    variants = sys.argv[1:] or [""]
    libs = [(v or "default", build_variant(i, v)) for i, v in enumerate(variants)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    for name, d, t in tables():
        N = t.N
        g = torch.Generator("cuda").manual_seed(1)
        qkv = torch.randn(N, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
        lut = 0.5 * torch.randn(64, 2 * d, device="cuda", generator=g)
        dout = torch.randn(N, d, device="cuda", generator=g).to(torch.bfloat16)
        tau = torch.ones(1, device="cuda")
        # fp32 SIMT reference on the same (bf16-rounded) inputs
        o_ref, lse_ref = ops.sra_fwd(qkv.float(), lut, tau, t, 0.01, 8)
        dq_ref, _ = ops.sra_bwd(qkv.float(), lut, tau, t, 0.01, 8, o_ref, lse_ref, dout.float())
        fb, bb = N * d * 8 + N * 32, N * d * 14 + N * 32
        print(f"== {name}: N={N} d={d}  fwd alg bytes {fb / 1e6:.1f} MB, bwd {bb / 1e6:.1f} MB")
        nbu = L.lib().gdmae_sra_bin_units_bytes(L.i64(N))
        for vn, lib in libs:
            units = torch.empty(nbu // 4, dtype=torch.int32, device="cuda")
            assert lib.gdmae_sra_bin_units(L.P(t.row_info), L.i64(N), L.P(units), st()) == 0
            out = torch.empty(N, d, device="cuda", dtype=torch.bfloat16)
            lse = torch.empty(N, 8, device="cuda")
            lrr = torch.empty(N, 24, device="cuda")
            qkvdw = torch.empty(4 * N * d, device="cuda", dtype=torch.bfloat16)
            dqkv = torch.empty(N, 3 * d, device="cuda", dtype=torch.bfloat16)
            dts = torch.zeros(1, dtype=torch.float64, device="cuda")
            # the window-major operands (in the fused layer the GEMM epilogues write them), then the kernels alone are timed
            rc = lib.gdmae_sra_relayout(L.P(qkv), L.P(lut), L.P(t.row_info), L.P(tau), L.f32(0.01), L.i64(N), d, L.P(dout), L.P(lse_ref),
                                        L.P(qkvdw), L.P(lrr), st())
            assert rc == 0, lib.gdmae_last_error()

            def fwd():
                rc = lib.gdmae_sra_fwd_win(L.P(qkvdw), L.P(units), L.i64(N), d, None, 1, L.P(out), L.P(lrr if FUSED else lse), FUSED, st())
                assert rc == 0, lib.gdmae_last_error()

            def bwd():
                rc = lib.gdmae_sra_bwd_win(L.P(qkvdw), L.P(lrr), L.P(units), L.i64(N), d, L.P(tau), L.f32(0.01), L.P(dqkv), L.P(dts), st())
                assert rc == 0, lib.gdmae_last_error()

            FUSED = 0
            fwd()                                                  # lse by token, for the check below
            FUSED = 1                                              # timed as the encoder layer calls it: lse into the row records
            tf = timeit(fwd, flush)
            tb = timeit(bwd, flush)
            ef = float((out.float() - o_ref).abs().max())
            el = float((lse - lse_ref).abs().max())
            to = ctypes.c_int(0)
            lib.gdmae_sra_wait_timeouts(ctypes.byref(to))
            assert to.value == 0, "bounded wait ran out"
            eb = float((dqkv.float() - dq_ref).abs().max()) / max(float(dq_ref.abs().max()), 1e-12)
            if hasattr(lib, "gdmae_sra_prof_read"):
                buf = (ctypes.c_ulonglong * 16)()
                lib.gdmae_sra_prof_read(buf, 1)
                fwd()
                torch.cuda.synchronize()
                lib.gdmae_sra_prof_read(buf, 1)
                p = list(buf)
                nb = max(p[5], 1)
                print(f"      profile (cycles per bin, producer / first and last math warp; {p[5]} bin iterations): producer wait_empty "
                      f"{p[0] / nb:.0f} issue {p[1] / nb:.0f} | math0 wait_full "
                      f"{p[6] / nb:.0f} entries {p[7] / nb:.0f} ({p[8] / nb:.2f} of {p[9] / nb:.1f} per bin) | math-last wait_full {p[10] / nb:.0f} "
                      f"entries {p[11] / nb:.0f} ({p[12] / nb:.2f}) | CTA total avg {p[13] / 148:.0f} max {p[14]} prologue avg {p[15] / 148:.0f} cycles")
            print(f"   {vn:28s} fwd {tf:7.1f} us ({fb / tf / 1e3 / 6538.3:.3f} of peak)   bwd {tb:7.1f} us ({bb / tb / 1e3 / 6538.3:.3f})   "
                  f"fwd max|diff| {ef:.2e} lse {el:.2e}  bwd rel {eb:.2e}")


if __name__ == "__main__":
    main()
