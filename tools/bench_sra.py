"""Micro-benchmark / tuning sweep of the SRA attention kernels on realistic window tables.

Builds variants of csrc/sra_attention.cu with different -D tunables (bin size, channel slice, CTA
size), then times forward and backward of each on the three pyramid scales of a synthetic
Waymo-shape batch (B=8) with CUDA events, flushing L2 between launches.
  python tools/bench_sra.py > gpurun_out/bench_sra.txt
"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import ops, _lib as L  # noqa: E402
from gd_mae_b200.pcdet.utils.spconv_utils import spconv, plan_pyramid  # noqa: E402
from oracle import gdmae_oracle as O  # noqa: E402  (synthetic scene generator only)

CSRC = os.path.join(ROOT, "gd-mae_b200", "csrc")
OUT = os.path.join(ROOT, "gpurun_out", "variants")
VARIANTS = {  # name: (BIN, SLICE, FWD_THREADS, BWD_THREADS, MIN_CTAS)
    "b64_s64_t256": (64, 64, 256, 128, 2),
}


def build_variant(name, cfg):
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, f"libsra_{name}.so")
    flags = [f"-DSRA_BIN={cfg[0]}", f"-DSRA_SLICE={cfg[1]}", f"-DSRA_FWD_THREADS={cfg[2]}", f"-DSRA_BWD_THREADS={cfg[3]}",
             f"-DSRA_MIN_CTAS={cfg[4]}"]
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-o", so, os.path.join(CSRC, "sra_attention.cu"), os.path.join(CSRC, "api.cu")] + flags
    subprocess.run(cmd, check=True, cwd=CSRC)
    lib = ctypes.CDLL(so)
    lib.gdmae_last_error.restype = ctypes.c_char_p
    return lib


def tables():
    cfg = O.make_cfg("waymo_ssl")
    pts = torch.from_numpy(O.synth_batch(list(range(8)), cfg)).cuda()
    ps = ops.dynamic_voxelize(pts, cfg["pc_range"], cfg["voxel"], cfg["grid"], 8)
    noise = torch.rand(ps.n_pillars, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    mask = ops.random_mask(noise, ps.batch_offsets_dev, 8, 0.85)
    nvis = sum(int((ps.batch_offsets[b + 1] - ps.batch_offsets[b]) * (1 - 0.85)) for b in range(8))
    _, idx, grid, _ = ops.visible_sites(ps.voxel_coords, mask, nvis, 8, 468, 468)
    sp = spconv.SparseConvTensor(None, idx, [468, 468], 8, {"rank_grid": grid})
    plan_pyramid(sp, 2)
    out = [("scale1", 128, sp)]
    d1 = sp.down()
    sp2 = spconv.SparseConvTensor(None, d1.indices, d1.spatial_shape, 8, d1.struct)
    out.append(("scale2", 256, sp2))
    d2 = sp2.down()
    out.append(("scale3", 256, spconv.SparseConvTensor(None, d2.indices, d2.spatial_shape, 8, d2.struct)))
    return [(n, d, s.window_tables()[0]) for n, d, s in out]


def timeit(fn, flush, iters=12):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tabs = tables()
    libs = {n: build_variant(n, c) for n, c in VARIANTS.items()}
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    for name, d, t in tabs:
        N = t.N
        g = torch.Generator("cuda").manual_seed(1)
        qkv = torch.randn(N, 3 * d, device="cuda", generator=g)
        lut = 0.5 * torch.randn(64, 2 * d, device="cuda", generator=g)
        dout = torch.randn(N, d, device="cuda", generator=g)
        tau = torch.ones(1, device="cuda")
        cnt = (t.win_off[1:] - t.win_off[:-1]).float()
        cnt = cnt[cnt > 0]
        fb, bb = N * d * 16 + N * 8, N * d * 32 + N * 8
        print(f"== {name}: N={N} d={d} windows={cnt.numel()} mean={cnt.mean():.1f} max={int(cnt.max())} "
              f"sum n^2={int((cnt * cnt).sum())}  fwd alg bytes {fb / 1e6:.1f} MB -> {fb / 6538.3e3:.1f} us at peak")
        ref = None
        for vn, lib in libs.items():
            out = torch.empty(N, d, device="cuda")
            lse = torch.empty(N, 8, device="cuda")
            dqkv = torch.empty_like(qkv)
            dts = torch.zeros(1, dtype=torch.float64, device="cuda")
            work = torch.empty(N, 8, device="cuda")

            def fwd():
                rc = lib.gdmae_sra_attention_fwd(L.P(qkv), L.P(lut), L.P(t.row_info), L.i64(N), d, 8, L.P(tau), L.f32(0.01), None, 0, L.P(out),
                                                 L.P(lse), st())
                assert rc == 0, lib.gdmae_last_error()

            def bwd():
                rc = lib.gdmae_sra_attention_bwd(L.P(qkv), L.P(lut), L.P(t.row_info), L.i64(N), d, 8, L.P(tau), L.f32(0.01), None, 0, L.P(out),
                                                 L.P(lse), L.P(dout), L.P(dqkv), L.P(dts), L.P(work), st())
                assert rc == 0, lib.gdmae_last_error()

            tf = timeit(fwd, flush)
            tb = timeit(bwd, flush)
            if ref is None:
                ref = (out.clone(), dqkv.clone())
            err = max(float((out - ref[0]).abs().max()), float((dqkv - ref[1]).abs().max()))
            print(f"   {vn:16s} fwd {tf:7.1f} us ({fb / tf / 1e3 / 6538.3:.3f} of peak)   bwd {tb:7.1f} us ({bb / tb / 1e3 / 6538.3:.3f})   "
                  f"max|diff vs first| {err:.2e}")
        # tensor-core forward of the main library: bf16 q/k/v in, bf16 out
        qkv_b = qkv.to(torch.bfloat16)
        out_b = torch.empty(N, d, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(N, 8, device="cuda")
        fn = L.lib().gdmae_sra_attention_fwd_tc

        def fwd_tc():
            L.check(fn(L.P(qkv_b), L.P(lut), L.P(t.row_info), L.P(t.bin_units()), L.i64(N), d, 8, L.P(tau), L.f32(0.01), None, 1, L.P(out_b), L.P(lse), st()),
                    "tc")

        tf = timeit(fwd_tc, flush)
        fbb = N * d * 8 + N * 32
        print(f"   {'tensor-core fwd':16s} fwd {tf:7.1f} us ({fbb / tf / 1e3 / 6538.3:.3f} of peak, bf16 io {fbb / 1e6:.1f} MB)   "
              f"max|diff| {float((out_b.float() - ref[0]).abs().max()):.2e}")


if __name__ == "__main__":
    main()
