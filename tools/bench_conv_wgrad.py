"""Decoder conv weight gradient (gdmae_conv3x3_wgrad) at the Waymo map size against cuDNN's wgrad: correctness on a small
map and timing (CUDA events, L2 flushed between launches) at B=8, 468 x 468.
  GDMAE_WGRAD_SHIFT=0|1 python tools/bench_conv_wgrad.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import _lib as L  # noqa: E402


def wgrad(dy, x, dw, acc=0):
    B, Y, X, _ = x.shape
    L.check(L.lib().gdmae_conv3x3_wgrad(L.P(dy), L.P(x), B, Y, X, 384, 128, L.P(dw), acc, L.stream()), "gdmae_conv3x3_wgrad")


def main():
    g = torch.Generator("cuda").manual_seed(0)
    for B, Y, X in ((2, 37, 70), (1, 9, 130)):
        x = torch.randn(B, Y, X, 384, device="cuda", generator=g).to(torch.bfloat16)
        dy = torch.randn(B, Y, X, 128, device="cuda", generator=g).to(torch.bfloat16)
        dw = torch.empty(128, 3, 3, 384, device="cuda")
        wgrad(dy, x, dw)
        ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (128, 384, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
        err = float((dw.permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max())
        print(f"shift={os.environ.get('GDMAE_WGRAD_SHIFT', '1')} B={B} Y={Y} X={X}: rel err {err:.2e}", flush=True)
    B, Y, X = 8, 468, 468
    x = torch.randn(B, Y, X, 384, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(B, Y, X, 128, device="cuda", generator=g).to(torch.bfloat16)
    dw = torch.empty(128, 3, 3, 384, device="cuda")
    w = torch.randn(128, 384, 3, 3, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timeit(fn, n=6):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return sorted(ts)[len(ts) // 2]

    t_own = timeit(lambda: wgrad(dy, x, dw))
    t_lib = timeit(lambda: torch.ops.aten.convolution_backward(dy.permute(0, 3, 1, 2), x.permute(0, 3, 1, 2), w, None, [1, 1], [1, 1], [1, 1],
                                                                False, [0, 0], 1, [False, True, False]))
    flops = 2.0 * B * Y * X * 9 * 384 * 128
    print(f"B={B} {Y}x{X}: own wgrad {t_own:.0f} us ({flops / t_own / 1e6:.0f} TFLOP/s), cuDNN wgrad {t_lib:.0f} us ({flops / t_lib / 1e6:.0f} TFLOP/s)")
    to = __import__("ctypes").c_int(0)
    L.lib().gdmae_tc_gemm_timeouts(__import__("ctypes").byref(to))
    print("timeouts", to.value)


if __name__ == "__main__":
    main()
