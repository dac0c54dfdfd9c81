"""One forward and one backward launch of the window-major SRA kernels per pyramid scale (after a warm-up launch) - the
target of `ncu --set full -k regex:sra_(fwd|bwd)_mma`:
  ncu --set full --clock-control none --import-source on -k regex:sra_.*_mma -o gpurun_out/r2_sra python tools/ncu_sra.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import ops  # noqa: E402
from bench_sra import tables  # noqa: E402


def main():
    only = sys.argv[1:] or ["scale2"]
    for name, d, t in tables():
        if name not in only:
            continue
        N = t.N
        g = torch.Generator("cuda").manual_seed(1)
        qkv = torch.randn(N, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
        lut = 0.5 * torch.randn(64, 2 * d, device="cuda", generator=g)
        dout = torch.randn(N, d, device="cuda", generator=g).to(torch.bfloat16)
        tau = torch.ones(1, device="cuda")
        for _ in range(2):
            o, lse = ops.sra_fwd(qkv, lut, tau, t, 0.01, 8, out_dtype=torch.bfloat16)
            ops.sra_bwd(qkv, lut, tau, t, 0.01, 8, None, lse, dout)
        torch.cuda.synchronize()
        print(name, N, d, float(o.float().abs().mean()))


if __name__ == "__main__":
    main()
