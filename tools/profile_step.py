"""Kernel-time breakdown of the MAE pre-train step with torch.profiler (quick look between ncu runs).
usage: python tools/profile_step.py [--dtype bf16] [--steps 2] > gpurun_out/step_profile.txt"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import config  # noqa: E402
from gd_mae_b200.trainer import MAETrainer  # noqa: E402
from oracle import gdmae_oracle as O  # noqa: E402  (synthetic scene generator only)

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
cfg = config.builtin_cfg("waymo_ssl")
model = config.build_mae_model(cfg).cuda()
trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=100)
pts = torch.from_numpy(O.synth_batch(list(range(args.batch)), O.make_cfg("waymo_ssl"))).cuda()
import contextlib  # noqa: E402
ac = contextlib.nullcontext()
config.set_precision(model, args.dtype, dense_spatial_features=args.dtype == "fp32")
for _ in range(4):
    with ac:
        trainer.step({"points": pts, "batch_size": args.batch})
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        with ac:
            trainer.step({"points": pts, "batch_size": args.batch})
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=70))
