"""Per-step device times of the bench workload (CUDA events around every step), with the bench's kernel timers on or off.
usage: python tools/step_times.py [--timers 1] [--steps 30]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import _lib, config  # noqa: E402
from gd_mae_b200.trainer import MAETrainer  # noqa: E402
from oracle import gdmae_oracle as O  # noqa: E402  (synthetic scene generator only)

ap = argparse.ArgumentParser()
ap.add_argument("--timers", type=int, default=0)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--nogc", type=int, default=0)
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
cfg = config.builtin_cfg("waymo_ssl")
model = config.build_mae_model(cfg).cuda()
config.set_precision(model, "bf16", dense_spatial_features=False)
trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=200)
ocfg = O.make_cfg("waymo_ssl")
batches = [torch.from_numpy(O.synth_batch([8 * k + i for i in range(8)], ocfg)).cuda() for k in range(4)]
for i in range(5):
    trainer.step({"points": batches[i % 4], "batch_size": 8})
torch.cuda.synchronize()
if args.nogc:
    import gc
    gc.collect()
    gc.disable()
if args.timers:
    _lib.KERNEL_TIMERS = {}
    _lib.lib().gdmae_timing_enable(1)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
evs[0].record()
for i in range(args.steps):
    trainer.step({"points": batches[i % 4], "batch_size": 8})
    evs[i + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
print(f"timers={args.timers} nogc={args.nogc} mean {sum(ts) / len(ts):.2f} ms  steps: " + " ".join(f"{t:.1f}" for t in ts))
