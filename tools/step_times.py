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
import time  # noqa: E402
evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
host = [0.0] * (args.steps + 1)
mall = [0] * (args.steps + 1)
evs[0].record()
host[0] = time.perf_counter()
mall[0] = torch.cuda.memory_stats().get("num_device_alloc", 0)
for i in range(args.steps):
    trainer.step({"points": batches[i % 4], "batch_size": 8})
    evs[i + 1].record()
    host[i + 1] = time.perf_counter()
    mall[i + 1] = torch.cuda.memory_stats().get("num_device_alloc", 0)
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
print(f"timers={args.timers} nogc={args.nogc} mean {sum(ts) / len(ts):.2f} ms  steps: " + " ".join(f"{t:.1f}" for t in ts))
med = sorted(ts)[len(ts) // 2]
for i, t_ in enumerate(ts):
    if t_ > 1.08 * med:
        # device time of the step, host time to enqueue it (and the two before it: the host runs up to two steps ahead), cudaMallocs
        print(f"  hiccup step {i}: device {t_:.1f} ms, host enqueue {1e3 * (host[i + 1] - host[i]):.1f} ms (previous "
              f"{1e3 * (host[i] - host[i - 1]) if i else 0:.1f}, {1e3 * (host[i - 1] - host[i - 2]) if i > 1 else 0:.1f}), "
              f"cudaMallocs in step {mall[i + 1] - mall[i]} (previous {mall[i] - mall[i - 1] if i else 0})")
# host time to ENQUEUE one step (queue empty at entry, so nothing inside step() waits for the device) and the device time of
# the same step started from an empty queue: enqueue > device means the step is launch bound whenever the host cannot run ahead
hq, dq = [], []
for i in range(6):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    trainer.step({"points": batches[i % 4], "batch_size": 8})
    b.record()
    hq.append(1e3 * (time.perf_counter() - t0))
    torch.cuda.synchronize()
    dq.append(a.elapsed_time(b))
print("host enqueue per step from an empty queue (ms): " + " ".join(f"{t:.1f}" for t in hq))
print("device span of the same steps (ms):            " + " ".join(f"{t:.1f}" for t in dq))
