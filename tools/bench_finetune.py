"""BASELINE config 4 (Waymo gd_mae_iou finetune: DynVFE + SPTBackbone + SSTBEVBackbone + CenterHead, no masking) on one
B200: frames/s of the full training step (forward, CenterHead targets + losses, backward, clip + adam_onecycle) on synthetic
Waymo-shape frames with 64 random ground-truth boxes each.
  python tools/bench_finetune.py [--batch 4] [--steps 10] [--dtype bf16] > gpurun_out/finetune_bench.json
One JSON line; device-timed with CUDA events, inputs resident in HBM.  The dense convolutions of the BEV backbone / head are
cuDNN (library) - this is a measurement of the finetune path's own kernels working at 7x the MAE token count (35 k / 25.6 k /
11 k tokens per frame, windows of all three drop levels)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import config  # noqa: E402
from gd_mae_b200.trainer import MAETrainer  # noqa: E402
from oracle import gdmae_oracle as O  # noqa: E402  (synthetic scene generator only)


def random_gt_boxes(seed, n=64):
    r = np.random.RandomState(seed)
    cls = r.randint(1, 4, n)
    dims = np.array([[4.7, 2.1, 1.7], [0.9, 0.9, 1.7], [1.8, 0.8, 1.7]])[cls - 1] * r.uniform(0.8, 1.2, (n, 3))
    c = np.concatenate([r.uniform(-74, 74, (n, 2)), r.uniform(-1, 2, (n, 1))], 1)
    return np.concatenate([c, dims, r.uniform(-np.pi, np.pi, (n, 1)), cls[:, None]], 1).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    cfg = config.builtin_cfg("waymo_iou")
    torch.manual_seed(0)
    model = config.build_mae_model(cfg).cuda()
    config.set_precision(model, a.dtype)
    trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=100)
    ocfg = O.make_cfg("waymo_ssl")
    batches = []
    for k in range(2):
        pts = torch.from_numpy(O.synth_batch([100 * k + i for i in range(a.batch)], ocfg)).cuda()
        gt = torch.from_numpy(np.stack([random_gt_boxes(100 * k + i) for i in range(a.batch)], 0)).cuda()
        batches.append((pts, gt))

    def step(i):
        pts, gt = batches[i % 2]
        return trainer.step({"points": pts, "batch_size": a.batch, "gt_boxes": gt})

    for i in range(a.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        loss = step(a.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    toks = None
    print(json.dumps({"metric": "finetune_frames_per_sec", "workload": "waymo_gd_mae_iou_finetune_synthetic_160k_pt", "value": a.batch / (ms * 1e-3),
                      "unit": "frames/s", "ms_per_step": ms, "frames_per_step": a.batch, "dtype": a.dtype, "steps": a.steps,
                      "warmup": a.warmup, "final_loss": float(loss), "params": trainer.n_params,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
