"""Small target for Nsight Compute: N training steps of the bench workload, nothing else.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py --steps 3
  ncu --set full --clock-control none --import-source on -k regex:sra_fwd_kernel -s 12 -c 3 -o gpurun_out/prof_sra_fwd python tools/ncu_target.py --steps 2"""
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
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--dtype", default="bf16")
args = ap.parse_args()
cfg = config.builtin_cfg("waymo_ssl")
model = config.build_mae_model(cfg).cuda()
config.set_precision(model, args.dtype, dense_spatial_features=args.dtype == "fp32")
trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=100)
pts = torch.from_numpy(O.synth_batch(list(range(args.batch)), O.make_cfg("waymo_ssl"))).cuda()
for _ in range(args.steps):
    loss = trainer.step({"points": pts, "batch_size": args.batch})
torch.cuda.synchronize()
print("loss", float(loss))
