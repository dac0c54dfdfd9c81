"""Roofline check of the input-side kernel (csrc/augment.cu) at the bench workload's size: B=8 Waymo-shape frames,
~1.27 M points of 6 floats.  CUDA events, L2 flushed between launches.  usage: python tools/bench_augment.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import ops  # noqa: E402
from oracle import gdmae_oracle as O  # noqa: E402  (synthetic scene generator + checker only)

cfg = O.make_cfg("waymo_ssl")
batch = O.synth_batch(list(range(8)), cfg)
pts = torch.from_numpy(batch).cuda()
N = pts.shape[0]
np.random.seed(0)
rows, src, off = [], [], 0
counts = np.bincount(batch[:, 0].astype(np.int64), minlength=8)
prm = []
for b in range(8):
    p = O.draw_world_aug_params(n_points=int(counts[b]))
    prm.append(p)
    rows.append([float(p["flip_x"]), float(p["flip_y"]), np.float32(np.cos(p["rotation"])), np.float32(np.sin(p["rotation"])),
                 np.float32(p["scaling"]), 0.0])
    src.append(p["perm"] + off)
    off += int(counts[b])
params = torch.tensor(np.asarray(rows, dtype=np.float32)).cuda()
src_index = torch.from_numpy(np.concatenate(src).astype(np.int32)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=20):
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


# parity at full size against the oracle, frame 0
out = ops.world_augment(pts, params, src_index).cpu().numpy()
f0 = batch[batch[:, 0] == 0][:, 1:]
ref = O.world_augment(f0, prm[0]["flip_x"], prm[0]["flip_y"], prm[0]["rotation"], prm[0]["scaling"], prm[0]["perm"])
err = np.abs(out[:ref.shape[0], 1:4] - ref[:, :3]).max() / np.abs(ref[:, :3]).max()
assert err <= 1e-6 and np.array_equal(out[:ref.shape[0], 4:], ref[:, 3:]), err
alg = N * pts.shape[1] * 4 * 2
t_plain = timeit(lambda: ops.world_augment(pts, params, None))
t_shuf = timeit(lambda: ops.world_augment(pts, params, src_index))
print(f"N={N} rows of {pts.shape[1]} floats; algorithmic bytes {alg / 1e6:.1f} MB (+{N * 4 / 1e6:.1f} MB index with shuffle); parity vs oracle (frame 0) rel {err:.1e}")
print(f"augment only : {t_plain:7.1f} us  {alg / t_plain / 1e3:7.0f} GB/s  {alg / t_plain / 1e3 / peak:.2f} of {peak:.0f} GB/s")
print(f"with shuffle : {t_shuf:7.1f} us  {(alg + 4 * N) / t_shuf / 1e3:7.0f} GB/s  {(alg + 4 * N) / t_shuf / 1e3 / peak:.2f} of peak (random 24-byte row gather)")
