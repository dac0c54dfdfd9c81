"""Where the GPU idles inside one training step: kernel timeline of the bench workload (torch profiler / CUPTI), gaps between
consecutive kernels on the stream, attributed to the kernel that FOLLOWS the gap (the one the host was late to launch).
usage: python tools/gpu_gaps.py > gpurun_out/gpu_gaps.txt"""
import collections
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gd_mae_b200  # noqa: E402,F401
from gd_mae_b200 import config  # noqa: E402
from gd_mae_b200.trainer import MAETrainer  # noqa: E402
from oracle import gdmae_oracle as O  # noqa: E402  (synthetic scene generator only)

torch.backends.cudnn.benchmark = True
cfg = config.builtin_cfg("waymo_ssl")
model = config.build_mae_model(cfg).cuda()
config.set_precision(model, "bf16", dense_spatial_features=False)
trainer = MAETrainer(model, cfg.OPTIMIZATION, total_steps=100)
pts = torch.from_numpy(O.synth_batch(list(range(8)), O.make_cfg("waymo_ssl"))).cuda()
PREFETCH = "--no-prefetch" not in sys.argv
nxt = {"points": pts, "batch_size": 8}


def one_step():
    """like bench.py: the next batch's index structures are prefetched while this step's backward is queued"""
    global nxt
    bd, nxt = nxt, {"points": pts, "batch_size": 8}
    trainer.step(bd, next_batch=nxt if PREFETCH else None)


for _ in range(5):
    one_step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:      # GPU activity only: CPU-side op recording would slow the host
    for _ in range(3):
        one_step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
adam = [i for i, e in enumerate(ev) if e["name"].startswith("adam")]
seg = ev[adam[-2] + 1: adam[-1] + 1]          # the last full step
busy = sum(e["dur"] for e in seg)
span = seg[-1]["ts"] + seg[-1]["dur"] - seg[0]["ts"]
print(f"last step: {len(seg)} GPU activities, busy {busy / 1e3:.2f} ms, span {span / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms")
gaps = collections.defaultdict(lambda: [0, 0.0])
big = []
end = seg[0]["ts"] + seg[0]["dur"]
for prev, e in zip(seg, seg[1:]):
    g = e["ts"] - end
    if g > 2:
        k = e["name"][:70]
        gaps[k][0] += 1
        gaps[k][1] += g
        big.append((g, prev["name"][:50], e["name"][:50]))
    end = max(end, e["ts"] + e["dur"])
print("gap time by the kernel that follows the gap:")
for k, (c, t) in sorted(gaps.items(), key=lambda x: -x[1][1])[:25]:
    print(f"  {t:8.1f} us  {c:4d}x  {k}")
print("largest single gaps (us, after -> before):")
for g, a, b in sorted(big, reverse=True)[:25]:
    print(f"  {g:8.1f}  {a}  ->  {b}")
# in-step (warm cache, steady-state clocks) time per kernel name: what the step really spends, next to the cold ncu list
busy_by = collections.defaultdict(lambda: [0, 0.0])
for e in seg:
    busy_by[e["name"][:90]][0] += 1
    busy_by[e["name"][:90]][1] += e["dur"]
print("in-step GPU time by kernel (last step):")
for k, (c, t) in sorted(busy_by.items(), key=lambda x: -x[1][1])[:int(os.environ.get("GAPS_TOP", "200"))]:
    print(f"  {t:8.1f} us  {c:4d}x  {100 * t / busy:5.1f} %  {k}")
