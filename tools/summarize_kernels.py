"""Per-kernel DRAM traffic table from an ncu metric pass over the step:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,
      smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,
      launch__grid_size,launch__block_size --clock-control none --csv --log-file X python tools/ncu_target.py --steps 3
  python tools/summarize_kernels.py X > profiles/r1_ncu_kernels_vNN.md
The last full step (between the last two adam_kernel launches) is aggregated per kernel name: launches, summed time, DRAM bytes
read / written per launch, achieved DRAM rate (actual traffic / time) and its fraction of the measured peak."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6650.0
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
iid, ik, im, iu, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
launch = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= iv:
        continue
    d = launch.setdefault(int(r[iid]), {"name": r[ik]})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    d[r[im]] = v * scale
ls = list(launch.values())
adam = [i for i, d in enumerate(ls) if d["name"].startswith("adam")]
seg = ls[adam[-2] + 1: adam[-1] + 1]
agg = collections.OrderedDict()
for d in seg:
    a = agg.setdefault(d["name"], collections.defaultdict(float))
    a["n"] += 1
    for k, v in d.items():
        if k != "name":
            a[k] += v
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
print(f"launches in the step: {len(seg)}; summed kernel time {tot / 1e3:.2f} ms (ncu, cold caches, serialised); peak = {peak:.0f} GB/s (measured copy)\n")
print("| kernel | launches | total us | DRAM read MB / launch | DRAM write MB / launch | DRAM GB/s (actual) | of peak | issue active % | regs |")
print("|---|---|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"])[:60]:
    n, t = a["n"], a["gpu__time_duration.sum"]
    rd, wr = a["dram__bytes_read.sum"] / n, a["dram__bytes_write.sum"] / n
    gbs = (a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / (t * 1e-6) / 1e9 if t > 0 else 0.0
    print(f"| `{name[:80]}` | {int(n)} | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {gbs / peak:.2f} | "
          f"{a['smsp__issue_active.avg.pct_of_peak_sustained_active'] / n:.0f} | {a['launch__registers_per_thread'] / n:.0f} |")
