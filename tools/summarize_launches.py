"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python
tools/ncu_target.py --steps 3`) into the per-kernel share table kept under profiles/.
usage: python tools/summarize_launches.py gpurun_out/launches.csv profiles/r1_launches_one_step_vNN.csv > profiles/r1_step_summary_vNN.md
The LAST full step (launches between the last two `adam_kernel` launches) is extracted; its per-launch rows are written
to the second argument."""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [(r[ik], float(r[iv].replace(",", "")) / 1000.0) for r in rows[start + 1:] if len(r) > iv]
adam = [i for i, (n, _) in enumerate(names) if n.startswith("adam")]
seg = names[adam[-2] + 1: adam[-1] + 1]
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel", "gpu__time_duration_us"])
    for i, (n, t) in enumerate(seg):
        w.writerow([i, n, f"{t:.3f}"])
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in seg:
    agg[n][0] += 1
    agg[n][1] += t
tot = sum(v[1] for v in agg.values())
own = ("sra_", "vfe", "bn_", "tail_", "dense_fill", "gather", "colsum", "add_ln", "bias_gelu", "chamfer", "segment", "vox_", "mask_",
       "win_", "vis_", "down_", "subm_", "up_map", "rank_grid", "pos_lut", "cast_bf16", "dtau", "adam", "sumsq", "group_points",
       "scatter", "bucket", "fill", "partial")
mine = [(n, c, t) for n, (c, t) in agg.items() if any(k in n.split("(")[0] for k in own) and "cutlass" not in n and "cub::" not in n
        and "native::" not in n and "nvjet" not in n and "cublas" not in n]
print(f"launches in the step: {len(seg)}; summed kernel time {tot / 1000:.2f} ms (cold-cache, serialised: compare SHARES); "
      f"hand-written kernels: {sum(c for _, c, _ in mine)} launches, {sum(t for _, _, t in mine) / 1000:.2f} ms "
      f"({100 * sum(t for _, _, t in mine) / tot:.0f}% of kernel time)\n")
print("| share | total us | launches | kernel |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print(f"| {100 * t / tot:.1f}% | {t:.1f} | {c} | `{n[:90]}` |")
