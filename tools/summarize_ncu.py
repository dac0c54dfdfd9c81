"""ncu report(s) -> markdown summary for profiles/ (read on the CPU box: `ncu -i … --page raw --csv`).
  python tools/summarize_ncu.py gpurun_out/x.ncu-rep [--title "…"] > profiles/r2_ncu_x.md
  python tools/summarize_ncu.py --launches gpurun_out/launches.csv [--last-fraction 0.333] > profiles/r2_step_summary.md"""
import argparse
import collections
import csv
import io
import subprocess
import sys

METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (legacy HMMA counter)"),
           ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
           ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("smsp__inst_executed.sum", "warp instructions"),
           ("lts__t_sector_hit_rate.pct", "L2 hit rate %")]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def report(rep, title):
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n\nsource: `{rep}` (`ncu --set full --clock-control none --import-source on`, one B200; per-launch values are"
          " cold-cache and serialised)\n")
    for r in rows:
        print(f"## `{r[idx['Kernel Name']][:110]}`\n\n| metric | value |\n|---|---|")
        for m, label in METRICS:
            if m in idx and r[idx[m]] != "":
                print(f"| {label} (`{m}`) | {r[idx[m]]} {units[idx[m]]} |")
        try:
            t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            unit = units[idx["gpu__time_duration.sum"]]
            t_us = t * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(unit, 1.0)
            rd, wr = float(r[idx["dram__bytes_read.sum"]].replace(",", "")), float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * scale.get(units[idx["dram__bytes_read.sum"]], 1.0) + wr * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
            print(f"| DRAM bytes / duration | {tot / t_us / 1e3:.0f} GB/s |")
        except Exception:
            pass
        print()


def launches(path, last_fraction):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 2:]
    iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    for r in data:
        if len(r) > iV:
            per.setdefault(int(r[iID]), {"name": r[iK]})[r[iM]] = float(r[iV].replace(",", ""))
    ids = sorted(per)
    last = ids[int(len(ids) * (1 - last_fraction)):]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i in last:
        d = per[i]
        nm = d["name"].split("(")[0][-72:]
        a = agg[nm]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0) / 1e3
        a[2] += d.get("dram__bytes_read.sum", 0) / 1e6
        a[3] += d.get("dram__bytes_write.sum", 0) / 1e6
    tot = sum(a[1] for a in agg.values())
    own = sum(a[1] for n, a in agg.items() if not any(t in n for t in ("cutlass", "nvjet", "at::", "elementwise", "cub::", "Memset", "nccl", "void_bf16", "ndhwc", "tf32gemm")))
    print(f"# kernels of one training step (last {len(last)} launches of `{path}`)\n\n`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
          f"dram__bytes_write.sum --clock-control none`; cold-cache, serialised launches: compare SHARES.\n\ntotal {tot:.0f} us, hand-written "
          f"kernels {own:.0f} us ({100 * own / tot:.1f} %)\n\n| kernel | launches | us | share | DRAM read MB | DRAM write MB | DRAM GB/s |\n|---|---|---|---|---|---|---|")
    for nm, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if a[1] < 0.002 * tot:
            continue
        print(f"| `{nm}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[2]:.1f} | {a[3]:.1f} | {(a[2] + a[3]) / a[1] * 1e3:.0f} |")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("rep", nargs="?")
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--launches")
    ap.add_argument("--last-fraction", type=float, default=1 / 3)
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, a.last_fraction)
    else:
        report(a.rep, a.title)
