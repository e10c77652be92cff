#!/usr/bin/env python
"""Summarise ncu CSV exports into the tracked profiles/ directory.

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>
  raw page    : ncu -i report.ncu-rep --page raw --csv > raw.csv

usage: summarize_ncu.py launches <launches.csv> <out.md> [timesteps]
       summarize_ncu.py raw <raw.csv> <out.md>
"""
import collections
import csv
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def short(name):
    name = name.replace("void ", "").replace("mmf::<unnamed>::", "").replace("unnamed>::", "")
    return name.split("(")[0][:48]


def launches(path, out, timesteps=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mv, gs, bs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        agg.setdefault((short(r[kn]), r[gs], r[bs]), []).append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path})\n\n`gpu__time_duration.sum`, `--clock-control none`; serialised, cold-cache per-launch times: "
                f"use the SHARES, not the absolutes.\n\n{n} launches, {tot / 1e3:.1f} us total"
                + (f" = {tot / 1e3 / timesteps:.1f} us and {n / timesteps:.0f} launches per timestep ({timesteps} timesteps)" if timesteps else "")
                + "\n\n| kernel | grid | block | launches | avg us | total us | share |\n|---|---|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[0]}` | {k[1]} | {k[2]} | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |\n")


def raw(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    h, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full capture ({path})\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[h.index('Kernel Name')])}` grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in RAW_KEYS:
                if k in h:
                    f.write(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |\n")
            st = sorted(((float(r[i].replace(",", "") or 0), n) for i, n in enumerate(h)
                         if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")), reverse=True)[:5]
            f.write("\ntop stall reasons (warps per issue-active cycle): " + ", ".join(f"{n.split('stalled_')[1].split('_per')[0]} {v:.2f}" for v, n in st) + "\n\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)
    else:
        raw(sys.argv[2], sys.argv[3])
