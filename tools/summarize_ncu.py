#!/usr/bin/env python
"""Summarise ncu CSV exports into the tracked profiles/ directory.

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>
  raw page    : ncu -i report.ncu-rep --page raw --csv > raw.csv

usage: summarize_ncu.py launches <launches.csv> <out.md> [timesteps]
       summarize_ncu.py hot <source_page.csv> <out.md>      (ncu -i rep --page source --csv --print-source cuda,sass)
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


EXTRA_SUFFIX_KEYS = [
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def hot_lines(path, out, top=40):
    """ncu -i rep --page source --csv --print-source cuda,sass  ->  warp-stall samples per CUDA source line."""
    rows = list(csv.reader(open(path, errors="replace")))
    cur, h = None, None
    samples, execd, text, stalls = collections.Counter(), collections.Counter(), {}, collections.defaultdict(collections.Counter)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            h = r
            isamp, iex = h.index("# Samples"), h.index("Instructions Executed")
            scols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            continue
        if h is None or len(r) <= isamp or r[0] in ("", "Function Name") or r[2] != "-":
            continue
        try:
            n, e = int(r[isamp]), int(r[iex])
        except ValueError:
            continue
        key = (cur, int(r[0]))
        samples[key] += n; execd[key] += e; text[key] = r[1].strip()[:70]
        for i in scols:
            try:
                stalls[key][h[i][6:]] += int(r[i])
            except ValueError:
                pass
    tot, totex = sum(samples.values()) or 1, sum(execd.values()) or 1
    allst = collections.Counter()
    for k, c in stalls.items():
        if not (k[0] == "mmf_ptx.cuh" and "try_wait" in text[k] or "spins" in text[k]):
            allst.update(c)
    with open(out, "w") as f:
        f.write(f"# warp-stall samples per CUDA source line ({path})\n\n{tot} samples, {totex} warp instructions executed.\n\n"
                "Stall mix outside the barrier-wait loops: " + ", ".join(f"{a} {100 * b / (sum(allst.values()) or 1):.1f}%" for a, b in allst.most_common(8)) + "\n\n"
                "| file:line | samples | executed | source | top stalls |\n|---|---:|---:|---|---|\n")
        for k, n in samples.most_common(top):
            f.write(f"| {k[0]}:{k[1]} | {100 * n / tot:.1f}% | {100 * execd[k] / totex:.1f}% | `{text[k].replace('|', '/')}` | "
                    + " ".join(f"{a}:{b}" for a, b in stalls[k].most_common(3)) + " |\n")


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
            for k in RAW_KEYS + EXTRA_SUFFIX_KEYS:
                hit = [i for i, n in enumerate(h) if n == k or n.endswith("." + k)]      # some sections prefix the metric name
                if hit:
                    f.write(f"| {k} | {r[hit[0]]} | {units[hit[0]]} |\n")
            st = sorted(((float(r[i].replace(",", "") or 0), n) for i, n in enumerate(h)
                         if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")), reverse=True)[:5]
            f.write("\ntop stall reasons (warps per issue-active cycle): " + ", ".join(f"{n.split('stalled_')[1].split('_per')[0]} {v:.2f}" for v, n in st) + "\n\n")


if __name__ == "__main__":
    if sys.argv[1] == "hot":
        hot_lines(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)
    else:
        raw(sys.argv[2], sys.argv[3])
