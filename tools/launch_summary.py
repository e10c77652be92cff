"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py file.csv [first] [count]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    seq.append((re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mmf::<unnamed>::", ""), v))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else len(seq)
seq = seq[first:first + count]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in seq:
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(seq)} launches, {tot:.1f} us in kernels")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}% {v[0]:5d} x {v[1] / v[0]:7.1f}  {k[:100]}")
