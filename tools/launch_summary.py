"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel launches / total time / share."""
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; iN = h.index("Kernel Name"); iV = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = re.sub(r"\(.*", "", r[iN]).replace("void ", "").replace("cb::", "")
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[iV].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"# ncu launch list (gpu__time_duration.sum, --clock-control none), {len(rows) - 1} launches of `{sys.argv[2] if len(sys.argv) > 2 else 'bench.py'}`")
print("# (cold-cache, serialised per-launch times: compare SHARES with bench.py's live CUDA-event shares, not absolutes)")
print(f"# total {tot} ns over {len(rows) - 1} launches")
print("kernel,launches,total_ns,share")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{a[0]},{a[1]},{a[1] / tot:.4f}")
