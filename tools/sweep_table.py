"""Summarise an ncu --metrics --csv sweep (one row per launch) into a per-kernel table."""
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; iN = h.index("Kernel Name"); iM = h.index("Metric Name"); iV = h.index("Metric Value"); iI = h.index("ID")
by = collections.OrderedDict()
for r in rows[1:]:
    by.setdefault(r[iI], {"name": r[iN]})[r[iM]] = float(r[iV].replace(",", ""))
agg = collections.OrderedDict()
for k, d in by.items():
    n = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("cb::", "")
    key = (n, int(d.get("launch__grid_size", 0)))
    a = agg.setdefault(key, collections.Counter())
    a["n"] += 1
    for m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum"):
        a[m] += d.get(m, 0)
    a["tensor"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
    a["issue"] += d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)
    a["regs"] = d.get("launch__registers_per_thread", 0)
tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
print(f"{'kernel':46s} {'grid':>7s} {'n':>3s} {'us':>9s} {'%':>5s} {'dramGB':>7s} {'GB/s':>7s} {'ltsGB':>7s} {'lts GB/s':>8s} {'tens%':>5s} {'iss%':>5s} {'regs':>4s}")
for (n, g), a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = a["gpu__time_duration.sum"]; db = a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]
    print(f"{n[:46]:46s} {g:7d} {a['n']:3d} {t/1e3:9.1f} {100*t/tot:5.1f} {db/1e9:7.3f} {db/t:7.0f} {a['lts__t_bytes.sum']/1e9:7.2f} {a['lts__t_bytes.sum']/t:8.0f} {a['tensor']/a['n']:5.1f} {a['issue']/a['n']:5.1f} {int(a['regs']):4d}")
print(f"total {tot/1e6:.3f} ms")
