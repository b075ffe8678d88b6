"""Build profiles/traffic.json (ncu DRAM bytes per launch of the main kernel families) from the raw csv of tools/ncu_sweep.sh.

The sweep profiles ONE 3840-frame PPO minibatch step, so the launches of a kernel appear in execution order and the families
that share a kernel (forward conv / dgrad) are told apart by their position:
  k_conv_umma<2,16>: 4 forward, then 4 dgrad (42x42)
  k_conv_umma<4,32>: 4 forward at 21x21, 4 forward at 11x11, 4 dgrad at 11x11, then 5 dgrad at 21x21
  k_wgrad_umma<4,32>: 4 at 11x11, then 5 at 21x21 (backward runs the last ConvSequence first)
usage: python tools/make_traffic.py gpurun_out/<tag>_sweep_learner_raw.csv <tag> > profiles/traffic.json"""
import collections, csv, json, re, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
iI, iN, iM, iV = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    if r[iM] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        k = (int(r[iI]), re.sub(r"\(.*", "", r[iN]).replace("void ", "").replace("cb::", "").replace(" ", ""))
        per[k] = per.get(k, 0.0) + float(r[iV].replace(",", ""))
seq = collections.defaultdict(list)
for (i, name), v in sorted(per.items()):
    seq[name].append(v)
unit = None
for r in rows[1:]:
    if r[iM] == "dram__bytes_read.sum":
        unit = r[h.index("Metric Unit")]
        break
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
mean = lambda xs: int(sum(xs) / len(xs) * mult)
out = {"_source": f"profiles/{sys.argv[2]}_ncu_sweep_learner_mb3840.txt (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one 3840-frame "
                  "minibatch step, mean per launch of the family; tools/make_traffic.py)"}
c16, c32, w32 = seq["k_conv_umma<2,16>"], seq["k_conv_umma<4,32>"], seq["k_wgrad_umma<4,32>"]
assert len(c16) == 8 and len(c32) == 17 and len(w32) == 9, (len(c16), len(c32), len(w32))
out["conv_fwd<cin16,cout16>@42x42"] = mean(c16[:4]); out["conv_dgrad<cin16,cout16>@42x42"] = mean(c16[4:])
out["conv_wgrad<cin16,cout16>@42x42"] = mean(seq["k_wgrad_umma<2,16>"])
out["conv_fwd<cin32,cout32>@21x21"] = mean(c32[:4]); out["conv_fwd<cin32,cout32>@11x11"] = mean(c32[4:8])
out["conv_dgrad<cin32,cout32>@11x11"] = mean(c32[8:12]); out["conv_dgrad<cin32,cout32>@21x21"] = mean(c32[12:])
out["conv_wgrad<cin32,cout32>@11x11"] = mean(w32[:4]); out["conv_wgrad<cin32,cout32>@21x21"] = mean(w32[4:])
out["conv_dgrad<cin32,cout16>@42x42"] = mean(seq["k_conv_umma<4,16>"]); out["conv_wgrad<cin16,cout32>@42x42"] = mean(seq["k_wgrad_umma<2,32>"])
out["conv0_pool_fwd@84"] = mean(seq["k_conv0_pool_umma"])
pools = sorted(k for k in seq if k.startswith("k_conv_pool_umma"))
for k in pools:
    out["conv_pool_fwd<cin16,cout32>@42" if k.startswith("k_conv_pool_umma<2") else "conv_pool_fwd<cin32,cout32>@21"] = mean(seq[k])
out["pool_bwd_wgrad0@84"] = mean(seq["k_pool_bwd_wgrad0"])
pb = seq["k_pool_bwd"]
out["pool_bwd@21"], out["pool_bwd@42"] = int(min(pb) * mult), int(max(pb) * mult)
print(json.dumps(out, indent=1))
