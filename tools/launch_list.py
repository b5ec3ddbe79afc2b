"""tools/launch_list.py launches.csv -- per-kernel summary (count, mean duration, share) of an
`ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
mu = h.index("Metric Unit")
gs = h.index("Grid Size") if "Grid Size" in h else None
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] in ("ns", "nsecond") else (v * 1e3 if r[mu] in ("ms", "msecond") else v)
    # keyed by grid size too: bench.py runs the same kernels on the 4096-car workload and on its 16x "saturated" replica
    agg[r[kn][:58] + ("  grid " + r[gs].strip("()").split(",")[0] if gs is not None else "")].append(v)
tot = sum(sum(v) for v in agg.values())
print("%-70s %6s %12s %8s" % ("kernel", "n", "mean us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-70s %6d %12.2f %7.1f%%" % (k[:70], len(v), sum(v) / len(v), 100 * sum(v) / tot))
