"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of step)."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mst::", "")
    t = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
unit = rows[rows.index(hdr) + 1][hdr.index("Metric Unit")] if len(rows) > rows.index(hdr) + 1 else "ns"
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, total {tot:.0f} {unit} (cold-cache, serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {n:5d} launches {t:14.0f} {unit} {100*t/tot:6.2f}%")
