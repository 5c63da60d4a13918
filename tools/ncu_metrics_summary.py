"""Summarise an `ncu --metrics ... --csv` capture: one line per launch + per-kernel-family totals."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
hdr = next(r for r in rows if "Kernel Name" in r)
I = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launches = collections.OrderedDict()
for r in rows:
    if r is hdr or not r[I["ID"]].isdigit():
        continue
    d = launches.setdefault(int(r[I["ID"]]), {"name": re.sub(r"\(.*", "", r[I["Kernel Name"]]).replace("void ", "").replace("mst::", "")})
    v = r[I["Metric Value"]].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        pass
    unit = r[I["Metric Unit"]]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(unit, 1)
    d[r[I["Metric Name"]]] = v * mult if isinstance(v, float) else v
fam = collections.OrderedDict()
print(f"{'kernel':28s} {'grid':>6s} {'time_us':>9s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'dram_GB/s':>9s} {'L2_MB':>8s} {'tensor%':>7s} {'warps%':>6s} {'dram%':>6s}")
for i, d in launches.items():
    t = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    print(f"{d['name']:28s} {int(d.get('launch__grid_size', 0)):6d} {t*1e6:9.1f} {rd/1e6:10.1f} {wr/1e6:10.1f} {(rd+wr)/t/1e9 if t else 0:9.0f} "
          f"{d.get('lts__t_bytes.sum', 0)/1e6:8.1f} {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):7.1f} "
          f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):6.1f} {d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):6.1f}")
    f = fam.setdefault(d["name"].split("<")[0], [0, 0.0, 0.0, 0.0])
    f[0] += 1; f[1] += t; f[2] += rd + wr; f[3] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) * t
print("\n# per family (times are ncu's serialised cold-cache durations)")
tot = sum(f[1] for f in fam.values())
for k, f in fam.items():
    print(f"{k:28s} {f[0]:3d} launches {f[1]*1e6:9.1f} us ({100*f[1]/tot:5.1f}%)  dram traffic {f[2]/1e6:9.1f} MB  {f[2]/f[1]/1e9:7.0f} GB/s  time-weighted tensor pipe {f[3]/f[1]:5.1f}%")
