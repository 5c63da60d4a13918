"""Per-kernel-family DRAM traffic from an ncu per-launch metrics CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum ... --csv`, long format) -> profiles/ncu_traffic.json (read by bench.py's roofline.traffic).

    python tools/ncu_traffic.py gpurun_out/forward_metrics_r1f.csv b32_256 [launches_per_forward]
The capture holds several identical forwards; only the LAST `launches_per_forward` launches (one warm forward) are used.
"""
import collections, csv, json, os, re, sys

path_csv, key = sys.argv[1], sys.argv[2]
per_fwd = int(sys.argv[3]) if len(sys.argv) > 3 else 54
rows = [r for r in csv.reader(open(path_csv)) if len(r) > 8]
hdr = next(r for r in rows if "Kernel Name" in r)
I = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
launches = collections.OrderedDict()
for r in rows:
    if r is hdr or not r[I["ID"]].isdigit():
        continue
    name = re.sub(r"\(.*", "", r[I["Kernel Name"]]).replace("void ", "").replace("mst::", "").split("<")[0]
    d = launches.setdefault(int(r[I["ID"]]), {"name": name})
    try:
        d[r[I["Metric Name"]]] = float(r[I["Metric Value"]].replace(",", "")) * mult.get(r[I["Metric Unit"]], 1.0)
    except ValueError:
        pass
last = list(launches.values())[-per_fwd:]
fam = collections.OrderedDict()
for d in last:
    f = fam.setdefault(d["name"], {"launches": 0, "rd": 0.0, "wr": 0.0, "us": 0.0})
    f["launches"] += 1
    f["rd"] += d.get("dram__bytes_read.sum", 0.0)
    f["wr"] += d.get("dram__bytes_write.sum", 0.0)
    f["us"] += d.get("gpu__time_duration.sum", 0.0)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
data = json.load(open(out)) if os.path.exists(out) else {}
data[key] = {k: {"launches": v["launches"], "dram_bytes_per_launch": (v["rd"] + v["wr"]) / v["launches"],
                 "dram_read_per_launch": v["rd"] / v["launches"], "dram_write_per_launch": v["wr"] / v["launches"],
                 "us_per_launch_under_ncu": v["us"] / v["launches"], "source": os.path.basename(path_csv)} for k, v in fam.items()}
json.dump(data, open(out, "w"), indent=1, sort_keys=True)
for k, v in sorted(data[key].items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"] * kv[1]["launches"]):
    print(f"{k:28s} {v['launches']:3d} launches  {v['dram_bytes_per_launch']/1e6:9.1f} MB/launch  {v['us_per_launch_under_ncu']:8.1f} us/launch (ncu)")
