"""Per-kernel-family DRAM traffic from an `ncu --set full` report -> profiles/ncu_traffic.json (read by bench.py's roofline.traffic).

    python tools/ncu_traffic.py gpurun_out/prof_fwd_full.ncu-rep b32_256
"""
import csv, io, json, os, re, subprocess, sys

rep, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
units = rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
fam = {}
for r in rows[2:]:
    name = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
    name = re.sub(r"^mst::", "", name).split("<")[0].split("(")[0]
    rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    dur = float(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
    f = fam.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0, "us_under_ncu": 0.0})
    f["launches"] += 1
    f["dram_bytes"] += rd + wr
    f["dram_read"] += rd
    f["dram_write"] += wr
    f["us_under_ncu"] += dur
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["dram_bytes"] / v["launches"],
                 "dram_read_per_launch": v["dram_read"] / v["launches"], "dram_write_per_launch": v["dram_write"] / v["launches"],
                 "us_per_launch_under_ncu": v["us_under_ncu"] / v["launches"], "source": os.path.basename(rep)} for k, v in fam.items()}
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
for k, v in sorted(data[key].items(), key=lambda kv: -kv[1]["dram_bytes_per_launch"] * kv[1]["launches"]):
    print(f"{k:28s} {v['launches']:3d} launches  {v['dram_bytes_per_launch']/1e6:9.1f} MB/launch  {v['us_per_launch_under_ncu']:8.1f} us/launch (ncu)")
