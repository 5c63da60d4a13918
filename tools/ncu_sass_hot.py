"""Hottest SASS instructions of an ncu --set full report: by warp-stall samples and by executed count.
   python tools/ncu_sass_hot.py rep.ncu-rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[i0]; body = [r for r in rows[i0 + 1:] if len(r) == len(h)]
si, wi, ei = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
tot_s = sum(int(r[wi]) for r in body) or 1; tot_e = sum(int(r[ei]) for r in body) or 1
print(f"{len(body)} instructions, {tot_s} stall samples, {tot_e} warp-instructions executed")
idx = {id(r): k for k, r in enumerate(body)}
print("--- by stall samples")
for r in sorted(body, key=lambda r: -int(r[wi]))[:n]:
    print(f"{100*int(r[wi])/tot_s:5.1f}%  exec {100*int(r[ei])/tot_e:5.1f}%  [{idx[id(r)]:5d}] {r[si].strip()[:100]}")
print("--- by executed count")
for r in sorted(body, key=lambda r: -int(r[ei]))[:n]:
    print(f"exec {100*int(r[ei])/tot_e:5.1f}%  stall {100*int(r[wi])/tot_s:5.1f}%  [{idx[id(r)]:5d}] {r[si].strip()[:100]}")
