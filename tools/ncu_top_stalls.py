"""Print the instructions with the most warp-stall samples from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
si, src, ie, ad = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Address')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[si].isdigit()]
sass = [r for r in data if r[ad].strip()]
tot = sum(int(r[si]) for r in sass)
print('total samples', tot, 'sass instr', len(sass))
agg = {}
for r in sass:
    for i in stall_cols:
        if r[i].isdigit():
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(sass, key=lambda r: -int(r[si]))[:n]:
    st = {hdr[i]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}% ex={r[ie]:>8s} {r[src][:84]:84s} {st}")
