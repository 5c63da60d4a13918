"""Tensor-core / TMA / TMEM instruction counts per kernel of libmst_b200.so (cuobjdump -sass) -> profiles/r2_sass_tensor_ops.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mastermetastyletransfer_b200", "libmst_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, cnt, ex = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    mm = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?(UTCHMMA|UTCQMMA|UTCCP|UTMALDG|UTMASTG|UBLKCP|UTCBAR|LDTM|STTM|HMMA|SYNCS)\b[^;]*;", line)
    if mm:
        cnt.setdefault(cur, collections.Counter())[mm.group(1)] += 1
        ex.setdefault((cur, mm.group(1)), mm.group(0).replace("*/", "").strip())
lines = ["# Tensor-core / TMA / TMEM instruction counts per kernel in libmst_b200.so (cuobjdump -sass, sm_100a), with one example each.",
         "# UTCHMMA = tcgen05.mma (kind::f16; a tmem[...] first operand = A from tensor memory), UTCCP = tcgen05.cp, UTMALDG / UTMASTG =",
         "# cp.async.bulk.tensor load / store (TMA), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,",
         "# HMMA = mma.sync (legacy kernels), SYNCS = mbarrier ops.", ""]
for k, d in cnt.items():
    if not any(op in d for op in ("UTCHMMA", "UTCCP", "UTMALDG", "UTMASTG", "HMMA")):
        continue
    name = re.sub(r"\(.*", "", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip())
    lines.append(name)
    lines.append("    " + "  ".join(f"{op}={n}" for op, n in sorted(d.items())))
    for op in ("UTCHMMA", "UTCCP", "UTMALDG", "UTMASTG", "HMMA"):
        if (k, op) in ex:
            lines.append(f"    e.g. {ex[(k, op)]}")
open(os.path.join(ROOT, "profiles", "r2_sass_tensor_ops.txt"), "w").write("\n".join(lines) + "\n")
print(len(lines), "lines")
