"""Summarise an `ncu --set full` report (.ncu-rep): headline metrics, stall reasons, memory request efficiency, hottest source lines.
   python tools/ncu_report_summary.py gpurun_out/x.ncu-rep [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    f = lambda k: float(d[k].replace(",", "")) if d.get(k) not in (None, "", "n/a") else float("nan")
    print("kernel:", d["Kernel Name"][:90], " grid", d.get("Grid Size"), "block", d.get("Block Size"))
    print(f"  duration {f('gpu__time_duration.sum'):.1f} us | dram rd {f('dram__bytes_read.sum'):.1f} wr {f('dram__bytes_write.sum'):.1f} ({units[hdr.index('dram__bytes_read.sum')]}) "
          f"| dram {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}% | tensor pipe {f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in d else float('nan'):.1f}% "
          f"| regs {d.get('launch__registers_per_thread')} | occupancy {f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}% | ipc {f('sm__inst_executed.avg.per_cycle_active') if 'sm__inst_executed.avg.per_cycle_active' in d else float('nan'):.2f}")
    stalls = {k: f(k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:7]
    print("  stalls (warps per issue):", ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for k, v in top))
    for k in ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
              "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum"):
        if k in d:
            print(f"    {k} = {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))) if nl > 0 else []
if rows:
    h = rows[0]
    try:
        si = h.index("Source"); wi = next(i for i, n in enumerate(h) if n.startswith("Warp Stall Sampling (All"))
        ii = next(i for i, n in enumerate(h) if n.startswith("Instructions Executed"))
        body = [r for r in rows[1:] if len(r) > wi and r[wi].replace(",", "").isdigit()]
        tot = sum(int(r[wi].replace(",", "")) for r in body) or 1
        print("  hottest source lines (share of warp-stall samples):")
        for r in sorted(body, key=lambda r: -int(r[wi].replace(",", "")))[:nl]:
            print(f"    {100*int(r[wi].replace(',', ''))/tot:5.1f}%  {r[si].strip()[:140]}")
    except (ValueError, StopIteration):
        print("  (no source page)")
