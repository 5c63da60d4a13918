#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU).  Every ncu command only after the same program exited 0 without ncu.
#   1. launch list of the bench command                  -> gpurun_out/r2_launches_bench_b32_256.csv
#   2. per-launch metrics of one eager forward            -> gpurun_out/r2_forward_metrics_b32_256.csv  (-> profiles/ncu_traffic.json)
#   3. `--set full` captures of one representative launch of each heavy family, summarised ON THE BOX (headline metrics, stall
#      reasons, hottest SASS); the reports themselves are deleted (large with imported source).
K='regex:gemm_tc|mlp_fused|attn_fused|attn_core|window_attn|conv_rows|conv_cm|conv_band|layernorm|instnorm|patch_embed|cast_bf16|upsample'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size
BENCH="python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 --config5 0"
$BENCH > gpurun_out/r2_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench_b32_256.csv $BENCH > gpurun_out/r2_ncu_bench.log 2>&1
python tools/prof_forward.py > gpurun_out/r2_plain_forward.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/r2_forward_metrics_b32_256.csv python tools/prof_forward.py > gpurun_out/r2_ncu_forward.log 2>&1
# one representative launch per family from the THIRD forward (per forward: 9 mlp_fused, 6 attn_fused, 8 gemm_tc, 5 conv_cm, 4 conv_rows)
for spec in "mlp_c128_pre:mlp_fused:18" "mlp_c256_pre:mlp_fused:22" "attn_fused_c256_ws8:attn_fused:16" "conv_cm_128to128:conv_cm:11" "conv_cm_128to64:conv_cm:14" "attn_core_ws8:attn_core:4" "conv_rows_32ch:conv_rows:10"; do
  label=${spec%%:*}; rest=${spec#*:}; name=${rest%%:*}; skip=${rest##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o /tmp/one_$label python tools/prof_forward.py > /dev/null 2>&1
  python tools/ncu_report_summary.py /tmp/one_$label.ncu-rep 0 > gpurun_out/r2_ncu_full_${label}.txt 2>&1
  python tools/ncu_sass_hot.py /tmp/one_$label.ncu-rep 14 >> gpurun_out/r2_ncu_full_${label}.txt 2>&1
done
rm -f /tmp/*.ncu-rep
ls -la gpurun_out/ | grep r2_
