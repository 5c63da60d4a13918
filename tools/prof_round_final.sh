#!/bin/bash
# Round-end profiling pass (run under gpurun, one GPU).  Every ncu command only after the same program exited 0 without ncu.
#   1. launch list of the bench command                  -> gpurun_out/launches_${TAG}.csv
#   2. per-launch metrics of one eager forward            -> gpurun_out/forward_metrics_${TAG}.csv
#   3. `--set full` captures of one representative launch of each heavy family, summarised ON THE BOX (gpurun_out/ is capped at
#      64 MiB and reports with imported source are large): headline metrics, stall reasons, hottest SASS; reports then deleted.
# DRAM traffic per kernel family (roofline.traffic) is derived here, from (2), with tools/ncu_traffic.py.
TAG=${1:-r1e}
K='regex:gemm_tc|mlp_fused|window_attn|conv_rows|conv_band|layernorm|instnorm|patch_embed|cast_bf16'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size
set -x
python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 > gpurun_out/ncu_bench_${TAG}.log 2>&1
python tools/prof_forward.py > gpurun_out/plain2_${TAG}.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/forward_metrics_${TAG}.csv python tools/prof_forward.py > gpurun_out/ncu_${TAG}.log 2>&1
# hottest SASS of the three heaviest families: one representative launch each (re-captured alone: small reports)
# label:kernel regex:launches to skip (second forward: 9 mlp_fused, 8 window_attn, 19 gemm_tc, 4 conv_rows launches per forward)
for spec in "mlp_c128:mlp_fused:9" "attn_ws7:window_attn:8" "gemm_qkv_tma:gemm_tc:19" "gemm_conv128to64_tma:gemm_tc:34" "conv_rows_32ch:conv_rows:6"; do
  label=${spec%%:*}; rest=${spec#*:}; name=${rest%%:*}; skip=${rest##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o /tmp/one_$label python tools/prof_forward.py > /dev/null 2>&1
  python tools/ncu_report_summary.py /tmp/one_$label.ncu-rep 0 > gpurun_out/ncu_full_${label}_${TAG}.txt 2>&1
  python tools/ncu_sass_hot.py /tmp/one_$label.ncu-rep 14 >> gpurun_out/ncu_full_${label}_${TAG}.txt 2>&1
done
rm -f /tmp/*.ncu-rep
ls -la gpurun_out/
