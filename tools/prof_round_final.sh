#!/bin/bash
# Round-end profiling pass (run under gpurun, one GPU).  Every ncu command only after the same program exited 0 without ncu.
#   1. launch list of the bench command                  -> gpurun_out/launches_${TAG}.csv
#   2. per-launch metrics of one eager forward            -> gpurun_out/forward_metrics_${TAG}.csv
#   3. `--set full` capture of ONE whole forward (third)  -> gpurun_out/prof_fwd_full_${TAG}.ncu-rep  (traffic, stalls, source)
TAG=${1:-r1e}
K='regex:gemm_tc|mlp_fused|window_attn|conv_rows|conv_band|layernorm|instnorm|patch_embed|cast_bf16'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size
set -x
python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 > gpurun_out/ncu_bench_${TAG}.log 2>&1
python tools/prof_forward.py > gpurun_out/plain2_${TAG}.log 2>&1 || exit 1
ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/forward_metrics_${TAG}.csv python tools/prof_forward.py > gpurun_out/ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 108 -c 54 -f -o gpurun_out/prof_fwd_full_${TAG} python tools/prof_forward.py > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*
