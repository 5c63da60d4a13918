#!/bin/bash
# Steps 1-2 of tools/prof_round_r2.sh only (launch list of the bench command, per-launch metrics of one eager forward): refreshes
# profiles/r2_launches_bench_b32_256.* and profiles/ncu_traffic.json after a change of the launch sequence, without the --set full captures.
K='regex:gemm_tc|mlp_fused|attn_fused|attn_core|window_attn|conv_rows|conv_cm|conv_band|layernorm|instnorm|patch_embed|cast_bf16|upsample'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size
BENCH="python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 --config5 0"
$BENCH > gpurun_out/r2_plain_bench.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench_b32_256.csv $BENCH > gpurun_out/r2_ncu_bench.log 2>&1
python tools/prof_forward.py > gpurun_out/r2_plain_forward.log 2>&1 || exit 1
timeout 600 ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/r2_forward_metrics_b32_256.csv python tools/prof_forward.py > gpurun_out/r2_ncu_forward.log 2>&1
ls -la gpurun_out/ | grep r2_
