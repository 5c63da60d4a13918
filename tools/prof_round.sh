set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size
python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --train-steps 0 --cpu-baseline 0 > gpurun_out/ncu_bench.log 2>&1
python tools/prof_forward.py > gpurun_out/plain2.log 2>&1 && \
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/forward_metrics_r1.csv python tools/prof_forward.py > gpurun_out/ncu.log 2>&1
python tools/train_step_prof.py 8 > gpurun_out/plain3.log 2>&1 && \
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/train_metrics_r1.csv python tools/train_step_prof.py 8 > gpurun_out/train_prof.log 2>&1
python tools/prof_kernels.py gemm > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 2 -f -o gpurun_out/prof_gemm_r1c python tools/prof_kernels.py gemm > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
