set -x
python tools/prof_forward.py > gpurun_out/plain5.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:window_attn -s 16 -c 1 -f -o gpurun_out/prof_attn_r1 python tools/prof_forward.py > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_fused -s 18 -c 1 -f -o gpurun_out/prof_mlp_r1 python tools/prof_forward.py > gpurun_out/ncu_mlp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_band -s 10 -c 1 -f -o gpurun_out/prof_band_r1 python tools/prof_forward.py > gpurun_out/ncu_band.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 56 -c 1 -f -o gpurun_out/prof_gemm_qkv_r1 python tools/prof_forward.py > gpurun_out/ncu_gemmqkv.log 2>&1
ls -la gpurun_out/*.ncu-rep
