#!/bin/bash
# ncu --set full of the fused attention kernel (one launch per bench shape), after the same command ran clean without ncu
mkdir -p gpurun_out
python tools/attn_prof.py > gpurun_out/attn_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 1 -c 1 -o gpurun_out/r2_attn_fused_c128 -f python tools/attn_prof.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn.log; ls -la gpurun_out/*.ncu-rep
