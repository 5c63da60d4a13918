#!/bin/bash
# fused attention kernel: correctness (TMA + cp.async producers), timing per stagger mode, in-kernel phase clocks
mkdir -p gpurun_out
for st in 1 0 2; do
  MST_ATTN_STAGGER=$st timeout 300 python tools/attn_fused_check.py > gpurun_out/attn_check_st$st.log 2>&1; echo "stagger $st rc=$?"
  grep -E "BAD|time:|ALL OK|FAILED|worst|per-|qkv diff|Error|error" gpurun_out/attn_check_st$st.log | head -30
done
MST_ATTN_TMA=0 timeout 300 python tools/attn_fused_check.py > gpurun_out/attn_check_notma.log 2>&1; echo "notma rc=$?"
grep -E "BAD|ALL OK|FAILED|Error|error" gpurun_out/attn_check_notma.log | head -20
MST_LIB_PATH=mastermetastyletransfer_b200/libmst_b200_prof.so timeout 120 python tools/attn_prof.py 2>&1 | grep -E "af prof|----" | head -16
