#!/bin/bash
# fused attention kernel: correctness (TMA + cp.async producers), timing, in-kernel phase clocks
mkdir -p gpurun_out
timeout 300 python tools/attn_fused_check.py > gpurun_out/attn_check_tma.log 2>&1; echo "tma rc=$?"
grep -E "BAD|time:|ALL OK|FAILED|worst|per-|qkv diff|Error|error" gpurun_out/attn_check_tma.log | head -30
MST_ATTN_TMA=0 timeout 300 python tools/attn_fused_check.py > gpurun_out/attn_check_notma.log 2>&1; echo "notma rc=$?"
grep -E "BAD|ALL OK|FAILED|Error|error" gpurun_out/attn_check_notma.log | head -20
MST_LIB_PATH=mastermetastyletransfer_b200/libmst_b200_prof.so timeout 120 python tools/attn_prof.py 2>&1 | grep -E "af prof|----" | awk 'NR%2==1 || /----/' | head -12
