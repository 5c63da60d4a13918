#!/bin/bash
# experiment: MMA-issue variants of conv_rows_kernel (timing only; mode 5 gives wrong results by design)
for m in 0 1 2 5; do echo "mode $m"; MST_ROWS_MODE=$m timeout 100 python tools/gemm_bench.py rows 2>&1 | head -5; done
