#!/bin/bash
# full GPU suite, then an A/B of the GEMM's TMA output stores inside the graphed step on the same box
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all_$tag.log 2>&1
echo "all gpu tests: $(tail -1 gpurun_out/pytest_all_$tag.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_all_$tag.log | head -10
for v in 1 0 1 0; do
  MST_GEMM_TMA_OUT=$v timeout 200 python bench.py --train-steps 0 --cpu-baseline 0 --config5 0 2>/dev/null > gpurun_out/bench_${tag}_$v.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${tag}_$v.json').read().strip().splitlines()[-1]); print('TMA_OUT=$v', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['kernel_families']['gemm_tc_kernel'])"
done
