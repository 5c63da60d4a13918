#!/bin/bash
# patch-embedding kernel: tests with the TMA stores and with direct stores, timing of both, then path tests + bench line
tag=${1:-r2q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "patch_embed" > gpurun_out/pytest_pe_$tag.log 2>&1
echo "patch_embed tests (TMA stores): $(tail -1 gpurun_out/pytest_pe_$tag.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_pe_$tag.log | head -10
MST_PATCH_EMBED_TMA=0 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "patch_embed" > gpurun_out/pytest_pe0_$tag.log 2>&1
echo "patch_embed tests (direct stores): $(tail -1 gpurun_out/pytest_pe0_$tag.log)"
timeout 300 python tools/debug/patch_embed_time.py 2>&1 | tail -4
MST_PATCH_EMBED_TMA=0 timeout 300 python tools/debug/patch_embed_time.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_path.py -x -q -m gpu > gpurun_out/pytest_path_$tag.log 2>&1
echo "path gpu tests: $(tail -1 gpurun_out/pytest_path_$tag.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_path_$tag.log | head -10
timeout 900 python bench.py --train-steps 0 --cpu-baseline 0 --config5 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print(json.dumps(d["summary"]))
for k, v in d["kernel_families"].items(): print("  ", k, v)
PY
