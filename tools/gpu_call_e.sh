#!/bin/bash
# uint8 boundary tests + path tests (host entry points) + bench line
tag=${1:-r2h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "patch_embed or images or conv3x3 or mlp or instnorm or instance" > gpurun_out/pytest_k_$tag.log 2>&1
echo "kernel gpu tests: $(tail -1 gpurun_out/pytest_k_$tag.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_k_$tag.log | head -20
timeout 900 python -m pytest tests/test_gpu_path.py -x -q -m gpu > gpurun_out/pytest_path_$tag.log 2>&1
echo "path gpu tests: $(tail -1 gpurun_out/pytest_path_$tag.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_path_$tag.log | head -20
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${tag}_err.log
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print(json.dumps(d["summary"]))
print("e2e", json.dumps(d["e2e"]))
PY
