#!/bin/bash
# full GPU suite with the fused attention kernel in the engine + bench line
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all_$tag.log 2>&1
echo "all gpu tests: $(tail -1 gpurun_out/pytest_all_$tag.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_all_$tag.log | head -20
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${tag}_err.log
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print(json.dumps(d["summary"]))
    print("roofline", json.dumps(d["roofline"]))
    for k, v in d["kernel_families"].items(): print("  ", k, v)
    print("loss_forward", json.dumps(d["loss_forward"]))
    print("gpu_eager", json.dumps(d["gpu_eager_baseline"]))
    print("config5", json.dumps({k: v for k, v in d["config5"].items() if k != "e2e"}))
    print("train", json.dumps({k: (v["ms_per_step"] if isinstance(v, dict) else v) for k, v in d["training"].items()}))
except Exception as e:
    print("no bench line:", e)
PY
