#!/bin/bash
# r2 call B: new parity tests (512^2, batch-32, end-to-end gradient gate, 10-step trajectory, two graphs on one trainer) + the new bench line
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_train.py -q -m gpu -s -k "512 or batch32 or full_training_step or trajectory or two_graphs" > gpurun_out/pytest_new_$tag.log 2>&1
echo "new tests: $(tail -1 gpurun_out/pytest_new_$tag.log)"
grep -E "end-to-end gradient|oracle total|kernel total|worst step|^FAILED|^ERROR|Error|assert " gpurun_out/pytest_new_$tag.log | head -40
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; echo "bench rc=$?"
tail -3 gpurun_out/bench_${tag}_err.log
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print(json.dumps(d["summary"]))
    print("loss_forward", json.dumps(d["loss_forward"]))
    print("gpu_eager", json.dumps(d["gpu_eager_baseline"]))
    print("config5", json.dumps({k: v for k, v in d["config5"].items() if k != "e2e"}))
except Exception as e:
    print("no bench line:", e)
PY
