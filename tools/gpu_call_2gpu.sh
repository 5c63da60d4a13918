#!/bin/bash
# 2-GPU box: the NCCL correctness tests, the fixed MLP kernel tests, and the bench line at N=2 (replica self-check)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -q -m gpu -k "multi or nccl or ranks or mlp" > gpurun_out/pytest_2gpu.log 2>&1
echo "2-gpu + mlp tests: $(tail -1 gpurun_out/pytest_2gpu.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_2gpu.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2_err.log; echo "bench N=2 rc=$?"
tail -3 gpurun_out/bench_n2_err.log
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
    print(json.dumps(d["summary"]))
    print("train", json.dumps({k: (v["ms_per_step"] if isinstance(v, dict) else v) for k, v in d["training"].items()}))
except Exception as e:
    print("no bench line:", e)
PY
