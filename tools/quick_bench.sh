#!/bin/bash
# fast iteration loop for gpurun: forward bench only (no CPU baseline, no training), per-launch detail into gpurun_out/
tag=${1:-q}
MST_BENCH_DETAIL=gpurun_out/detail_$tag.txt timeout 300 python bench.py --train-steps 0 --cpu-baseline 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log
tail -c 400 gpurun_out/bench_${tag}_err.log
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("img/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 4))
for k, v in d["kernel_families"].items():
    print(f"  {k:26s} {v['launches']:3d} {v['ms']:8.4f} ms {v['share']:.3f} {v['tflops']} {v['gbs']}")
PY
