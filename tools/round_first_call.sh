#!/bin/bash
# First gpurun call of a round: the B200 tests written at the end of the previous round without a GPU (alternate style-transformer
# configurations, optimiser checkpoint / resume, fast adaptation) FIRST and on their own, then the whole GPU suite, smoke(), and the
# default bench line.  Short summary on stdout, full logs in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round_first_call.sh r2a'
tag=${1:-first}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_alternates.py tests/test_zz_gpu_resume.py -q -m gpu > gpurun_out/pytest_zz_$tag.log 2>&1
echo "zz tests: $(tail -1 gpurun_out/pytest_zz_$tag.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_zz_$tag.log | head -20
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_all_$tag.log 2>&1
echo "all gpu tests: $(tail -1 gpurun_out/pytest_all_$tag.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_$tag.log)"
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_${tag}_err.log; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("img/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
          "train", d.get("training"))
except Exception as e:
    print("no bench line:", e)
PY
