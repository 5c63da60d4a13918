#!/bin/bash
# kernel + path parity tests, then the quick forward bench; short summary on stdout, full logs in gpurun_out/
tag=${1:-c}
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py tests/test_zz_gpu_alternates.py tests/test_zz_gpu_resume.py -x -q -m gpu > gpurun_out/pytest_$tag.log 2>&1
tail -3 gpurun_out/pytest_$tag.log
tools/quick_bench.sh $tag 2>/dev/null | head -12
