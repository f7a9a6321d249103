#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ab2.sh default tfirst nofence
unset FFN_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu.log
FFN_B200_DEBUG=128 timeout 300 python bench.py --skip-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; grep -o '"value": [0-9.]*' gpurun_out/bench_quick.json | head -3; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_quick.json
FFN_B200_DEBUG=128 timeout 300 python tools/profile_e2e.py 2>&1 | head -24
