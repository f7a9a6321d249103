#!/bin/bash
# A/B of the product library against variants/libffn_b200_<name>.so: roles profile (predict batch 48), then GPU tests.
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset FFN_B200_LIB; else export FFN_B200_LIB=$PWD/variants/libffn_b200_$v.so; fi
  echo "== $v"
  timeout 200 python tools/profile_roles.py 2>&1 | grep -E '^\{|rror' | cut -c1-420
done
unset FFN_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 100 python tools/trace_tiles.py 3 > gpurun_out/trace3_win.log 2>&1; tail -3 gpurun_out/trace3_win.log
FFN_B200_TMAP=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "predict or golden" 2>&1 | tail -2
timeout 300 python bench.py --skip-extras > gpurun_out/bench_win.json 2> gpurun_out/bench_win.err; cut -c1-300 gpurun_out/bench_win.json; grep -o '"device_only": {[^}]*}' gpurun_out/bench_win.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_win.json
