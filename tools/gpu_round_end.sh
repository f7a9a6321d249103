#!/bin/bash
# One-GPU evidence run: GPU tests, chains-vs-sequential on the 250^3 canvas, one `ncu --set full` capture of the
# flood kernel, the launch list and the bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu.log
FFN_B200_DEBUG=64 timeout 300 python tools/chain_debug.py 250 64 > gpurun_out/chain_debug.log 2>&1; echo "chain_debug rc $?"; tail -6 gpurun_out/chain_debug.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ffn_flood_kernel -c 1 -f -o gpurun_out/r02_flood \
  python bench.py --max-seeds 60 --skip-extras --skip-e2e --warmup 1 > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/ncu_full_bench.log
python tools/ncu_summary.py gpurun_out/r02_flood.ncu-rep "ncu --set full --clock-control none, first ffn_flood_kernel launch of: bench.py --max-seeds 60 --skip-extras --skip-e2e --warmup 1" > gpurun_out/r02_ncu_full_flood.txt 2>&1
ls -la gpurun_out/r02_flood.ncu-rep
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; cat gpurun_out/r02_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv \
  python bench.py --skip-extras > gpurun_out/launch_bench.log 2>&1; echo "launch list rc $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>&1; cat gpurun_out/r02_bench_ref.json
