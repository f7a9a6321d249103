#!/bin/bash
# One-GPU evidence run of the product build: GPU tests, smoke, the bench line, scheduler stress sweep (chains vs the
# sequential run under the FFN_B200_DEBUG knobs), the ncu launch list of the bench command.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; cat gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
SWEEP_CHAINS=3,4 timeout 300 python tools/chain_sweep.py 250 0 32 256 288 > gpurun_out/sweep_final.log 2>&1; echo "sweep rc $?"
grep -h "predict_chains\|segment_all_chains" gpurun_out/sweep_final.log | cut -c1-330
if [ -z "$SKIP_NCU" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv \
  python bench.py --skip-extras --skip-e2e > gpurun_out/launch_bench.log 2>&1; echo "launch list rc $?"; tail -2 gpurun_out/launch_bench.log | cut -c1-300
fi
