#!/bin/bash
mkdir -p gpurun_out
FFN_B200_DEBUG=128 timeout 600 python bench.py --skip-extras > gpurun_out/probe2_bench.json 2> gpurun_out/probe2_bench.err; tail -5 gpurun_out/probe2_bench.err; cut -c1-400 gpurun_out/probe2_bench.json
FFN_B200_DEBUG=128 timeout 600 python tools/profile_e2e.py > gpurun_out/probe2_e2e.log 2>&1; head -60 gpurun_out/probe2_e2e.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ffn_flood_kernel -s 1 -c 1 -f -o gpurun_out/r02_flood \
  python bench.py --max-seeds 60 --skip-extras --skip-e2e --warmup 1 > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu rc $?"; tail -3 gpurun_out/ncu_full_bench.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/r02_flood.ncu-rep "ncu --set full --clock-control none -k regex:ffn_flood_kernel -s 1 -c 1 (the timed segment_all launch; the first launch is the warm-up) of: bench.py --max-seeds 60 --skip-extras --skip-e2e" > gpurun_out/r02_ncu_full_flood.txt 2>&1
grep "gpu__time_duration.sum\|tensor_cycles\|dram__bytes_read.sum \[\|dram__bytes_write.sum \[" gpurun_out/r02_ncu_full_flood.txt
