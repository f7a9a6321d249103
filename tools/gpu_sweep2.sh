#!/bin/bash
mkdir -p gpurun_out
export FFN_B200_LIB=$PWD/variants/libffn_b200_c5.so
timeout 300 python tools/chain_sweep.py 250 32 0 256 512 288 544 > gpurun_out/sweep2_c5.log 2>&1; echo "sweep rc $?"
grep -h "segment_all_chains" gpurun_out/sweep2_c5.log | python -c "
import sys, json
for ln in sys.stdin:
  d = json.loads(ln)
  if 'debug' not in d: print(d); continue
  sp = d['spec']
  print(d['segment_all_chains'], d['debug'], d['steps_per_s'], d['all_equal'], d['ctr_diff'], sp['early_runs'], sp['early_runs_discarded'], sp['steps_discarded'], sp['rounds'], sp['chain_rounds_free'], sp['chain_rounds_waiting'])
"
