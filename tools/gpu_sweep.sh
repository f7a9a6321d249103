#!/bin/bash
# One-GPU A/B of the chain-count / accumulator-slot variants (tools/build_variants.py) against the product build,
# all in one call on one box; then where Runner.run's end-to-end time goes.  Outputs under gpurun_out/.
mkdir -p gpurun_out
for v in default c3s2 c4 c5; do
  if [ "$v" = default ]; then unset FFN_B200_LIB; else export FFN_B200_LIB=$PWD/variants/libffn_b200_$v.so; fi
  timeout 200 python tools/chain_sweep.py 250 ${SWEEP_KNOBS:-0 32} > gpurun_out/sweep_$v.log 2>&1; echo "sweep $v rc $?"
  grep -h "predict_chains\|segment_all_chains" gpurun_out/sweep_$v.log | cut -c1-400
done
unset FFN_B200_LIB
timeout 150 python tools/profile_e2e.py > gpurun_out/profile_e2e.log 2>&1; echo "profile_e2e rc $?"; head -60 gpurun_out/profile_e2e.log
