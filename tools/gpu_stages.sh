#!/bin/bash
# Runs each probe stage in its own process (a faulting kernel poisons its CUDA context only).
mkdir -p gpurun_out
for st in "$@"; do
  echo "=== stage $st" >> gpurun_out/stages.log
  timeout 300 python tools/gpu_probe.py $st >> gpurun_out/stages.log 2>&1
  cp gpurun_out/probe.json gpurun_out/probe_$st.json 2>/dev/null
done
