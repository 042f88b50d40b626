#!/bin/bash
# On the GPU box: repeat the sharded parity tests in fresh processes (each run creates its engines on
# non-blocking streams and renders from a cold start - the condition under which the intermittent
# first-window failure of round 2 showed up).   stress_multigpu.sh [runs] [-k expression]
runs=${1:-50}; expr=${2:-"fused_exchange or cut_path or root_ramp"}
mkdir -p gpurun_out
for i in $(seq 1 $runs); do
  timeout 120 python -m pytest tests/test_multigpu.py -m gpu -q -k "$expr" 2>&1 | grep -E "passed|failed|first diff" | head -3
done | sort | uniq -c | tee gpurun_out/stress_multigpu.txt
