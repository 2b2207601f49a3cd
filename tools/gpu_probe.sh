#!/bin/bash
# Runs every probe group in its own process under a timeout; logs go to gpurun_out/probe/.
# usage: tools/gpu_probe.sh [group ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/probe
export PYTHONPATH="$PWD:$PYTHONPATH"
groups="$@"
[ -z "$groups" ] && groups="gemm_kk gemm_tails gemm_kmn gemm_mnmn gemm_epi ln loss attn gemm_perf"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe/gpu.txt 2>&1
for g in $groups; do
  echo "=== $g ===" | tee gpurun_out/probe/$g.log
  timeout 300 python tools/gpu_probe.py $g >> gpurun_out/probe/$g.log 2>&1
  echo "exit=$?" >> gpurun_out/probe/$g.log
  tail -n 40 gpurun_out/probe/$g.log | grep -E "PASS|FAIL|GROUP|PERF|exit=|Error|error" | tail -n 30
done
