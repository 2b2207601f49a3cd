#!/bin/bash
# scratch driver for one gpurun call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/gelu
export PYTHONPATH="$PWD:$PYTHONPATH"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gelu/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/gelu/pytest_kernels.log 2>&1
echo "pytest kernels exit=$?"; tail -n 5 gpurun_out/gelu/pytest_kernels.log
for v in new prev new prev; do
  if [ $v = prev ]; then export CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_prev.so; else unset CLIPA_B200_LIB; fi
  echo "--- $v"; timeout 300 python tools/gpu_probe.py gemm_epi_perf 2>&1 | grep -E "PERF|Error|error" | tee -a gpurun_out/gelu/epi_perf_$v.log
done
for v in new prev; do
  if [ $v = prev ]; then export CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_prev.so; else unset CLIPA_B200_LIB; fi
  timeout 600 python bench.py --global-batch 4096 --micro-batch 4096 --steps 6 --warmup 3 --no-cpu-baseline --op-table gpurun_out/gelu/op_table_$v.json > gpurun_out/gelu/bench_$v.json 2> gpurun_out/gelu/bench_$v.err
  echo "bench $v exit=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/gelu/bench_$v.json').read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'], d['clocks'])"
done
unset CLIPA_B200_LIB
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/gelu/pytest_model.log 2>&1
echo "pytest model exit=$?"; tail -n 5 gpurun_out/gelu/pytest_model.log
