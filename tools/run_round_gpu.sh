#!/bin/bash
# scratch driver for one gpurun call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/dact
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > gpurun_out/dact/pytest_gemm.log 2>&1
echo "pytest gemm exit=$?"; tail -n 5 gpurun_out/dact/pytest_gemm.log
CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_d52.so timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" > gpurun_out/dact/pytest_gemm_d52.log 2>&1
echo "pytest gemm d52 exit=$?"; tail -n 3 gpurun_out/dact/pytest_gemm_d52.log
for v in new d52 prev new d52 prev; do
  if [ $v = new ]; then unset CLIPA_B200_LIB; else export CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_$v.so; fi
  echo "--- $v"; timeout 300 python tools/gpu_probe.py gemm_epi_perf 2>&1 | grep -E "PERF|Error|error" | grep -E "dgrad|Error|error" | tee -a gpurun_out/dact/epi_perf_$v.log
done
for v in new d52; do
  if [ $v = new ]; then unset CLIPA_B200_LIB; else export CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_$v.so; fi
  timeout 600 python bench.py --global-batch 4096 --micro-batch 4096 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/dact/bench_$v.json 2> gpurun_out/dact/bench_$v.err
  echo "bench $v exit=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/dact/bench_$v.json').read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'], d['clocks'])"
done
unset CLIPA_B200_LIB
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/dact/pytest_model.log 2>&1
echo "pytest model exit=$?"; tail -n 3 gpurun_out/dact/pytest_model.log
