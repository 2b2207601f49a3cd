#!/bin/bash
# scratch driver for one gpurun call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ln
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "layernorm or ln" > gpurun_out/ln/pytest_ln.log 2>&1
echo "pytest ln exit=$?"; tail -n 5 gpurun_out/ln/pytest_ln.log
for v in new prev new prev; do
  if [ $v = prev ]; then export CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_prev.so; else unset CLIPA_B200_LIB; fi
  echo "--- $v"; timeout 300 python tools/prof_ln.py 2>&1 | grep -E "PERF|Error|error" | tee -a gpurun_out/ln/perf_$v.log
done
unset CLIPA_B200_LIB
timeout 600 python bench.py --global-batch 4096 --micro-batch 4096 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ln/bench_new.json 2> gpurun_out/ln/bench_new.err
echo "bench exit=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/ln/bench_new.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ln/pytest_all.log 2>&1
echo "pytest all exit=$?"; tail -n 5 gpurun_out/ln/pytest_all.log
