#!/bin/bash
# scratch driver for one gpurun call
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/pack
export PYTHONPATH="$PWD:$PYTHONPATH"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/pack/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" > gpurun_out/pack/pytest_attn.log 2>&1
echo "pytest attn exit=$?"; tail -n 5 gpurun_out/pack/pytest_attn.log
echo "--- new"; timeout 300 python tools/prof_attn_text.py 2>&1 | tee gpurun_out/pack/perf_new.log
echo "--- old"; CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_oldattn.so timeout 300 python tools/prof_attn_text.py 2>&1 | tee gpurun_out/pack/perf_old.log
for v in new old; do
  if [ $v = old ]; then export CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_oldattn.so; else unset CLIPA_B200_LIB; fi
  timeout 600 python bench.py --global-batch 4096 --micro-batch 4096 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/pack/bench_$v.json 2> gpurun_out/pack/bench_$v.err
  echo "bench $v exit=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/pack/bench_$v.json').read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'], d['clocks'])"
done
unset CLIPA_B200_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pack/pytest_all.log 2>&1
echo "pytest all exit=$?"; tail -n 5 gpurun_out/pack/pytest_all.log
