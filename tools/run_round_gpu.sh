#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/attnf2
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > gpurun_out/attnf2/pytest_attn.log 2>&1
echo "pytest attn exit=$?"; tail -n 3 gpurun_out/attnf2/pytest_attn.log
for v in new prev new prev; do
  if [ $v = prev ]; then export CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_b200_prev.so; else unset CLIPA_B200_LIB; fi
  echo "--- $v"; timeout 300 python tools/prof_attn_text.py 2>&1 | grep -E "PERF|Error|error" | tee -a gpurun_out/attnf2/perf_$v.log
done
