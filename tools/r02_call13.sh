#!/bin/bash
cd "$(dirname "$0")/.."
out=gpurun_out/r02c13
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 300 python tools/probe_e2e.py vitb16_i64_t16_gb16k 2048 2>&1 | grep -E "PERF|rror" | tee $out/probe_e2e.log
timeout 300 python tools/probe_e2e.py vitl14_i81_t16_gb32k 4096 2>&1 | grep -E "PERF|rror" | tee -a $out/probe_e2e.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_reference_caller_gpu.py -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest exit=$?"; tail -n 3 $out/pytest.log | cut -c1-200
true
