#!/bin/bash
# recompute epilogue with two staging tiles + hoisted activation branch: kernel tests, sustained probe
cd "$(dirname "$0")/.."
out=gpurun_out/r02c15
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or epilogue or deriv or act or block" > $out/pytest_gemm.log 2>&1; echo "pytest exit=$?"; tail -n 3 $out/pytest_gemm.log | cut -c1-200
timeout 300 python tools/probe_gemm_sustained.py 4096 1024 300 2>&1 | grep -E "PERF|rror" | tee $out/gemm_sustained.log
true
