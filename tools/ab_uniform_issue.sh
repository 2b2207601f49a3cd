#!/bin/bash
# A/B of the warp-uniform tcgen05 issue mode (-DCLIPA_UNIFORM_ISSUE=1, see csrc/ptx.cuh) against the default
# single-lane issuer, on ONE box.  Build the variant library BEFORE the gpurun call (no nvcc time on the GPU box):
#   CLIPA_B200_NVCC_FLAGS="-DCLIPA_UNIFORM_ISSUE=1" CLIPA_B200_LIB_NAME=libclipa_b200_uni.so python clipa_b200/build.py
#   gpurun --timeout 900 -- 'bash tools/ab_uniform_issue.sh'
# Every step runs under `timeout`: a wrong barrier protocol in the variant would hang, not crash.
cd "$(dirname "$0")/.."
out=gpurun_out/uniform_issue
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
UNI=$PWD/clipa_b200/lib/libclipa_b200_uni.so
[ -f "$UNI" ] || { echo "missing $UNI (build it first)"; exit 1; }
CLIPA_B200_LIB=$UNI timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > $out/pytest_kernels_uni.log 2>&1
echo "variant kernel tests exit=$?"; tail -n 3 $out/pytest_kernels_uni.log
for v in uni base uni base; do
  if [ $v = uni ]; then export CLIPA_B200_LIB=$UNI; else unset CLIPA_B200_LIB; fi
  echo "--- $v"
  timeout 200 python tools/gpu_probe.py gemm_epi_perf 2>&1 | grep -E "PERF|rror" | tee -a $out/epi_perf_$v.log
  timeout 200 python tools/gpu_probe.py gemm_perf 2>&1 | grep -E "PERF|rror" | tee -a $out/gemm_perf_$v.log
  timeout 200 python tools/prof_attn_text.py 2>&1 | grep -E "PERF|rror" | tee -a $out/attn_perf_$v.log
done
for v in uni base; do
  if [ $v = uni ]; then export CLIPA_B200_LIB=$UNI; else unset CLIPA_B200_LIB; fi
  timeout 400 python bench.py --global-batch 4096 --micro-batch 4096 --steps 6 --warmup 3 --no-cpu-baseline > $out/bench_$v.json 2> $out/bench_$v.err
  echo "bench $v exit=$?"; python -c "
import json
d=json.loads(open('$out/bench_$v.json').read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'], d['clocks'])"
done
unset CLIPA_B200_LIB
CLIPA_B200_LIB=$UNI timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q > $out/pytest_model_uni.log 2>&1
echo "variant model tests exit=$?"; tail -n 3 $out/pytest_model_uni.log
