#!/bin/bash
# round-2 GPU call 1: test suite, library baselines (unmodified reference on torch CUDA kernels), uniform-issue A/B
cd "$(dirname "$0")/.."
out=gpurun_out/r02c1
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $out/gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 4 $out/pytest_gpu.log
for spec in "vitl14_i81_t16_gb32k 2048" "vitb16_i64_t16_gb16k 2048" "vitl14_i256_t32_gb16k 512" "vith14_i36_t8_gb64k 2048"; do
  set -- $spec
  timeout 400 python tools/library_baseline.py --workload $1 --batch $2 --steps 3 --warmup 2 > $out/library_$1.json 2> $out/library_$1.err
  echo "library $1 exit=$?"; tail -n 1 $out/library_$1.json; tail -n 2 $out/library_$1.err | grep -i error
done
UNI=$PWD/clipa_b200/lib/libclipa_b200_uni.so
CLIPA_B200_LIB=$UNI timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > $out/pytest_kernels_uni.log 2>&1
echo "uni kernel tests exit=$?"; tail -n 2 $out/pytest_kernels_uni.log
for v in base uni; do
  if [ $v = uni ]; then export CLIPA_B200_LIB=$UNI; else unset CLIPA_B200_LIB; fi
  timeout 200 python tools/gpu_probe.py gemm_epi_perf 2>&1 | grep -E "PERF|rror" > $out/epi_perf_$v.log
  timeout 200 python tools/gpu_probe.py gemm_perf 2>&1 | grep -E "PERF|rror" > $out/gemm_perf_$v.log
  timeout 200 python tools/prof_attn_text.py 2>&1 | grep -E "PERF|rror" > $out/attn_perf_$v.log
done
for v in base uni base uni; do
  if [ $v = uni ]; then export CLIPA_B200_LIB=$UNI; else unset CLIPA_B200_LIB; fi
  timeout 400 python bench.py --global-batch 4096 --micro-batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e >> $out/bench_$v.json 2>> $out/bench_$v.err
  echo "bench $v exit=$?"; tail -n 1 $out/bench_$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['roofline']['achieved'])"
done
unset CLIPA_B200_LIB
paste -d'\n' $out/epi_perf_base.log $out/epi_perf_uni.log | head -40
