#!/bin/bash
# flash backward with dQ kept in TMEM across key tiles (A/B against the scratch path), forward mask mode 2 vs 0
cd "$(dirname "$0")/.."
out=gpurun_out/r02c18
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 300 python tools/probe_flash.py check > $out/probe_flash_check.log 2>&1; grep -E "FAIL|GROUP|rror" $out/probe_flash_check.log | head -20
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn or attention or flash" > $out/pytest_attn.log 2>&1; echo "pytest exit=$?"; tail -n 3 $out/pytest_attn.log | cut -c1-200
for v in "b200 1" "b200 0" "mask2 1" "old 1" "b200 1"; do
  set -- $v
  echo "== lib $1 dq_tmem=$2" | tee -a $out/ab.log
  CLIPA_FLASH_DQ_TMEM=$2 CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_$1.so timeout 200 python tools/probe_flash.py perf 2>&1 | grep PERF | sed 's/  mma.sync.*//' | tee -a $out/ab.log
done
true
