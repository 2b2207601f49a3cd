#!/bin/bash
# flash kernels: producer warps moved to SM sub-partitions 2/3 (12-warp CTA) vs the default role layout, one box
cd "$(dirname "$0")/.."
out=gpurun_out/r02c27
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_roles1.so timeout 300 python tools/probe_flash.py check > $out/probe_flash_check_roles1.log 2>&1; grep -E "FAIL|GROUP|rror" $out/probe_flash_check_roles1.log | head -20
for v in b200 roles1 b200 roles1; do
  echo "== lib $v" | tee -a $out/ab.log
  CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_$v.so timeout 200 python tools/probe_flash.py perf 2>&1 | grep PERF | sed 's/  mma.sync.*//' | tee -a $out/ab.log
done
true
