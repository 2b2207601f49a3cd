#!/bin/bash
# A/B of the flash-forward worker variants (one box): old source, default (chunk kinds + packed fp32 + split loops),
# scalar math, per-element masks, both
cd "$(dirname "$0")/.."
out=gpurun_out/r02c17
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 300 python tools/probe_flash.py check > $out/probe_flash_check.log 2>&1; grep -E "FAIL|GROUP|rror" $out/probe_flash_check.log | head -20
for v in old b200 m1s0 m0s1 m0s0 old b200; do
  echo "== lib $v" | tee -a $out/ab.log
  CLIPA_B200_LIB=$PWD/clipa_b200/lib/libclipa_$v.so timeout 200 python tools/probe_flash.py perf 2>&1 | grep PERF | sed 's/  mma.sync.*//' | tee -a $out/ab.log
done
true
