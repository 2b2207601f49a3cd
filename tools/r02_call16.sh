#!/bin/bash
# attention workers: warp-uniform chunk classification, per-chunk masks + packed fp32 in the flash forward,
# delta cache in the flash backward: parity probe, kernel tests, perf probe, ncu of the L=257 kernels
cd "$(dirname "$0")/.."
out=gpurun_out/r02c16
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 300 python tools/probe_flash.py check > $out/probe_flash_check.log 2>&1; grep -E "FAIL|GROUP|rror" $out/probe_flash_check.log | head -20
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn or attention or flash" > $out/pytest_attn.log 2>&1; echo "pytest exit=$?"; tail -n 3 $out/pytest_attn.log | cut -c1-200
timeout 300 python tools/probe_flash.py perf > $out/probe_flash_perf.log 2>&1; grep -E "PERF" $out/probe_flash_perf.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash -s 2 -c 2 -o $out/flash_l257 python tools/prof_flash.py 256 257 16 64 > $out/ncu_flash.log 2>&1; echo "ncu flash exit=$?"
true
