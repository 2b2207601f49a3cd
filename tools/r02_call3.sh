#!/bin/bash
# round-2 GPU call 3: ncu --set full (source-level) of the flash attention kernels; new same-mode parity tests
cd "$(dirname "$0")/.."
out=gpurun_out/r02c3
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash -s 2 -c 2 -o $out/flash_l257 python tools/prof_flash.py 256 257 16 64 > $out/ncu_flash.log 2>&1; echo "ncu exit=$?"; tail -n 3 $out/ncu_flash.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q > $out/pytest_model.log 2>&1; echo "pytest exit=$?"; tail -n 30 $out/pytest_model.log
cp gpurun_out/parity_report.jsonl $out/parity_report.jsonl 2>/dev/null
