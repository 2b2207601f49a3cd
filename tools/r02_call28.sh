#!/bin/bash
# last check of the committed state: full GPU suite + smoke + flash parity probe
cd "$(dirname "$0")/.."
out=gpurun_out/r02c28
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 2 $out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 1 $out/smoke.log | cut -c1-200
timeout 300 python tools/probe_flash.py check > $out/probe_flash_check.log 2>&1; grep -E "FAIL|GROUP|rror" $out/probe_flash_check.log | head
true
