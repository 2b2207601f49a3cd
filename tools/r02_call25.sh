#!/bin/bash
# final verification of the round: full GPU suite, smoke, the driver's default bench command
cd "$(dirname "$0")/.."
out=gpurun_out/r02c25
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 3 $out/pytest_gpu.log | cut -c1-200
cp gpurun_out/parity_report.jsonl $out/parity_report.jsonl 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 1 $out/smoke.log | cut -c1-200
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench default exit=$?"; tail -n 1 $out/bench_default.json | cut -c1-1500
tail -n 3 $out/bench_default.err | grep -i -E "error|Traceback|memory"
true
