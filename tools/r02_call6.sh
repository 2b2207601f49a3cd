#!/bin/bash
# round-2 GPU call 6: tower heads/tails kernels, full GPU suite, smoke, headline bench line with library baseline
cd "$(dirname "$0")/.."
out=gpurun_out/r02c6
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 25 $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 3 $out/smoke.log
timeout 900 python bench.py --global-batch 4096 --micro-batch 4096 --steps 4 --warmup 3 --op-table $out/op_table_vitl14_gb4096.json > $out/bench_gb4096.json 2> $out/bench_gb4096.err; echo "bench exit=$?"; tail -n 3 $out/bench_gb4096.err
tail -n 1 $out/bench_gb4096.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms mfu', round(d['config']['model_flops_utilization'],3), d['clocks'], 'e2e', d.get('e2e',{}).get('value'), 'lib', d.get('library_baseline'), 'cpu', d.get('cpu_baseline',{}).get('value'), d['roofline']['frac'], d['roofline']['frac_algorithmic'], d['gpu_launches'])"
true
