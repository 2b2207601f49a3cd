#!/bin/bash
# round-2 GPU call 8: fp32 parity path, LN-fused bias gradients, full GPU suite, shard bench
cd "$(dirname "$0")/.."
out=gpurun_out/r02c8
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 25 $out/pytest_gpu.log
cp gpurun_out/parity_report.jsonl $out/parity_report.jsonl 2>/dev/null
timeout 900 python bench.py --global-batch 4096 --micro-batch 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline --op-table $out/op_table_vitl14_gb4096.json > $out/bench_gb4096.json 2> $out/bench_gb4096.err; echo "bench exit=$?"; tail -n 3 $out/bench_gb4096.err
tail -n 1 $out/bench_gb4096.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms mfu', round(d['config']['model_flops_utilization'],3), d['clocks']['sm_mhz'], 'e2e', d.get('e2e',{}).get('value'), d['roofline']['frac'], d['roofline']['frac_algorithmic'], d['gpu_launches'])"
true
