#!/bin/bash
# round-2 GPU call 19: full validation + shard benches after the recompute-epilogue and attention-worker changes
cd "$(dirname "$0")/.."
out=gpurun_out/r02c19
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 6 $out/pytest_gpu.log | cut -c1-200
cp gpurun_out/parity_report.jsonl $out/parity_report.jsonl 2>/dev/null
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 2 $out/smoke.log | cut -c1-200
for spec in "vitl14_i81_t16_gb32k 4096 1024" "vitb16_i64_t16_gb16k 2048 2048" "vitl14_i256_t32_gb16k 2048 512" "vith14_i36_t8_gb64k 8192 2048"; do
  set -- $spec
  timeout 900 python bench.py --workload $1 --global-batch $2 --micro-batch $2 --steps 4 --warmup 3 --no-cpu-baseline --library-batch $3 --op-table $out/op_table_$1.json > $out/library_vs_ours_$1.json 2> $out/bench_$1.err
  echo "bench $1 exit=$?"; tail -n 1 $out/library_vs_ours_$1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); lb=d.get('library_baseline',{}); print('$1', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms mfu', round(d['config']['model_flops_utilization'],3), 'hbm', d['config']['peak_hbm_gb'], d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1), 'library', lb.get('value'), lb.get('unavailable'), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['frac_algorithmic'],3))"
  tail -n 3 $out/bench_$1.err | grep -i -E "error|Traceback"
done
true
