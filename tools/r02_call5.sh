#!/bin/bash
# round-2 GPU call 5: packed flash tiles + 128-row fwd tiles; CustomTextCLIP / clip-on-device / state_dict plumbing; benches
cd "$(dirname "$0")/.."
out=gpurun_out/r02c5
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 300 python tools/probe_flash.py check > $out/probe_flash_check.log 2>&1; echo "probe check exit=$?"; grep -E "FAIL|GROUP|tile|sample|rror" $out/probe_flash_check.log | head -60
timeout 300 python tools/probe_flash.py perf > $out/probe_flash_perf.log 2>&1; echo "probe perf exit=$?"; grep -E "PERF|rror" $out/probe_flash_perf.log
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 12 $out/pytest_gpu.log
for spec in "vith14_i36_t8_gb64k 8192" "vitl14_i256_t32_gb16k 2048"; do
  set -- $spec
  timeout 500 python bench.py --workload $1 --global-batch $2 --micro-batch $2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline --op-table $out/op_table_$1.json > $out/bench_shard_$1.json 2> $out/bench_shard_$1.err
  echo "bench $1 exit=$?"; tail -n 1 $out/bench_shard_$1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms  mfu', round(d['config']['model_flops_utilization'],3), 'hbm', d['config']['peak_hbm_gb'], d['clocks']['sm_mhz'], d['roofline']['frac'], d['roofline']['frac_algorithmic'])"
  tail -n 5 $out/bench_shard_$1.err
done
true
