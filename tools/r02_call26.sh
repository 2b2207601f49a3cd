#!/bin/bash
# memory-policy safety check: the other BASELINE workloads at N=1 with the automatic GradCache chunk sizes
cd "$(dirname "$0")/.."
out=gpurun_out/r02c26
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
for wl in vitb16_i64_t16_gb16k vitl14_i256_t32_gb16k vith14_i36_t8_gb64k; do
  timeout 600 python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline --no-e2e > $out/bench_$wl.json 2> $out/bench_$wl.err
  echo "$wl exit=$?"; tail -n 1 $out/bench_$wl.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']; print('   ', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],1), 'ms mfu', round(c['model_flops_utilization'],3), 'hbm', c['peak_hbm_gb'], d['clocks']['sm_mhz'], 'micro', c['micro_batch'], 'policy', c.get('activation_policy_vision(save_ln,drop_o,keep_mlp_blocks)'), c.get('activation_policy_text(save_ln,drop_o,keep_mlp_blocks)'))"
  tail -n 3 $out/bench_$wl.err | grep -i -E "error|Traceback|memory"
done
true
