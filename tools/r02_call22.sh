#!/bin/bash
# (rerun after the grad-mode fix) keep-MLP activation level: policy test, N=1 global-batch-32k A/B (micro-batch 4096 recompute vs 2048 keep), shard check
cd "$(dirname "$0")/.."
out=gpurun_out/r02c22
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "activation_policy or trainstep or gradcache or chunk" > $out/pytest_policy.log 2>&1; echo "pytest exit=$?"; tail -n 3 $out/pytest_policy.log | cut -c1-200
run() { # tag, args...
  tag=$1; shift
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline --no-e2e "$@" > $out/bench_$tag.json 2> $out/bench_$tag.err
  echo "$tag exit=$?"; tail -n 1 $out/bench_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']; print('   ', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],1), 'ms mfu', round(c['model_flops_utilization'],3), 'hbm', c['peak_hbm_gb'], d['clocks']['sm_mhz'], 'micro', c['micro_batch'], 'policy', c.get('activation_policy_vision(save_ln,drop_o,keep_mlp_blocks)'), c.get('activation_policy_text(save_ln,drop_o,keep_mlp_blocks)'), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['frac_algorithmic'],3))"
  tail -n 3 $out/bench_$tag.err | grep -i -E "error|Traceback|memory"
}
run gb32k_mb4096_recompute --micro-batch 4096 --keep-mlp 0
run gb32k_mb2048_keep --micro-batch 2048
run gb32k_mb4096_auto --micro-batch 4096
run shard4096_auto --global-batch 4096
run shard4096_recompute --global-batch 4096 --keep-mlp 0
true
