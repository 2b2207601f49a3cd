#!/bin/bash
# after moving the activation-policy arithmetic into a pure function: model tests + two short bench runs
cd "$(dirname "$0")/.."
out=gpurun_out/r02c29
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_reference_caller_gpu.py -m gpu -q > $out/pytest_model.log 2>&1; echo "pytest exit=$?"; tail -n 2 $out/pytest_model.log | cut -c1-200
for a in "--global-batch 4096" "--global-batch 8192"; do
  timeout 300 python bench.py $a --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline --no-e2e 2> $out/err.log | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']; print('   ', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],1), 'ms hbm', c['peak_hbm_gb'], 'micro', c['micro_batch'], c['schedule'][:12], 'policy', c.get('activation_policy_vision(save_ln,drop_o,keep_mlp_blocks)'), c.get('activation_policy_text(save_ln,drop_o,keep_mlp_blocks)'))"
done
true
