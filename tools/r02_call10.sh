#!/bin/bash
# round-2 GPU call 10 (8 GPUs): full-global-batch bench lines of the headline config and BASELINE configs 2, 4, 5;
# overlap A/B on the headline; 1-GPU shard-size run on the same node
cd "$(dirname "$0")/.."
out=gpurun_out/r02c10
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
port=29800
run8() {  # tag workload extra-args...
  tag=$1; wl=$2; shift 2
  port=$((port + 1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --workload $wl --steps 4 --warmup 3 "$@" > $out/bench_8gpu_${wl}${tag}.json 2> $out/bench_8gpu_${wl}${tag}.err
  rc=$?
  echo "8gpu $wl$tag exit=$rc"; tail -n 1 $out/bench_8gpu_${wl}${tag}.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('  ', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms mfu', round(d['config']['model_flops_utilization'],3), 'hbm', d['config']['peak_hbm_gb'], d['clocks']['sm_mhz'], 'e2e', (d.get('e2e') or {}).get('value'), d['config']['schedule'][:20])
except Exception as e: print('  no line', e)"
  [ $rc -ne 0 ] && tail -n 4 $out/bench_8gpu_${wl}${tag}.err
  return $rc
}
run8 "" vitl14_i81_t16_gb32k
CLIPA_OVERLAP=0 run8 _overlap_off vitl14_i81_t16_gb32k --no-e2e
run8 "" vitb16_i64_t16_gb16k
run8 "" vitl14_i256_t32_gb16k
run8 "" vith14_i36_t8_gb64k --micro-batch 8192 || run8 _gradcache vith14_i36_t8_gb64k --micro-batch 4096
timeout 600 python bench.py --global-batch 4096 --micro-batch 4096 --steps 4 --warmup 3 --no-cpu-baseline --no-library-baseline > $out/bench_1gpu_shard_same_node.json 2> $out/bench_1gpu_shard.err; echo "1gpu shard exit=$?"
tail -n 1 $out/bench_1gpu_shard_same_node.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'])"
true
