cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
for w in "vitb16_i64_t16_gb16k 2048" "vitl14_i256_t32_gb16k 2048" "vith14_i36_t8_gb64k 8192"; do
set -- $w
python bench.py --workload $1 --global-batch $2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --op-table gpurun_out/op_table_$1.json > gpurun_out/bench_$1.log 2> gpurun_out/bench_$1.err
tail -1 gpurun_out/bench_$1.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'B', d['config']['global_batch'], 'ms', round(d['ms_per_step'],1), 'pairs/s', round(d['value'],1), 'mfu', round(d['config']['model_flops_utilization'],3), 'gemm TF', round(d['roofline']['achieved'],1), 'peak GB', d['config']['peak_hbm_gb'], 'loss', round(d['config']['loss_last'],3))
except Exception as e: print('$1 FAILED', e)"
grep -v Warn gpurun_out/bench_$1.err | tail -3 | cut -c1-300
done
