cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
for pol in off on off on; do
python bench.py --global-batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --save-ln $pol 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('save_ln', '$pol', 'ms', round(d['ms_per_step'],1), 'pairs/s', round(d['value'],1), 'peak_hbm_gb', d['config']['peak_hbm_gb'], d['clocks']['sm_mhz'])"
done
