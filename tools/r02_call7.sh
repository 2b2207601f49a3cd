#!/bin/bash
# round-2 GPU call 7 (2 GPUs): full GPU suite, 2-rank NCCL check (overlapped bucketed all-reduce, GradCache), 2-GPU bench
cd "$(dirname "$0")/.."
out=gpurun_out/r02c7
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 15 $out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/ddp_check.py > $out/ddp_check_2gpu.log 2>&1; echo "ddp_check exit=$?"; grep -E "ddp_check|Error|error" $out/ddp_check_2gpu.log | tail -n 8
for ov in 1 0; do
  CLIPA_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --global-batch 8192 --micro-batch 4096 --steps 4 --warmup 3 --no-e2e > $out/bench_2gpu_gb8192_overlap$ov.json 2> $out/bench_2gpu_overlap$ov.err; echo "bench overlap=$ov exit=$?"
  tail -n 1 $out/bench_2gpu_gb8192_overlap$ov.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('overlap=$ov', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms mfu', round(d['config']['model_flops_utilization'],3), d['clocks']['sm_mhz'])"
done
true
