#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/n8
export PYTHONPATH="$PWD:$PYTHONPATH"
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > gpurun_out/n8/gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/n8/bench_8gpu.json 2> gpurun_out/n8/bench_8gpu.err
echo "bench8 exit=$?"; tail -c 2500 gpurun_out/n8/bench_8gpu.json; tail -n 5 gpurun_out/n8/bench_8gpu.err
