#!/bin/bash
# 2-GPU sanity of what the driver's scaling run does at N=2: GradCache chunks of 2048 with kept MLP activations + NCCL
cd "$(dirname "$0")/.."
out=gpurun_out/r02c23
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/ddp_check.py 2>&1 | grep -E "ddp_check|rror" | tee $out/ddp_check_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > $out/bench_2gpu.json 2> $out/bench_2gpu.err; echo "bench 2gpu exit=$?"; tail -n 1 $out/bench_2gpu.json | cut -c1-900
tail -n 5 $out/bench_2gpu.err | grep -i -E "error|Traceback|memory"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > $out/bench_2gpu_reference.json 2> $out/bench_2gpu_reference.err; echo "reference arm 2gpu exit=$?"; tail -n 1 $out/bench_2gpu_reference.json | cut -c1-300
true
