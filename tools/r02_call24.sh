#!/bin/bash
# 8-GPU headline line with the final kernels (default flags, as the driver's scaling run launches it)
cd "$(dirname "$0")/.."
out=gpurun_out/r02c24
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline > $out/bench_8gpu.json 2> $out/bench_8gpu.err; echo "bench 8gpu exit=$?"; tail -n 1 $out/bench_8gpu.json | cut -c1-1200
tail -n 5 $out/bench_8gpu.err | grep -i -E "error|Traceback|memory"
true
