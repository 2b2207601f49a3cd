cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_8gpu_gb32k.log 2> gpurun_out/bench_8gpu_gb32k.err
tail -1 gpurun_out/bench_8gpu_gb32k.log | cut -c1-400; grep -v Warning gpurun_out/bench_8gpu_gb32k.err | tail -15 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_4gpu_gb32k.log 2> gpurun_out/bench_4gpu_gb32k.err
tail -1 gpurun_out/bench_4gpu_gb32k.log | cut -c1-400; grep -v Warning gpurun_out/bench_4gpu_gb32k.err | tail -5 | cut -c1-300
