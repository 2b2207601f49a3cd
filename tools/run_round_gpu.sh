cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --workload vitb32_i36_t16_gb256 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_small.log
python bench.py --global-batch 4096 --micro-batch 2048 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_l14_gb4096.log
nvidia-smi --query-gpu=memory.used --format=csv
