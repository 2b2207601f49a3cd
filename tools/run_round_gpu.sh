cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --op-table gpurun_out/op_table_gb32k.json 2>&1 | tail -3 | tee gpurun_out/bench_l14_gb32k.log
python bench.py --global-batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --op-table gpurun_out/op_table_gb4096.json 2>&1 | tail -3 | tee gpurun_out/bench_l14_gb4096_plain.log
nvidia-smi --query-gpu=memory.used --format=csv
