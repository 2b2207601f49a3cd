cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
bash tools/gpu_probe.sh gemm_epi 2>&1 | grep -E "FAIL|GROUP|exit=|rror" | head
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --global-batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --op-table gpurun_out/op_table_gb4096.json 2>&1 | tail -1 | tee gpurun_out/bench_l14_gb4096_plain.log
python bench.py --steps 3 --warmup 3 --op-table gpurun_out/op_table_gb32k.json 2>&1 | tail -1 | tee gpurun_out/bench_l14_gb32k.log
