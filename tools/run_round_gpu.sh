cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu_probe.sh gemm_epi_perf 2>&1 | grep -E "PERF|exit=|rror"
python bench.py --global-batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --op-table gpurun_out/op_table_gb4096.json > gpurun_out/bench_l14_gb4096_plain.log 2> gpurun_out/bench_l14_gb4096_plain.err; tail -1 gpurun_out/bench_l14_gb4096_plain.log | cut -c1-300; tail -3 gpurun_out/bench_l14_gb4096_plain.err
