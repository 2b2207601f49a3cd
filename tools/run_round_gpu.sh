cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
bash tools/gpu_probe.sh gemm_kk gemm_tails gemm_kmn gemm_mnmn gemm_epi loss 2>&1 | grep -E "FAIL|GROUP|PERF|exit=|rror" | head -40
CLIPA_GEMM_MODE=2 bash tools/gpu_probe.sh gemm_kk gemm_tails gemm_kmn gemm_mnmn gemm_epi gemm_perf 2>&1 | grep -E "FAIL|GROUP|PERF|exit=|rror" | head -40
python bench.py --global-batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --op-table gpurun_out/op_table_gb4096.json 2>&1 | tail -1 | tee gpurun_out/bench_l14_gb4096_plain.log
