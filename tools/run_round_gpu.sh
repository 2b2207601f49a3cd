cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
sed -i 's/timeout 300 python/timeout 120 python/' tools/gpu_probe.sh
CLIPA_GEMM_MODE=2 bash tools/gpu_probe.sh gemm_kk gemm_tails gemm_kmn gemm_mnmn gemm_epi 2>&1 | grep -E "FAIL|GROUP|PERF|exit=|rror" | head -40
echo "--- perf mode 2 (2-CTA)"; CLIPA_GEMM_MODE=2 bash tools/gpu_probe.sh gemm_perf 2>&1 | grep -E "PERF|exit=|rror"
echo "--- perf mode 1 (1-CTA)"; CLIPA_GEMM_MODE=1 bash tools/gpu_probe.sh gemm_perf 2>&1 | grep -E "PERF|exit=|rror"
