cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
bash tools/gpu_probe.sh attn 2>&1 | grep -E "FAIL|GROUP|PERF|exit=|rror" | head -20
ncu --set full --clock-control none --import-source on -k regex:attn -s 4 -c 2 -o gpurun_out/prof_attn -f python tools/prof_attn.py > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 5 -o gpurun_out/prof_gemm -f python tools/prof_gemm.py > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
python bench.py --global-batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --op-table gpurun_out/op_table_gb4096.json 2>&1 | tail -1 | tee gpurun_out/bench_l14_gb4096_plain.log
ls -la gpurun_out/*.ncu-rep
