#!/bin/bash
# scratch driver for one gpurun call: final validation + profiles of the round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/final2
export PYTHONPATH="$PWD:$PYTHONPATH"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/final2/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/final2/pytest_gpu.log 2>&1
echo "pytest gpu exit=$?"; tail -n 4 gpurun_out/final2/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/final2/smoke.log 2>&1
echo "smoke exit=$?"; tail -n 2 gpurun_out/final2/smoke.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -c 5 -f -o gpurun_out/final2/prof_gemm3 python tools/prof_gemm.py 4096 > gpurun_out/final2/ncu_gemm.log 2>&1
echo "ncu gemm exit=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 2 -f -o gpurun_out/final2/prof_attn3 python tools/prof_attn.py > gpurun_out/final2/ncu_attn.log 2>&1
echo "ncu attn exit=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/final2/launches.csv python bench.py --global-batch 4096 --micro-batch 4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/final2/ncu_launch_bench.log 2>&1
echo "ncu launches exit=$?"; wc -l gpurun_out/final2/launches.csv
gzip -f gpurun_out/final2/launches.csv
timeout 900 python bench.py > gpurun_out/final2/bench_default.json 2> gpurun_out/final2/bench_default.err
echo "bench default exit=$?"; tail -c 600 gpurun_out/final2/bench_default.json
