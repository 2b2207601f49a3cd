#!/bin/bash
# round-2 GPU call 9: ncu evidence -- launch list of a step, --set full of the GEMMs, attention kernels, LayerNorm backward
cd "$(dirname "$0")/.."
out=gpurun_out/r02c9
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -n 4 $out/pytest_gpu.log
timeout 300 python tools/probe_flash.py check > $out/probe_flash_check.log 2>&1; grep -E "FAIL|GROUP" $out/probe_flash_check.log | head -20
timeout 300 python tools/probe_flash.py perf > $out/probe_flash_perf.log 2>&1; grep -E "PERF" $out/probe_flash_perf.log
timeout 200 python tools/prof_ln.py > $out/ln_perf.log 2>&1; grep PERF $out/ln_perf.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file $out/launches.csv python bench.py --global-batch 4096 --micro-batch 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline --no-e2e > $out/bench_under_ncu.log 2>&1; echo "launch list exit=$?"
python tools/summarize_launches.py $out/launches.csv > $out/launches_summary.txt 2>&1; head -n 30 $out/launches_summary.txt; gzip -f $out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 6 -c 6 -o $out/gemm python tools/prof_gemm.py 4096 > $out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 2 -o $out/attn_l82 python tools/prof_attn.py > $out/ncu_attn.log 2>&1; echo "ncu attn exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash -s 2 -c 2 -o $out/flash_l257 python tools/prof_flash.py 256 257 16 64 > $out/ncu_flash.log 2>&1; echo "ncu flash exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash -s 2 -c 2 -o $out/flash_l37_hd80 python tools/prof_flash.py 2048 37 16 80 > $out/ncu_flash80.log 2>&1; echo "ncu flash80 exit=$?"
ls -la $out
true
