#!/bin/bash
# round-2 GPU call 20: the driver's own commands (default bench, reference arm), launch list + ncu of the final kernels
cd "$(dirname "$0")/.."
out=gpurun_out/r02c20
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench default exit=$?"; tail -n 1 $out/bench_default.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "bench reference exit=$?"; tail -n 1 $out/bench_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file $out/launches.csv python bench.py --global-batch 4096 --micro-batch 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline --no-e2e > $out/bench_under_ncu.log 2>&1; echo "launch list exit=$?"
python tools/summarize_launches.py $out/launches.csv > $out/launches_summary.txt 2>&1; head -n 16 $out/launches_summary.txt; gzip -f $out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 7 -c 7 -o $out/gemm python tools/prof_gemm.py 4096 > $out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash -s 2 -c 2 -o $out/flash_l257 python tools/prof_flash.py 256 257 16 64 > $out/ncu_flash.log 2>&1; echo "ncu flash exit=$?"
true
