#!/bin/bash
# sustained epilogue probe at the shard shape + ncu --set full of the seven block GEMMs (incl. the c_fc recompute)
cd "$(dirname "$0")/.."
out=gpurun_out/r02c14
mkdir -p $out
export PYTHONPATH="$PWD:$PYTHONPATH"
timeout 300 python tools/probe_gemm_sustained.py 4096 1024 300 2>&1 | grep -E "PERF|rror" | tee $out/gemm_sustained.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 7 -c 7 -o $out/gemm_full python tools/prof_gemm.py 4096 > $out/ncu.log 2>&1; echo "ncu exit=$?"
timeout 200 torchrun --standalone --nproc-per-node 1 tests/ddp_check.py 2>&1 | grep -E "ddp_check|rror" | head
true
