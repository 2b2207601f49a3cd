cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv python bench.py --global-batch 4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt; head -30 gpurun_out/launches_summary.txt
gzip -f gpurun_out/launches.csv; ls -la gpurun_out/launches.csv.gz
