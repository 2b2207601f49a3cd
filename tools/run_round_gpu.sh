cd /root/repo
export PYTHONPATH=/root/repo
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py 2>&1 | grep -E "ddp_check|rror|Assert" | tee gpurun_out/ddp_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --global-batch 8192 --steps 2 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_2gpu_gb8192.log
