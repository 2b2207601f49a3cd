cd /root/repo
export PYTHONPATH=/root/repo
python tools/diag_adamw.py 2>&1 | grep -E "loss=|worst"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -4
