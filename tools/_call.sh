(timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header --tb=short 2>&1 | grep -E "^E  |passed|failed" | cut -c1-200 | head)
