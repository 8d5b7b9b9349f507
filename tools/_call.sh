(timeout 600 python -m pytest tests/test_gpu_jit.py -m gpu -q --no-header 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-300 | head -20)
