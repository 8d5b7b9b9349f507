(timeout 600 python -m pytest tests/test_gpu_cond.py tests/test_gpu_multi.py tests/test_gpu_palette.py -m gpu -q --no-header 2>&1 | grep -E "^E  |passed|failed|Error|error" | cut -c1-300 | head -30)
