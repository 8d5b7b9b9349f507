timeout 300 python tools/gpu_jit_debug.py 2>&1 | tail -30
