mkdir -p gpurun_out
for tc in 1 0; do for s in "2 256 4 32" "32 1024 16 32" "2 1024 8 64" "2 3072 16 32"; do JG_ATTN_TC=$tc timeout 120 python tools/gpu_attn_bwd_one.py $s 0 10 2>&1 | tail -4; done; done > gpurun_out/c28_attn.txt 2>&1
JG_ATTN_TC=1 timeout 120 python tools/gpu_attn_bwd_one.py 2 1024 8 32 1 5 2>&1 | tail -2 >> gpurun_out/c28_attn.txt
(timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_production_shapes.py tests/test_gpu_palette.py -m gpu -q --no-header -x 2>&1 | tail -3) >> gpurun_out/c28_attn.txt 2>&1
cat gpurun_out/c28_attn.txt
