mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_production_shapes.py tests/test_gpu_config2_full.py -m gpu -q --no-header -x 2>&1 | tail -3) > gpurun_out/c22_tests.txt 2>&1
(JG_FUSE_GN=1 timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_production_shapes.py tests/test_gpu_config2_full.py tests/test_gpu_palette.py -m gpu -q --no-header -x 2>&1 | tail -3) >> gpurun_out/c22_tests.txt 2>&1
for s in "32 256 256 128 128 3" "32 128 128 128 128 3" "32 128 128 256 256 3"; do timeout 120 python tools/gpu_conv_one.py $s fwdres 10 2>&1 | tail -1; done > gpurun_out/c22_conv.txt 2>&1
for m in fwd dgradgn; do for s in "32 256 256 64 64 3" "32 256 256 128 64 3"; do timeout 120 python tools/gpu_conv_one.py $s $m 10 2>&1 | tail -1; done; done >> gpurun_out/c22_conv.txt 2>&1
for e in stats 1 stats 1; do JG_FUSE_GN=$e timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('fuse=$e',d['value'],d['ms_per_step'],d['clocks']['sm_mhz'])"; done > gpurun_out/c22_bench.txt 2>&1
cat gpurun_out/c22_tests.txt gpurun_out/c22_conv.txt gpurun_out/c22_bench.txt
