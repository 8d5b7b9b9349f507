mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_production_shapes.py tests/test_gpu_config2_full.py -m gpu -q --no-header -x 2>&1 | tail -3) > gpurun_out/c21_tests.txt 2>&1
for e in 0 4; do for s in "32 256 256 128 128 3" "32 128 128 128 128 3" "32 128 128 256 256 3" "32 64 64 256 256 3" "32 32 32 512 512 3"; do JG_DBG_EPI=$e timeout 120 python tools/gpu_conv_one.py $s fwdres 10 2>&1 | tail -1 | sed "s/^/dbg=$e /"; done; done > gpurun_out/c21_fwdres.txt 2>&1
for e in 0 4 0 4; do JG_DBG_EPI=$e timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dbg=$e',d['value'],d['ms_per_step'],d['clocks']['sm_mhz'])"; done > gpurun_out/c21_bench.txt 2>&1
cat gpurun_out/c21_tests.txt gpurun_out/c21_fwdres.txt gpurun_out/c21_bench.txt
