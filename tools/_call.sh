(timeout 600 python -m pytest tests/test_gpu_jit.py -m gpu -q --no-header 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-300 | head -20)
timeout 300 python bench.py --config 6 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg6.json 2> gpurun_out/c37.err; tail -3 gpurun_out/c37.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_cfg6.json').read().strip().splitlines()[-1]);print('cfg6',d['value'],d['unit'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['clocks']['sm_mhz'])"
