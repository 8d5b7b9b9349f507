mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -x -k "attention" 2>&1 | tail -3)
(timeout 400 python -m pytest tests/test_gpu_widen_plumbing.py tests/test_gpu_refattn.py -m gpu -q --no-header 2>&1 | tail -3)
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg4_v2.json 2> gpurun_out/c33.err; tail -3 gpurun_out/c33.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_cfg4_v2.json').read().strip().splitlines()[-1]);print('cfg4',d['value'],d['unit'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['clocks']['sm_mhz'])"
