mkdir -p gpurun_out
timeout 300 python bench.py --config 6 --steps 10 --warmup 4 > gpurun_out/r02_bench_cfg6_v2.json 2> gpurun_out/c42.err; tail -3 gpurun_out/c42.err | cut -c1-300
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_cfg6_v2.json').read().strip().splitlines()[-1]);print('cfg6',d['value'],d['unit'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches'],d['config'].get('launch'),d['config'].get('loss_last'))"
timeout 300 python bench.py --config 6 --no-graph --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg6 eager',d['value'],d['ms_per_step'],d['config'].get('loss_last'))"
(timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -k "wider or attention" 2>&1 | grep -E "^E  |passed|failed" | cut -c1-300)
