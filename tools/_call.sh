mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_jit.py -m gpu -q --no-header 2>&1 | grep -E "^E  |passed|failed" | cut -c1-200 | head)
JG_BREAKDOWN=gpurun_out/r02_breakdown_cfg6.json timeout 300 python bench.py --config 6 --steps 10 --warmup 4 > gpurun_out/r02_bench_cfg6_v3.json 2>/dev/null
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_cfg6_v3.json').read().strip().splitlines()[-1]);print('cfg6',d['value'],d['unit'],d['ms_per_step'],d['e2e']['value'])
b=json.load(open('gpurun_out/r02_breakdown_cfg6.json'));print(b['instrumented_step_ms'],b['sum_call_ms'])
tot={}
for r in b['rows']:
    k=r['key'].split(' ')[0]; tot.setdefault(k,[0,0]); tot[k][0]+=r['ms']; tot[k][1]+=r['calls']
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][0])[:14]: print('%8.3f ms %4d  %s'%(v[0],v[1],k))"
