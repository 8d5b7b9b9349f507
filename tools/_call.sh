mkdir -p gpurun_out
for c in 3 4 5; do JG_BREAKDOWN=gpurun_out/r02_breakdown_cfg$c.json timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/c50_$c.json 2>/dev/null; python -c "
import json
b=json.load(open('gpurun_out/r02_breakdown_cfg$c.json'));print('cfg$c', b['instrumented_step_ms'],b['sum_call_ms'])
tot={}
for r in b['rows']:
    k=r['key'].split(' ')[0]; tot.setdefault(k,[0,0]); tot[k][0]+=r['ms']; tot[k][1]+=r['calls']
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][0])[:9]: print('   %8.3f ms %4d  %s'%(v[0],v[1],k))"; done
