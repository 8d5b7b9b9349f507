# Round-end evidence on one B200: the whole GPU suite, the default bench (both arms), every other configuration.
#   bash tools/final_evidence.sh      (writes gpurun_out/r02_final_*)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header > gpurun_out/r02_final_gpu_suite.log 2>&1; tail -3 gpurun_out/r02_final_gpu_suite.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2>/dev/null
for c in 3 4 5 6; do timeout 300 python bench.py --config $c --steps 10 --warmup 4 > gpurun_out/r02_final_bench_cfg$c.json 2>/dev/null; done
timeout 300 python bench.py --config 4 --size 512x384 --steps 5 --warmup 3 > gpurun_out/r02_final_bench_cfg4_512.json 2>/dev/null
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for f in bench bench_reference bench_cfg3 bench_cfg4 bench_cfg5 bench_cfg6 bench_cfg4_512; do python -c "
import json;d=json.loads(open('gpurun_out/r02_final_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['value'],2),d['unit'],round(d.get('ms_per_step',0),2),round(d['e2e']['value'],2),d.get('roofline',{}).get('frac'),d.get('clocks',{}).get('sm_mhz'))"; done
