mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_final_bench_8gpu.json 2> gpurun_out/c53.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_final_bench_8gpu.json').read().strip().splitlines()[-1]);print('8gpu',d.get('n_gpus'),round(d['value'],1),d.get('ms_per_step'),d.get('allreduce',{}).get('buckets'),d.get('clocks'))"; tail -3 gpurun_out/c53.err | cut -c1-200
