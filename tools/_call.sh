mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header 2>&1 | tail -3) > gpurun_out/r02_final_multi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_final_bench_2gpu.json 2> gpurun_out/c52.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference_2gpu.json 2>> gpurun_out/c52.err
cat gpurun_out/r02_final_multi.txt; for f in bench_2gpu bench_reference_2gpu; do python -c "
import json;d=json.loads(open('gpurun_out/r02_final_$f.json').read().strip().splitlines()[-1]);print('$f',d.get('n_gpus'),round(d['value'],2),d.get('ms_per_step'),d.get('cpu_baseline',{}).get('cores'))"; done; tail -2 gpurun_out/c52.err | cut -c1-200
