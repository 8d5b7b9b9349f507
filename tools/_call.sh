mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -x 2>&1 | tail -15) > gpurun_out/c23_multi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_v2.json 2> gpurun_out/r02_bench_2gpu_v2.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu_samebox.json 2>/dev/null
cat gpurun_out/c23_multi.txt; for f in gpurun_out/r02_bench_2gpu_v2.json gpurun_out/r02_bench_1gpu_samebox.json; do python -c "
import json,sys;d=json.loads(open('$f').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('allreduce'),d['clocks'])"; done; tail -3 gpurun_out/r02_bench_2gpu_v2.err
