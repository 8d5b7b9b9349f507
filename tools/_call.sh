mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_production_shapes.py -m gpu -q --no-header -x 2>&1 | tail -3)
(timeout 300 python -m pytest tests/test_gpu_palette.py tests/test_gpu_cond.py -m gpu -q --no-header 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-250 | head)
for s in "32 256 256 64 64 3" "32 256 256 128 128 3" "32 128 128 128 128 3" "32 128 128 256 256 3" "32 64 64 256 256 3" "32 32 32 512 512 3"; do for m in fwd fwdstats fwdres; do timeout 120 python tools/gpu_conv_one.py $s $m 10 2>&1 | tail -1; done; done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_v4.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['conv'])"
