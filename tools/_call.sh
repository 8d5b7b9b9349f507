mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_widen_cut.py tests/test_gpu_cond.py tests/test_gpu_prep.py -m gpu -q --no-header 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-250 | head -20)
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; tail -c 1500 gpurun_out/r02_bench_v3.json
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg3_v2.json 2> gpurun_out/r02_bench_cfg3_v2.err; tail -c 1200 gpurun_out/r02_bench_cfg3_v2.json; tail -3 gpurun_out/r02_bench_cfg3_v2.err
