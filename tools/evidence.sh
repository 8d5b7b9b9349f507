# Round evidence: tests, one-step ncu launch list with DRAM traffic, ncu --set full of the dominant kernels, benches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ev_pytest.log 2>&1; tail -2 gpurun_out/ev_pytest.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/gpu_step_once.py > gpurun_out/ev_step.log 2>&1
python tools/ncu_traffic.py gpurun_out/step_launches.csv gpurun_out/conv_traffic.json > gpurun_out/step_kernels.txt 2>&1; head -30 gpurun_out/step_kernels.txt; tail -1 gpurun_out/step_kernels.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -o gpurun_out/r01_ncu_conv_halo_128x128_256_final python tools/gpu_conv_one.py 32 256 256 128 128 3 fwd 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 3 -c 1 -o gpurun_out/r01_ncu_conv_halo_64x64_256_final python tools/gpu_conv_one.py 32 256 256 64 64 3 fwd 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_halo -s 4 -c 1 -o gpurun_out/r01_ncu_wgrad_halo_128x128_256_final python tools/gpu_conv_one.py 32 256 256 128 128 3 wgrad 2 > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_final.log 2>&1; tail -1 gpurun_out/bench_reference_final.log
