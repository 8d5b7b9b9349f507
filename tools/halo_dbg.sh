mkdir -p gpurun_out
for shape in "32 256 256 64 64 3" "32 256 256 128 128 3" "32 128 128 128 128 3" "32 256 256 128 64 3" "32 64 64 256 256 3" "32 32 32 512 512 3" "32 128 128 256 256 3" "32 256 256 64 128 3"; do
  python tools/gpu_conv_one.py $shape fwd 10 2>&1 | tail -1
done > gpurun_out/halo_v5.log 2>&1
cat gpurun_out/halo_v5.log
