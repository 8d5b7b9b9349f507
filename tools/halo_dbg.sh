mkdir -p gpurun_out
for shape in "32 256 256 64 128 1" "32 128 128 128 384 1" "32 256 256 64 192 1"; do
  for what in fwd; do python tools/gpu_conv_one.py $shape $what 12 2>&1 | tail -1; done
done > gpurun_out/epi_v2.log 2>&1
for shape in "32 256 256 64 64 3" "32 256 256 128 128 3" "32 128 128 128 128 3" "32 256 256 128 64 3"; do
  for what in fwd fwdres; do python tools/gpu_conv_one.py $shape $what 12 2>&1 | tail -1; done
done >> gpurun_out/epi_v2.log 2>&1
cat gpurun_out/epi_v2.log
