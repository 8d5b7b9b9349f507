# GroupNorm kernel timings (ncu launch list: per-kernel device time)
mkdir -p gpurun_out
: > gpurun_out/gn_sweep4.log
for shape in "32 256 64" "32 128 128" "32 64 256"; do
    echo "== shape $shape" >> gpurun_out/gn_sweep4.log
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gn_ -c 40 --csv --log-file /tmp/gnl.csv python tools/gpu_gn_one.py $shape 1 > /dev/null 2>&1
    python tools/ncu_launch_summary.py /tmp/gnl.csv | grep -v finalize | grep -v param_grad >> gpurun_out/gn_sweep4.log 2>&1
done
