mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dq_tc -s 1 -c 1 -f -o gpurun_out/r02_ncu_attn_bwd_dq_tc python tools/gpu_attn_bwd_one.py 32 1024 16 32 0 2 > gpurun_out/c30.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv_tc -s 1 -c 1 -f -o gpurun_out/r02_ncu_attn_bwd_dkv_tc python tools/gpu_attn_bwd_one.py 32 1024 16 32 0 2 >> gpurun_out/c30.log 2>&1
ls -la gpurun_out/r02_ncu_attn_bwd_*.ncu-rep
