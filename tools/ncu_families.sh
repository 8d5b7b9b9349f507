# Per-kernel ncu evidence of one config-2 training step: one `--set full` capture per kernel family,
# details page exported as CSV (small), the .ncu-rep kept only for the dominant families (64 MiB merge limit).
#   bash tools/ncu_families.sh          (on the GPU box; writes gpurun_out/r02_ncu_*)
mkdir -p gpurun_out
cap() {  # name regex skip keep_rep
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c 1 \
    -f -o gpurun_out/r02_ncu_$1 python tools/gpu_step_once.py > gpurun_out/ncu_$1.log 2>&1
  if [ -f gpurun_out/r02_ncu_$1.ncu-rep ]; then
    ncu -i gpurun_out/r02_ncu_$1.ncu-rep --page details --csv > gpurun_out/r02_ncu_$1.details.csv 2>/dev/null
    [ "$4" = "keep" ] || rm -f gpurun_out/r02_ncu_$1.ncu-rep
    echo "$1 ok"
  else
    echo "$1 FAILED"; tail -3 gpurun_out/ncu_$1.log
  fi
  rm -f gpurun_out/ncu_$1.log
}
cap conv_pair256      'conv_halo2_kernel<.int.256'  4 keep
cap conv_pair128      'conv_halo2_kernel<.int.128'  2 keep
cap wgrad_pair        'wgrad_halo2_kernel'          4 keep
cap conv_halo64       'conv_halo_kernel<.int.64'    4 keep
cap wgrad_halo64      'wgrad_halo_kernel'           2 drop
cap conv_1x1          'conv_fwd_kernel<.int.256'    2 drop
cap wgrad_generic     'conv_wgrad_kernel'           2 drop
cap gn_apply          'gn_apply_kernel'             4 drop
cap gn_bwd_apply      'gn_bwd_apply_kernel'         4 drop
cap gn_bwd_sums       'gn_bwd_sums_kernel'          4 drop
cap resample          'resample2x_kernel'           2 drop
cap attn_fwd          'attn_fwd_kernel'             0 drop
cap attn_bwd_dkv      'attn_bwd_dkv_kernel'         0 drop
cap adamw_ema         'adamw_ema_kernel'            0 drop
cap palette_loss      'palette_loss_fwd_kernel'     0 drop
cap noise_pack        'noise_pack_kernel'           0 drop
ls -la gpurun_out/r02_ncu_* | awk '{print $5, $9}'
