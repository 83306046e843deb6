#!/bin/bash
# ncu --set full captures of the top kernels inside one cfg2 bench step (final build) + the fused VQ kernel in isolation.
# Run on a B200 box from the repo root: bash tools/r02_capture.sh ; reports land in gpurun_out/.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off -f"
cap() { # name regex count [config]
  timeout 400 $NCU -k "regex:$2" -c "$3" -o "gpurun_out/r02_full_$1" python bench.py --ncu-step --config "${4:-cfg2}" --warmup 3 > "gpurun_out/r02_full_$1.log" 2>&1
  echo "$1 rc=$?"
}
cap halo_t 'conv_fwd_tc_halo_t_kernel' 6
cap halo2 'conv_fwd_tc_halo2_kernel' 4
cap wgrad_halo 'conv_wgrad_tc_halo_kernel' 4
cap gn 'gn_' 12
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:vq_fused_kernel -s 3 -c 1 -o gpurun_out/r02_full_vq_fused_k1024 python tools/vq_profile.py 16384 1024 normal > gpurun_out/r02_full_vq1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -k regex:vq_fused_kernel -s 3 -c 1 -o gpurun_out/r02_full_vq_fused_k8192 python tools/vq_profile.py 8192 8192 normal > gpurun_out/r02_full_vq2.log 2>&1
ls -la gpurun_out/*.ncu-rep
