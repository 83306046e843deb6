#!/bin/bash
# ncu --set full captures of the top kernels inside one cfg2 bench step (final build) + the fused VQ kernel in isolation.
# Run on a B200 box from the repo root: bash tools/r02_capture.sh.  Only the CSV exports (raw metrics page, details page) come
# back in gpurun_out/ (the reports themselves are 10-20 MB each; gpurun_out/ is limited to 64 MiB); the two small VQ reports
# are kept for the source page.
set -u
mkdir -p gpurun_out /tmp/rep
NCU="ncu --set full --clock-control none --import-source on -f"
export_rep() { # name
  ncu -i "/tmp/rep/$1.ncu-rep" --page raw --csv > "gpurun_out/r02_full_$1_raw.csv" 2>/dev/null
  ncu -i "/tmp/rep/$1.ncu-rep" --page details --csv > "gpurun_out/r02_full_$1_details.csv" 2>/dev/null
}
cap() { # name regex count [config]
  timeout 400 $NCU --profile-from-start off -k "regex:$2" -c "$3" -o "/tmp/rep/$1" python bench.py --ncu-step --config "${4:-cfg2}" --warmup 3 > "gpurun_out/r02_full_$1.log" 2>&1
  echo "$1 rc=$?"
  export_rep "$1"
}
cap halo_t 'conv_fwd_tc_halo_t_kernel' 5
cap halo2 'conv_fwd_tc_halo2_kernel' 3
cap wgrad_halo 'conv_wgrad_tc_halo_kernel' 3
cap gn 'gn_' 12
for shape in "16384 1024" "8192 8192"; do
  set -- $shape
  timeout 200 $NCU -k regex:vq_fused_kernel -s 3 -c 1 -o /tmp/rep/vq_fused_k$2 python tools/vq_profile.py $1 $2 normal > gpurun_out/r02_full_vq_k$2.log 2>&1
  export_rep vq_fused_k$2
  ncu -i /tmp/rep/vq_fused_k$2.ncu-rep --page source --csv > gpurun_out/r02_full_vq_fused_k$2_source.csv 2>/dev/null
done
du -sh gpurun_out
