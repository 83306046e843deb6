set -u
mkdir -p gpurun_out /tmp/rep
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
for cfg in cfg2 cfg4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${cfg}_final2.csv python bench.py --ncu-step --config $cfg --warmup 3 > gpurun_out/launches_${cfg}_final2.log 2>&1
  echo "launch list $cfg rc=$?"
  python tools/summarize_launches.py gpurun_out/launches_${cfg}_final2.csv > gpurun_out/launches_${cfg}_final2.txt 2>&1
  head -12 gpurun_out/launches_${cfg}_final2.txt
done
NCU="ncu --set full --clock-control none --import-source on -f"
for shape in "16384 1024" "8192 8192"; do
  set -- $shape
  timeout 200 $NCU -k regex:vq_fused_kernel -s 3 -c 1 -o /tmp/rep/vq_fused_k$2 python tools/vq_profile.py $1 $2 normal > gpurun_out/r02_full_vq_k$2.log 2>&1
  ncu -i /tmp/rep/vq_fused_k$2.ncu-rep --page raw --csv > gpurun_out/r02_full2_vq_fused_k$2_raw.csv 2>/dev/null
done
python tools/ncu_raw_summary.py gpurun_out/r02_full2_vq_fused_k1024_raw.csv gpurun_out/r02_full2_vq_fused_k8192_raw.csv 2>&1 | tail -12
du -sh gpurun_out
