set -x
for spec in "node_bwd_reduce:200" "node_fwd_kernel:200" "dw_tile_kernel:30" "dw_tile_wgrad:20" "maxpool_bwd:20" "bilinear_bwd:10" "conv_gemm_kernel<32>:60"; do
  name=${spec%%:*}; skip=${spec##*:}
  safe=$(echo $name | tr -c 'a-zA-Z0-9_' '_')
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$name" -s $skip -c 2 -o gpurun_out/bw_$safe -f python tools/profile_step.py > gpurun_out/ncu_bw_$safe.log 2>&1
  ncu -i gpurun_out/bw_$safe.ncu-rep --page raw --csv > gpurun_out/bw_${safe}_raw.csv 2>/dev/null
  ncu -i gpurun_out/bw_$safe.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/bw_${safe}_source.csv.gz
  rm -f gpurun_out/bw_$safe.ncu-rep
done
ls -la gpurun_out | head -40
