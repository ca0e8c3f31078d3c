#!/bin/bash
# Round-2 GPU call 19 (1 GPU): evidence of the final build — smoke, ncu launch list + per-launch DRAM traffic of the
# conv kernels of one eager step, ncu --set full of the two most frequent conv kernels, the three bench lines with
# baselines.
tag=r2c19
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/${tag}_smoke.log | cut -c1-200
timeout 400 ncu --profile-from-start off -k regex:"conv_|conv3" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/${tag}_conv_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_conv.log 2>&1
echo "ncu conv exit $?"
python tools/ncu_traffic.py gpurun_out/${tag}_conv_launches.csv gpurun_out/${tag}_roofline_traffic.json "round-2 final conv kernels (converged-warp issue, two MMA issuers, resident weights)" | cut -c1-300
gzip -f gpurun_out/${tag}_conv_launches.csv
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1
head -12 gpurun_out/${tag}_launches_summary.txt | cut -c1-160
gzip -f gpurun_out/${tag}_launches.csv
for k in 'conv3_kernel<\(int\)32>' 'conv_gemm2_kernel<\(int\)128>' 'conv3_kernel<\(int\)128>'; do
  n=$(echo $k | tr -cd 'a-z0-9_')
  timeout 300 ncu --profile-from-start off --kernel-name-base demangled -k regex:"$k" -c 2 --set full --clock-control none --import-source on \
      -o gpurun_out/${tag}_full_$n -f python tools/profile_step.py > gpurun_out/${tag}_ncu_full_$n.log 2>&1
  echo "ncu full $n exit $?"
  ncu -i gpurun_out/${tag}_full_$n.ncu-rep --page details > gpurun_out/${tag}_ncu_full_${n}_details.txt 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null | cut -c1-120
cp gpurun_out/${tag}_roofline_traffic.json profiles/roofline_traffic.json   # bench.py reads it for roofline.traffic
for wl in train search infer512; do
  timeout 900 python bench.py --workload $wl > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err
  echo "bench $wl exit $?"; grep '^{' gpurun_out/${tag}_bench_$wl.json | cut -c1-400
done
