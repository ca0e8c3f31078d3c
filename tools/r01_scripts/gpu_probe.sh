#!/bin/bash
# Diagnostic call: per-shape time table, per-launch ncu metrics of every dense-conv launch of one step, and
# ncu --set full of the implicit-GEMM kernel on the standalone bench-shape cases.
tag=${1:-probe}
mkdir -p gpurun_out
timeout 300 python tools/shape_table.py --top 120 > gpurun_out/${tag}_shape_table.txt 2> gpurun_out/${tag}_shape_table.err
timeout 400 ncu --profile-from-start off -k regex:conv_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/${tag}_conv_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_conv.log 2>&1
gzip -f gpurun_out/${tag}_conv_launches.csv
for cs in 17 18 21 23; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -c 2 \
      -o gpurun_out/${tag}_case$cs -f tests/csrc/_bin/test_conv $cs > gpurun_out/${tag}_case$cs.log 2>&1
  ncu -i gpurun_out/${tag}_case$cs.ncu-rep --page raw --csv > gpurun_out/${tag}_case${cs}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_case$cs.ncu-rep --page details > gpurun_out/${tag}_case${cs}_details.txt 2>/dev/null
  ncu -i gpurun_out/${tag}_case$cs.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_case${cs}_source.csv.gz
  rm -f gpurun_out/${tag}_case$cs.ncu-rep
done
head -50 gpurun_out/${tag}_shape_table.txt
