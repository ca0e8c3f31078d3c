#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench line, ncu launch list of one eager step, ncu --set full of the
# dominant kernels.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh v5'
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( time timeout 480 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?" >> gpurun_out/${tag}_bench.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1
gzip -f gpurun_out/${tag}_launches.csv
for spec in "conv_gemm2_kernel:40" "conv3_kernel:20" "conv_wgrad3_kernel:10"; do
  name=${spec%%:*}; skip=${spec##*:}
  safe=$(echo $name | tr -c 'a-zA-Z0-9_' '_')
  timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$name" -s $skip -c 2 \
      -o gpurun_out/${tag}_full_$safe -f python tools/profile_step.py > gpurun_out/${tag}_ncu_full_$safe.log 2>&1
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page raw --csv > gpurun_out/${tag}_full_${safe}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page details > gpurun_out/${tag}_full_${safe}_details.txt 2>/dev/null
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_full_${safe}_source.csv.gz
  rm -f gpurun_out/${tag}_full_$safe.ncu-rep
done
tail -3 gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_bench.json | cut -c1-600
