#!/bin/bash
# Round-2 GPU call 1 (1 GPU): parity at the BASELINE config + per-stage trace, current bench, A/B timing of the
# round-1 candidates, ncu --set full of the node backward reduce kernels, launch list of the current build.
tag=r2c1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_baseline_config.py -q -s ) > gpurun_out/${tag}_pytest_baseline_config.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_baseline_config.log
tail -5 gpurun_out/${tag}_pytest_baseline_config.log | cut -c1-300
timeout 600 python tools/parity_trace.py --out gpurun_out/${tag}_parity_trace.txt > gpurun_out/${tag}_parity_trace.log 2>&1
echo "trace exit $?"; tail -3 gpurun_out/${tag}_parity_trace.txt
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; cut -c1-300 gpurun_out/${tag}_bench.json
for cfg in "NPP_CONV_PAIR=1" "NPP_SE_BWD2=1" "NPP_PACK_TILES=1" "NPP_CE_BWD_SEP=1" "NPP_BILINEAR_SEP=0"; do
  safe=$(echo "$cfg" | tr -c 'A-Za-z0-9_=' '_')
  env $cfg timeout 300 python bench.py --steps 8 --no-cpu-baseline > gpurun_out/${tag}_ab_$safe.json 2> gpurun_out/${tag}_ab_$safe.err
  ms=$(python -c "import json,sys; print(json.load(open('gpurun_out/${tag}_ab_$safe.json'))['ms_per_step'])" 2>/dev/null || echo fail)
  echo "A/B $cfg -> $ms ms/step" | tee -a gpurun_out/${tag}_ab.txt
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1
gzip -f gpurun_out/${tag}_launches.csv
for spec in "node_bwd_reduce:150" "node_bwd_apply:150" "node_fwd:150"; do
  name=${spec%%:*}; skip=${spec##*:}
  safe=$(echo $name | tr -c 'a-zA-Z0-9_' '_')
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$name" -s $skip -c 2 -o gpurun_out/${tag}_full_$safe -f python tools/profile_step.py > gpurun_out/${tag}_ncu_full_$safe.log 2>&1
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page raw --csv > gpurun_out/${tag}_full_${safe}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page details > gpurun_out/${tag}_full_${safe}_details.txt 2>/dev/null
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_full_${safe}_source.csv.gz
  rm -f gpurun_out/${tag}_full_$safe.ncu-rep
done
ls gpurun_out | grep ${tag} | head -50
