#!/bin/bash
tag=${1:-c4}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?" >> gpurun_out/${tag}_bench.err
for spec in "dw_tile_kernel:20" "dw_tile_wgrad:10" "bilinear_bwd:4" "ce_bwd_kernel:0"; do
  name=${spec%%:*}; skip=${spec##*:}
  safe=$(echo $name | tr -c 'a-zA-Z0-9_' '_')
  timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$name" -s $skip -c 2 \
      -o gpurun_out/${tag}_full_$safe -f python tools/profile_step.py > gpurun_out/${tag}_ncu_full_$safe.log 2>&1
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page raw --csv > gpurun_out/${tag}_full_${safe}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_full_$safe.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_full_${safe}_source.csv.gz
  rm -f gpurun_out/${tag}_full_$safe.ncu-rep
done
cut -c1-700 gpurun_out/${tag}_bench.json
