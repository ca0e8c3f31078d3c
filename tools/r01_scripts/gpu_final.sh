#!/bin/bash
# A/B of the node-backward variants (short benches), parity tests of the chosen one, then the evidence set:
# smoke, full bench line, per-launch ncu metrics of the dense-conv kernels (roofline.traffic), launch list.
tag=${1:-fin}
mkdir -p gpurun_out
best=""; best_ms=999999
for cfg in "NPP_NODE_STRIPED=0" "NPP_NODE_STRIPED=0 NPP_NODE_CAT_GRADS=0" "NPP_NODE_STRIPES=2"; do
  safe=$(echo "$cfg" | tr -c 'A-Za-z0-9_=' '_')
  env $cfg timeout 300 python bench.py --steps 6 --no-cpu-baseline > gpurun_out/${tag}_ab_$safe.json 2> gpurun_out/${tag}_ab_$safe.err
  ms=$(python -c "import json,sys; print(json.load(open('gpurun_out/${tag}_ab_$safe.json'))['ms_per_step'])" 2>/dev/null || echo 999999)
  echo "A/B $cfg -> $ms ms/step" | tee -a gpurun_out/${tag}_ab.txt
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$best_ms') else 1)"; then best="$cfg"; best_ms=$ms; fi
done
echo "best: $best ($best_ms ms)" | tee -a gpurun_out/${tag}_ab.txt
( time env $best timeout 600 python -m pytest tests -m gpu -q --maxfail 5 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $? ($best)" >> gpurun_out/${tag}_pytest.log
tail -6 gpurun_out/${tag}_pytest.log | cut -c1-200
env $best timeout 200 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -2 gpurun_out/${tag}_smoke.log
env $best timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $? ($best)" >> gpurun_out/${tag}_bench.err
env $best timeout 300 ncu --profile-from-start off -k regex:"conv_|conv3" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/${tag}_conv_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_conv.log 2>&1
gzip -f gpurun_out/${tag}_conv_launches.csv
env $best timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1
gzip -f gpurun_out/${tag}_launches.csv
cut -c1-600 gpurun_out/${tag}_bench.json
