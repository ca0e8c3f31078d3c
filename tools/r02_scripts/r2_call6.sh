#!/bin/bash
# Round-2 GPU call 6 (1 GPU): whole GPU suite on the build with per-consumer gradient handles + gather-form CE backward,
# train bench, A/B against the previous behaviour (NPP_NODE_CAT_GRADS=0 / NPP_CE_BWD_ATOMIC=1), launch list.
tag=r2c6
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s --maxfail 15 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -25 | cut -c1-300
timeout 600 python bench.py --no-gpu-reference > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err
echo "bench train exit $?"; grep '^{' gpurun_out/${tag}_bench_train.json | cut -c1-300
for cfg in "NPP_CE_BWD_ATOMIC=1" "NPP_NODE_CAT_GRADS=0"; do
  safe=$(echo "$cfg" | tr -c 'A-Za-z0-9_=' '_')
  env $cfg timeout 300 python bench.py --steps 8 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_ab_$safe.json 2> gpurun_out/${tag}_ab_$safe.err
  ms=$(python -c "import json,sys; print([json.loads(l) for l in open('gpurun_out/${tag}_ab_$safe.json') if l.startswith('{')][0]['ms_per_step'])" 2>/dev/null || echo fail)
  echo "A/B $cfg -> $ms ms/step" | tee -a gpurun_out/${tag}_ab.txt
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1
gzip -f gpurun_out/${tag}_launches.csv
head -30 gpurun_out/${tag}_launches_summary.txt | cut -c1-160
