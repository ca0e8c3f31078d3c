#!/bin/bash
# Parity tests with the newest kernel selections; on failure bisect with the switches, bench with the first passing set.
tag=${1:-c6}
mkdir -p gpurun_out
run_tests() {  # $1 = label, rest = env assignments
  label=$1; shift
  ( time env "$@" timeout 600 python -m pytest tests -m gpu -q --maxfail 5 ) > gpurun_out/${tag}_pytest_$label.log 2>&1
  rc=$?
  echo "pytest exit $rc" >> gpurun_out/${tag}_pytest_$label.log
  echo "== $label: rc=$rc"; tail -8 gpurun_out/${tag}_pytest_$label.log | cut -c1-220
  return $rc
}
ENVS="X=1"
if ! run_tests default X=1; then
  ENVS="NPP_NODE_CAT_GRADS=0 NPP_DW_PIPE=0 NPP_NODE_STRIPED=0"
  for sw in NPP_NODE_CAT_GRADS NPP_DW_PIPE NPP_NODE_STRIPED; do
    if run_tests no_$sw $sw=0; then ENVS="$sw=0"; break; fi
  done
fi
echo "bench env: $ENVS"
env $ENVS timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $? ($ENVS)" >> gpurun_out/${tag}_bench.err
env $ENVS timeout 300 python tools/shape_table.py --top 150 > gpurun_out/${tag}_shape_table.txt 2> gpurun_out/${tag}_shape_table.err
env $ENVS timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt 2>&1
gzip -f gpurun_out/${tag}_launches.csv
cut -c1-700 gpurun_out/${tag}_bench.json
