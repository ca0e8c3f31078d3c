#!/bin/bash
tag=${1:-c5}
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
  for sw in NPP_NODE_STRIPED NPP_NODE_FUSED_FINALIZE NPP_STEM_IM2COL; do
    if run_tests no_$sw $sw=0; then ENVS="$sw=0"; break; fi
  done
fi
echo "bench env: $ENVS"
env $ENVS timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $? ($ENVS)" >> gpurun_out/${tag}_bench.err
env $ENVS timeout 300 python tools/shape_table.py --top 150 > gpurun_out/${tag}_shape_table.txt 2> gpurun_out/${tag}_shape_table.err
cut -c1-700 gpurun_out/${tag}_bench.json
