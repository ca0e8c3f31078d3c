#!/bin/bash
# Two-group conv epilogue (NPP_CONV_EPI2): all conv parity cases, then A/B timing of the bench shapes.
tag=${1:-r2epi2}
mkdir -p gpurun_out
CONV_CASE_TIMEOUT=40 bash tests/csrc/run_conv_cases.sh > gpurun_out/${tag}_conv_cases.log 2>&1
echo "run_conv_cases exit $?" >> gpurun_out/${tag}_conv_cases.log
grep -E "FAIL|exit code|run_conv_cases exit|timed out" gpurun_out/${tag}_conv_cases.log | head -20
out=gpurun_out/${tag}_ab.txt
: > $out
for cs in 17 18 19 20 21 22 23 24 25 32 33 34; do
  for v in 1 0; do
    echo "### case $cs NPP_CONV_EPI2=$v" >> $out
    NPP_CONV_EPI2=$v timeout 60 tests/csrc/_bin/test_conv $cs 2>&1 | grep -E "time  :|FAIL" | cut -c1-150 >> $out
  done
done
cat $out
