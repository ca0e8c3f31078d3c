#!/bin/bash
# A/B timing of the dense-conv kernels on the bench shapes (standalone binary, one process per case):
# legacy kernels vs the current defaults.  Output: gpurun_out/<tag>_conv_ab.txt
tag=${1:-ab}
mkdir -p gpurun_out
out=gpurun_out/${tag}_conv_ab.txt
: > $out
for cs in 17 18 19 20 21 22 23 24 25 32 33 34; do
  echo "### legacy" >> $out
  NPP_CONV_EPI8=0 NPP_CONV3=0 timeout 90 tests/csrc/_bin/test_conv $cs 2>&1 | grep -v "PASS" >> $out
  echo "### default" >> $out
  timeout 90 tests/csrc/_bin/test_conv $cs 2>&1 | grep -v "PASS" >> $out
done
cat $out
