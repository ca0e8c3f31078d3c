#!/bin/bash
# conv3_kernel wait-cycle profile (test-only -DNPP_C3_PROF build of the library under tests/csrc/_bin/prof).
tag=${1:-r2prof}
mkdir -p gpurun_out
out=gpurun_out/${tag}_conv3_prof.txt
: > $out
[ -f tests/csrc/_bin/prof/libnpp_b200.so ] || { echo "no profile build (tools/build_conv_prof.sh)"; exit 0; }
export LD_LIBRARY_PATH=$PWD/tests/csrc/_bin/prof:$LD_LIBRARY_PATH
for cs in 21 32 17 23 24 25; do
  timeout 90 tests/csrc/_bin/test_conv $cs 2>&1 | grep -v "PASS" >> $out
done
for t in 0 1 2; do
  echo "### case 21 NPP_CONV3_TILE=$t" >> $out
  NPP_CONV3_TILE=$t timeout 90 tests/csrc/_bin/test_conv 21 2>&1 | grep -v "PASS" >> $out
done
echo "### case 21 NPP_CONV3_2PROD=0" >> $out
NPP_CONV3_2PROD=0 timeout 90 tests/csrc/_bin/test_conv 21 2>&1 | grep -v "PASS" >> $out
echo "### case 21 NPP_CONV3=0 (generic 128-pixel kernel)" >> $out
NPP_CONV3=0 timeout 90 tests/csrc/_bin/test_conv 21 2>&1 | grep -v "PASS" >> $out
cat $out
