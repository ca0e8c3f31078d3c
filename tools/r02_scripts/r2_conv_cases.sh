#!/bin/bash
# All standalone conv parity cases (tcgen05 kernels vs the CUDA-core direct convolution) + timings, then the
# conv3_kernel wait-cycle profile (test-only -DNPP_C3_PROF build under tests/csrc/_bin/prof).
tag=${1:-r2conv}
mkdir -p gpurun_out
bash tests/csrc/run_conv_cases.sh > gpurun_out/${tag}_conv_cases.log 2>&1
echo "run_conv_cases exit $?" >> gpurun_out/${tag}_conv_cases.log
grep -E "FAIL|exit code|run_conv_cases exit" gpurun_out/${tag}_conv_cases.log | head -20
grep -A1 "bench shape" gpurun_out/${tag}_conv_cases.log | grep -E "case|time"
out=gpurun_out/${tag}_conv3_prof.txt
: > $out
[ -f tests/csrc/_bin/prof/libnpp_b200.so ] || { echo "no profile build (tools/build_conv_prof.sh)"; exit 0; }
export LD_LIBRARY_PATH=$PWD/tests/csrc/_bin/prof:$LD_LIBRARY_PATH
for cs in 21 32 17 23 18 20; do
  timeout 90 tests/csrc/_bin/test_conv $cs 2>&1 | grep -v "PASS" >> $out
done
cat $out
