#!/bin/bash
# Round-2 GPU call 20 (1 GPU): extended conv parity modes (ring / one producer / multi-tile) + compute-sanitizer
# memcheck of the restructured conv kernels on small cases.
tag=r2c20
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest_conv.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_conv.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest_conv.log | tail -8 | cut -c1-300
: > gpurun_out/${tag}_memcheck.txt
for cs in 2 4 12 27 30; do
  for env in "" "NPP_CONV3_MIN_TILES=1 NPP_CONV3_PAD_PCT=400" "NPP_CONV3_MIN_TILES=1 NPP_CONV3_PAD_PCT=400 NPP_CONV3_WRES=0"; do
    echo "### case $cs env: $env" >> gpurun_out/${tag}_memcheck.txt
    env $env timeout 300 compute-sanitizer --tool memcheck --print-limit 5 tests/csrc/_bin/test_conv $cs 2>&1 | grep -E "ERROR SUMMARY|Invalid|out of bounds|case .* (OK|FAILED)|Error" | head -8 >> gpurun_out/${tag}_memcheck.txt
  done
done
cat gpurun_out/${tag}_memcheck.txt | cut -c1-200
