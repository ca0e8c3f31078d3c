#!/bin/bash
tag=${1:-c2}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_conv.py -q --maxfail 8 ) > gpurun_out/${tag}_pytest_conv.log 2>&1
rc=$?
echo "pytest conv exit $rc" >> gpurun_out/${tag}_pytest_conv.log
tail -40 gpurun_out/${tag}_pytest_conv.log
bash tools/gpu_conv_ab.sh $tag > /dev/null 2>&1
grep "###\|case\|time" gpurun_out/${tag}_conv_ab.txt | cut -c1-250
if [ $rc -eq 0 ]; then
  bash tools/gpu_round.sh $tag
fi
