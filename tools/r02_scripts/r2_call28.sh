#!/bin/bash
# Round-2 GPU call 28 (1 GPU): final tree — full GPU suite, then the three default bench lines (with baselines).
tag=r2c28
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -8 | cut -c1-300
for wl in search infer512 train; do
  timeout 600 python bench.py --workload $wl > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err
  echo "bench $wl exit $?"; grep '^{' gpurun_out/${tag}_bench_$wl.json | cut -c1-260
done
