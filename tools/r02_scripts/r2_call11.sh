#!/bin/bash
# Round-2 GPU call 11 (1 GPU): two-stream schedule extended to the refinement stage / heads and to the supernet:
# whole GPU suite with NPP_TWO_STREAMS=1, the three workloads with the switch on and off.
tag=r2c11
mkdir -p gpurun_out
( time NPP_TWO_STREAMS=1 timeout 1500 python -m pytest tests -m gpu -q -s --maxfail 10 ) > gpurun_out/${tag}_pytest_two.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_two.log
grep -E "passed|failed|FAILED|ERROR|two streams vs one" gpurun_out/${tag}_pytest_two.log | tail -14 | cut -c1-330
for wl in train search infer512; do
  for two in 1 0; do
    NPP_TWO_STREAMS=$two timeout 400 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_${wl}_two$two.json 2> gpurun_out/${tag}_bench_${wl}_two$two.err
    echo "bench $wl two=$two exit $?: $(grep '^{' gpurun_out/${tag}_bench_${wl}_two$two.json | cut -c1-20) $(python -c "import json;d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_${wl}_two$two.json') if l.startswith('{')][0];print(d['ms_per_step'], d['value'], d['e2e']['value'])" 2>/dev/null)"
    tail -2 gpurun_out/${tag}_bench_${wl}_two$two.err | cut -c1-300
  done
done
