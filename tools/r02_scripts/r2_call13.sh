#!/bin/bash
# Round-2 GPU call 13 (1 GPU): wgrad companion stream (train) and MixedOp worker streams (search): tests + A/B.
tag=r2c13
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -12 | cut -c1-300
run () {  # name workload env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 400 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "bench $name exit $?: $(python -c "import json;d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{')][0];print(d['ms_per_step'], d['value'], d['e2e']['value'])" 2>/dev/null)"
  tail -2 gpurun_out/${tag}_bench_$name.err | cut -c1-300
}
run train_wgrad1 train NPP_WGRAD_STREAM=1
run train_wgrad0 train NPP_WGRAD_STREAM=0
run search_workers3 search NPP_BRANCH_STREAMS=3
run search_workers0 search NPP_BRANCH_STREAMS=0
run search_workers6 search NPP_BRANCH_STREAMS=6
run infer512 infer512 NPP_WGRAD_STREAM=1
