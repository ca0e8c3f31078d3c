#!/bin/bash
# Round-2 GPU call 27 (1 GPU): batched-load first pass of the separable bilinear backward: op / network / golden tests, train + infer A/B.
tag=r2c27
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_network.py tests/test_gpu_golden.py tests/test_gpu_eval_step.py -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -8 | cut -c1-300
run () {  # name workload env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 400 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "bench $name exit $?: $(python -c "import json;d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{\"')][0];print(d['ms_per_step'], d['value'], d['e2e']['value'])" 2>/dev/null)"
  tail -2 gpurun_out/${tag}_bench_$name.err | cut -c1-300
}
run train_bwdb1 train NPP_BILINEAR_BWD_BATCHED=1
run train_bwdb0 train NPP_BILINEAR_BWD_BATCHED=0
