#!/bin/bash
# Round-2 GPU call 21 (1 GPU): depthwise kernels with 8 / 4 / 2 channel-vector lanes per pixel: op tests, search tests,
# search + train bench.
tag=r2c21
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_search.py tests/test_gpu_search_step.py tests/test_gpu_network.py -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -8 | cut -c1-300
run () {  # name workload args...
  name=$1; wl=$2; shift; shift
  timeout 600 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-gpu-reference "$@" > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "bench $name exit $?: $(python -c "import json;d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{\"')][0];r=d.get('roofline') or {};print(d['ms_per_step'], d['value'], d['e2e']['value'], 'roofline', r.get('frac'), (r.get('wgrad') or {}).get('frac'), (r.get('hbm_class') or {}).get('frac')); c=r.get('classes') or {}; print({k:v for k,v in c.items() if 'dw' in k})" 2>/dev/null)"
  tail -2 gpurun_out/${tag}_bench_$name.err | cut -c1-300
}
run search search
run train train --no-kernel-table
