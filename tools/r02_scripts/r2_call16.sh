#!/bin/bash
# Round-2 GPU call 16 (1 GPU): dual-issuer wgrad kernels + fp32x2 statistics: conv cases, GPU tests, train bench.
tag=r2c16
mkdir -p gpurun_out
bash tools/r02_scripts/r2_conv_cases.sh ${tag} > gpurun_out/${tag}_conv_cases_stdout.txt 2>&1
grep -E "FAIL|exit code|run_conv_cases exit" gpurun_out/${tag}_conv_cases_stdout.txt | head
grep -E "epiprof" gpurun_out/${tag}_conv_cases_stdout.txt | cut -c1-330
grep -E "^\[case|time" gpurun_out/${tag}_conv_cases.log | grep -A1 "bench shape" | grep time | cut -c1-200
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -12 | cut -c1-300
run () {  # name workload args...
  name=$1; wl=$2; shift; shift
  timeout 600 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-gpu-reference "$@" > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "bench $name exit $?: $(python -c "import json;d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{')][0];r=d.get('roofline') or {};print(d['ms_per_step'], d['value'], d['e2e']['value'], 'roofline', r.get('frac'), (r.get('wgrad') or {}).get('frac'), (r.get('hbm_class') or {}).get('frac'))" 2>/dev/null)"
  tail -2 gpurun_out/${tag}_bench_$name.err | cut -c1-300
}
run train train
