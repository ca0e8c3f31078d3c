#!/bin/bash
# Round-2 GPU call 24 (1 GPU): final validation of the tree: full GPU suite, smoke, default train bench line.
tag=r2c24
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -8 | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/${tag}_smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err
echo "bench exit $?"; grep '^{' gpurun_out/${tag}_bench_train.json | cut -c1-300
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2c24_bench_train.json') if l.startswith('{"')][0]
r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], 'roofline', r['frac'], 'wgrad', r['wgrad']['frac'], 'hbm', r['hbm_class']['frac'], 'gpu_ref', d['gpu_reference'].get('speedup_vs_best'))
for k,v in r['classes'].items(): print(' ', k, v)
PY
