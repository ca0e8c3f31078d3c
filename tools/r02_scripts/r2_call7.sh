#!/bin/bash
# Round-2 GPU call 7 (1 GPU): programmatic dependent launch on every kernel — whole GPU suite, bench with PDL on / off.
tag=r2c7
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --maxfail 5 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -12 | cut -c1-300
timeout 400 python bench.py --no-gpu-reference --no-cpu-baseline > gpurun_out/${tag}_bench_pdl1.json 2> gpurun_out/${tag}_bench_pdl1.err
echo "bench pdl=1 exit $?"; grep '^{' gpurun_out/${tag}_bench_pdl1.json | cut -c1-260; tail -2 gpurun_out/${tag}_bench_pdl1.err | cut -c1-200
NPP_PDL=0 timeout 400 python bench.py --no-gpu-reference --no-cpu-baseline --no-kernel-table --steps 10 > gpurun_out/${tag}_bench_pdl0.json 2> gpurun_out/${tag}_bench_pdl0.err
echo "bench pdl=0 exit $?"; grep '^{' gpurun_out/${tag}_bench_pdl0.json | cut -c1-260
timeout 300 python bench.py --workload search --no-gpu-reference --no-cpu-baseline --no-kernel-table > gpurun_out/${tag}_bench_search_pdl1.json 2> gpurun_out/${tag}_bench_search_pdl1.err
echo "bench search pdl=1 exit $?"; grep '^{' gpurun_out/${tag}_bench_search_pdl1.json | cut -c1-260
timeout 300 python bench.py --workload infer512 --no-gpu-reference --no-cpu-baseline --no-kernel-table > gpurun_out/${tag}_bench_infer_pdl1.json 2> gpurun_out/${tag}_bench_infer_pdl1.err
echo "bench infer512 pdl=1 exit $?"; grep '^{' gpurun_out/${tag}_bench_infer_pdl1.json | cut -c1-260
