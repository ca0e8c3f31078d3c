#!/bin/bash
# Round-2 GPU call 3 (1 GPU): whole GPU test suite (incl. stage-wise parity at the BASELINE config, EvalStep, SearchStep,
# label synthesis, optimizer resume), smoke, bench of the three workloads, N1 timing.
tag=r2c3
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s --maxfail 12 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -25 | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -3 gpurun_out/${tag}_smoke.log | cut -c1-900
timeout 600 python bench.py > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err
echo "bench train exit $?"; tail -c 1500 gpurun_out/${tag}_bench_train.json; tail -3 gpurun_out/${tag}_bench_train.err | cut -c1-300
timeout 600 python bench.py --workload search > gpurun_out/${tag}_bench_search.json 2> gpurun_out/${tag}_bench_search.err
echo "bench search exit $?"; cut -c1-400 gpurun_out/${tag}_bench_search.json; tail -3 gpurun_out/${tag}_bench_search.err | cut -c1-300
timeout 600 python bench.py --workload infer512 > gpurun_out/${tag}_bench_infer512.json 2> gpurun_out/${tag}_bench_infer512.err
echo "bench infer512 exit $?"; cut -c1-400 gpurun_out/${tag}_bench_infer512.json; tail -3 gpurun_out/${tag}_bench_infer512.err | cut -c1-300
timeout 300 python tools/bench_pose_post.py > gpurun_out/${tag}_pose_post.txt 2>&1
tail -5 gpurun_out/${tag}_pose_post.txt
