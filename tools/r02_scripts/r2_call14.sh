#!/bin/bash
# Round-2 GPU call 14 (1 GPU): wave-scheduled cell primitives on worker streams (NPP_CELL_BRANCHES) in the derived net.
tag=r2c14
mkdir -p gpurun_out
( time NPP_CELL_BRANCHES=1 timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_golden.py tests/test_gpu_engine.py tests/test_gpu_two_streams.py tests/test_gpu_baseline_config.py tests/test_gpu_checkpoint.py -m gpu -q --maxfail 10 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -12 | cut -c1-300
run () {  # name workload env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 400 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "bench $name exit $?: $(python -c "import json;d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{')][0];print(d['ms_per_step'], d['value'], d['e2e']['value'])" 2>/dev/null)"
  tail -2 gpurun_out/${tag}_bench_$name.err | cut -c1-300
}
run train_cb0 train NPP_CELL_BRANCHES=0
run train_cb1_w2 train NPP_CELL_BRANCHES=1 NPP_BRANCH_STREAMS=2
run train_cb1_w4 train NPP_CELL_BRANCHES=1 NPP_BRANCH_STREAMS=4
run train_cb1_w6 train NPP_CELL_BRANCHES=1 NPP_BRANCH_STREAMS=6
run infer_cb1_w4 infer512 NPP_CELL_BRANCHES=1 NPP_BRANCH_STREAMS=4
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
