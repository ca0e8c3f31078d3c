#!/bin/bash
tag=r2c4
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_baseline_config.py tests/test_gpu_engine.py tests/test_gpu_labels.py tests/test_gpu_optim.py tests/test_gpu_search_step.py tests/test_gpu_checkpoint.py -q -s --maxfail 20 ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/${tag}_pytest.log | tail -25 | cut -c1-300
