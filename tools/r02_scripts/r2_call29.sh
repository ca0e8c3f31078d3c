#!/bin/bash
# Round-2 GPU call 29 (2 GPUs): final tree — two-rank numeric tests and the N=2 train bench.
tag=r2c29
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_dist.py -q -s ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR|2 ranks x" gpurun_out/${tag}_pytest.log | tail -6 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "bench n2 exit $?"; grep '^{' gpurun_out/${tag}_bench_n2.json | cut -c1-260; tail -1 gpurun_out/${tag}_bench_n2.err | cut -c1-200
