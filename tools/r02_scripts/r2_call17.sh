#!/bin/bash
# Round-2 GPU call 17 (2 GPUs): final conv kernels + worker-stream branches with per-worker peer communicators under SyncBN:
# 2-rank numeric tests, two-stream tests, N=2 and N=1 bench on the same box.
tag=r2c17
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_two_streams.py tests/test_gpu_network.py tests/test_gpu_engine.py -q -s ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|FAILED|ERROR|2 ranks x|two streams vs" gpurun_out/${tag}_pytest.log | tail -10 | cut -c1-300
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "bench n2 exit $?"; grep '^{' gpurun_out/${tag}_bench_n2.json | cut -c1-260; tail -2 gpurun_out/${tag}_bench_n2.err | cut -c1-200
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench n1 exit $?"; grep '^{' gpurun_out/${tag}_bench_n1.json | cut -c1-260
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table --workload search > gpurun_out/${tag}_bench_search_n2.json 2> gpurun_out/${tag}_bench_search_n2.err
echo "bench search n2 exit $?"; grep '^{' gpurun_out/${tag}_bench_search_n2.json | cut -c1-260; tail -2 gpurun_out/${tag}_bench_search_n2.err | cut -c1-200
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table --workload search > gpurun_out/${tag}_bench_search_n1.json 2> gpurun_out/${tag}_bench_search_n1.err
echo "bench search n1 exit $?"; grep '^{' gpurun_out/${tag}_bench_search_n1.json | cut -c1-260
