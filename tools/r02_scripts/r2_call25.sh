#!/bin/bash
# Round-2 GPU call 25 (8 GPUs): train bench at N=8, final build.
tag=r2c25
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-kernel-table > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
echo "bench n8 exit $?"; grep '^{' gpurun_out/${tag}_bench_n8.json | cut -c1-300; tail -2 gpurun_out/${tag}_bench_n8.err | cut -c1-200
