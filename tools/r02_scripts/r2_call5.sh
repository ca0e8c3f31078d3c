#!/bin/bash
# Round-2 GPU call 5 (8 GPUs): latency of the peer-memory exchange vs NCCL at 8 ranks, train bench at N=8.
tag=r2c5
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    tools/bench_peer.py > gpurun_out/${tag}_peer_latency.txt 2> gpurun_out/${tag}_peer_latency.err
echo "peer latency exit $?"; grep "floats" gpurun_out/${tag}_peer_latency.txt; tail -2 gpurun_out/${tag}_peer_latency.err | cut -c1-300
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-kernel-table > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
echo "bench n8 exit $?"; grep '^{' gpurun_out/${tag}_bench_n8.json | cut -c1-700; tail -3 gpurun_out/${tag}_bench_n8.err | cut -c1-300
