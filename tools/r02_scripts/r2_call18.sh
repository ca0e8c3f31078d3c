#!/bin/bash
# Round-2 GPU call 18 (8 GPUs): push-protocol SyncBN exchange latency at 8 ranks + train bench at N=8 (final build).
tag=r2c18
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
    tools/bench_peer.py > gpurun_out/${tag}_peer_latency_8gpu.txt 2> gpurun_out/${tag}_peer_latency_8gpu.err
echo "peer exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/${tag}_peer_latency_8gpu.txt | tail -8 | cut -c1-200
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-kernel-table > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
echo "bench n8 exit $?"; grep '^{' gpurun_out/${tag}_bench_n8.json | cut -c1-300; tail -2 gpurun_out/${tag}_bench_n8.err | cut -c1-200
