#!/bin/bash
# Round-2 GPU call 8 (2 GPUs): push-form peer exchange + PDL with early trigger — 2-rank numeric tests, exchange latency,
# N=2 and N=1 bench, ncu --set full of the C=32 3x3 kernel.
tag=r2c8
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_dist.py -q -s ) > gpurun_out/${tag}_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_dist.log
grep -E "passed|failed|FAILED|ERROR|2 ranks x" gpurun_out/${tag}_pytest_dist.log | tail -8 | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    tools/bench_peer.py > gpurun_out/${tag}_peer_latency.txt 2> gpurun_out/${tag}_peer_latency.err
grep floats gpurun_out/${tag}_peer_latency.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-kernel-table > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
echo "bench n2 exit $?"; grep '^{' gpurun_out/${tag}_bench_n2.json | cut -c1-260; tail -2 gpurun_out/${tag}_bench_n2.err | cut -c1-200
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench n1 exit $?"; grep '^{' gpurun_out/${tag}_bench_n1.json | cut -c1-260
NPP_PDL=0 timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_n1_pdl0.json 2> gpurun_out/${tag}_bench_n1_pdl0.err
echo "bench n1 pdl0 exit $?"; grep '^{' gpurun_out/${tag}_bench_n1_pdl0.json | cut -c1-260
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv3_kernel -c 2 -o gpurun_out/${tag}_full_conv3_32 -f python tools/profile_step.py > gpurun_out/${tag}_ncu_conv3.log 2>&1
ncu -i gpurun_out/${tag}_full_conv3_32.ncu-rep --page raw --csv > gpurun_out/${tag}_full_conv3_32_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full_conv3_32.ncu-rep --page details > gpurun_out/${tag}_full_conv3_32_details.txt 2>/dev/null
ncu -i gpurun_out/${tag}_full_conv3_32.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_full_conv3_32_source.csv.gz
rm -f gpurun_out/${tag}_full_conv3_32.ncu-rep
grep -E "conv3_kernel|Duration" gpurun_out/${tag}_full_conv3_32_details.txt | head -6
