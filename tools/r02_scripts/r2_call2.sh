#!/bin/bash
# Round-2 GPU call 2 (2 GPUs): 2-rank numeric tests (peer-memory SyncBN exchange), N=2 bench with the peer transport
# and with the NCCL transport (A/B).
tag=r2c2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${tag}_smi.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_dist.py -q -s -x ) > gpurun_out/${tag}_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_dist.log
tail -15 gpurun_out/${tag}_pytest_dist.log | cut -c1-400
run_bench () {  # name, env...
  name=$1; shift
  env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "bench $name exit $?"; cut -c1-260 gpurun_out/${tag}_bench_$name.json; tail -3 gpurun_out/${tag}_bench_$name.err | cut -c1-300
}
run_bench peer NPP_SYNCBN_PEER=1
run_bench nccl NPP_SYNCBN_PEER=0
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "bench n1 exit $?"; cut -c1-260 gpurun_out/${tag}_bench_n1.json
