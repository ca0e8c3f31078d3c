"""Latency of the SyncBN peer-memory exchange (csrc/peer.cu) against an NCCL all-reduce of the same vector, per call,
inside a CUDA graph (how the training step issues them).  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_peer.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npp_b200 import distributed as npp_dist  # noqa: E402
from npp_b200 import functional as F_  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
npp_dist.enable_sync_bn(True)
comm = npp_dist.peer_comm()
K = 500
for n in (128, 512, 2048, 4096):
    x = torch.randn(n, device="cuda")
    res = {}
    for name, fn in (("peer", lambda: comm.allreduce(x)), ("peer2", lambda: comm.allreduce(x[:n // 2], x[n // 2:])),
                     ("nccl", lambda: dist.all_reduce(x))):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(K):
                fn()
        dist.barrier()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[name] = 1e3 * e0.elapsed_time(e1) / K
        g.reset()
    if rank == 0:
        print("n=%5d floats, %d ranks: peer %.2f us, peer(2 vectors) %.2f us, nccl %.2f us per all-reduce" % (
            n, dist.get_world_size(), res["peer"], res["peer2"], res["nccl"]), flush=True)
comm.check()
npp_dist.enable_sync_bn(None)
dist.destroy_process_group()
