"""CPU, world_size 2 over gloo: the host side of the data-parallel path — the SyncBN exchange
protocol (raw per-channel sums forward, (sum g, sum g*xhat) backward), the flat-gradient all-reduce of
engine.TrainStep, rank-sharded synthetic batches and the integer evaluation all-reduce.  The arithmetic each rank
applies to the reduced vectors is restated from csrc/bn.cu (bn_finalize_kernel) / csrc/node.cu (bn_bwd_coef) and
checked against torch's BatchNorm over the concatenated batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        out[rank] = "ok"
    except Exception as e:  # surfaced by the parent
        import traceback
        out[rank] = "".join(traceback.format_exception(type(e), e, e.__traceback__))
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, out), nprocs=world, join=True)
    for r in range(world):
        assert out.get(r) == "ok", "rank %d: %s" % (r, out.get(r))


# ------------------------------------------------------------------------------------------------ workers
def _w_sync_bn_protocol(rank, world):
    from npp_b200 import distributed as npp_dist
    from npp_b200 import functional as F_
    assert npp_dist.world() == (rank, world)
    assert F_._sync_group() is None                     # off by default
    npp_dist.enable_sync_bn(True)
    try:
        assert F_._sync_group() is True
        gen = torch.Generator().manual_seed(7)          # both ranks generate the FULL batch, each keeps its shard
        full = torch.randn(2 * 3, 5, 4, 4, generator=gen, dtype=torch.float64)
        gy = torch.randn(full.shape, generator=gen, dtype=torch.float64)
        gamma = torch.rand(5, generator=gen, dtype=torch.float64) + 0.5
        beta = torch.randn(5, generator=gen, dtype=torch.float64)
        x, g = full[rank * 3:(rank + 1) * 3], gy[rank * 3:(rank + 1) * 3]
        # forward: local raw sums -> ONE all-reduce of 2C floats -> finalize (bn_finalize_kernel arithmetic)
        stats = torch.cat([x.sum((0, 2, 3)), (x * x).sum((0, 2, 3))])
        count = float(x.numel() // 5)
        count *= F_._allreduce_sum(stats)
        mean = stats[:5] / count
        var = (stats[5:] / count - mean * mean).clamp_min(0)
        invstd = 1.0 / torch.sqrt(var + 1e-5)
        y = (x - mean.view(1, -1, 1, 1)) * (gamma * invstd).view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
        fr = full.clone().requires_grad_(True)
        want = torch.nn.functional.batch_norm(fr, None, None, gamma, beta, True, 0.1, 1e-5)
        assert torch.allclose(y, want[rank * 3:(rank + 1) * 3].detach(), atol=1e-10)
        # backward: local (sum g, sum g*xhat) -> all-reduce -> dx (bn_bwd_coef arithmetic); d gamma / d beta stay local
        xhat = (x - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
        sums = torch.cat([g.sum((0, 2, 3)), (g * xhat).sum((0, 2, 3))])
        local = sums.clone()
        F_._allreduce_sum(sums)
        dx = (gamma * invstd).view(1, -1, 1, 1) * (g - sums[:5].view(1, -1, 1, 1) / count
                                                   - xhat * sums[5:].view(1, -1, 1, 1) / count)
        (want * gy).sum().backward()
        assert torch.allclose(dx, fr.grad[rank * 3:(rank + 1) * 3], atol=1e-10)
        # the parameter gradients of the full batch are the SUM of the ranks' local ones (DDP then averages)
        tot = local.clone()
        dist.all_reduce(tot)
        xh_full = (full - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
        assert torch.allclose(tot[:5], gy.sum((0, 2, 3)), atol=1e-10)
        assert torch.allclose(tot[5:], (gy * xh_full).sum((0, 2, 3)), atol=1e-10)
    finally:
        npp_dist.enable_sync_bn(None)
    assert F_._sync_group() is None
    model = torch.nn.Linear(2, 2)
    assert npp_dist.convert_sync_batchnorm(model) is model and F_._sync_group() is True
    npp_dist.enable_sync_bn(None)


def _w_flat_gradient_allreduce(rank, world):
    from npp_b200 import engine

    class _Shell(engine.TrainStep):      # the gradient exchange of TrainStep without its CUDA buffers
        def __init__(self, flat, world):
            self.flat_grads, self.world_size = flat, world

    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    _Shell(flat, world)._allreduce_grads()
    assert torch.equal(flat, torch.arange(10, dtype=torch.float32) * 1.5)   # mean over ranks of (1x, 2x)
    # the per-parameter fallback (no flat buffer) gives the same averages
    ps = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2, 2))]
    for i, p in enumerate(ps):
        p.grad = torch.full_like(p, float(rank + 1 + i))
    sh = _Shell(None, world)
    sh.opt = torch.optim.SGD(ps, lr=0.1)
    sh._allreduce_grads()
    assert torch.equal(ps[0].grad, torch.full((3,), 1.5)) and torch.equal(ps[1].grad, torch.full((2, 2), 2.5))


def _w_sharded_batches_and_eval(rank, world):
    from npp_b200 import engine
    from oracle import eval_ref as E
    mine = engine.synthetic_batch(2, 64, seed=1 + rank)
    again = engine.synthetic_batch(2, 64, seed=1 + rank)
    assert all(torch.equal(a, b) for a, b in zip(mine, again))            # deterministic per rank
    digest = torch.tensor([float(mine[0].double().sum())], dtype=torch.float64)
    both = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(both, digest)
    assert both[0].item() != both[1].item()                                # ranks see different images
    img, par = mine[0], mine[1]
    assert img.shape == (2, 3, 64, 64) and par.dtype == torch.int64 and int(par.max()) == 255
    # evaluation: per-rank int64 confusion histograms all-reduced == histogram of the union (bit-exact)
    rng = np.random.RandomState(5)
    logits = rng.randn(4, 7, 16, 16).astype(np.float32)
    label = rng.randint(0, 7, size=(4, 16, 16)).astype(np.int64)
    label[:, :2] = 255
    sl = slice(rank * 2, rank * 2 + 2)
    local = E.confusion_matrix(label[sl], logits[sl], (2, 7, 16, 16), 7, 255).astype(np.int64)
    t = torch.from_numpy(local)
    dist.all_reduce(t)
    assert np.array_equal(t.numpy(), E.confusion_matrix(label, logits, (4, 7, 16, 16), 7, 255).astype(np.int64))


# ------------------------------------------------------------------------------------------------ tests
def test_sync_bn_exchange_protocol_world2():
    _run("_w_sync_bn_protocol")


def test_flat_gradient_allreduce_world2():
    _run("_w_flat_gradient_allreduce")


def test_sharded_batches_and_eval_allreduce_world2():
    _run("_w_sharded_batches_and_eval")


def test_reference_arm_only_runs_on_rank0(monkeypatch, capsys):
    """bench.py --impl reference under torchrun: ranks other than 0 return without work or output."""
    sys.path.insert(0, ROOT)
    import bench
    import types
    monkeypatch.setenv("RANK", "1")
    bench.run_reference_arm(types.SimpleNamespace(steps=1, warmup=0, gpus=2))
    assert capsys.readouterr().out == ""
