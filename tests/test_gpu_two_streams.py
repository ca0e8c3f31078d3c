"""GPU: the two task streams on two CUDA streams (functional.TaskStreams, NPP_TWO_STREAMS=1) compute what the single-
stream schedule computes — eager forward / backward in fp32 validation mode, and a CUDA-graph-captured training step
(parallel graph branches, backward included) in bf16."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _fwd_bwd(two):
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.models.model_augment import Network
    F_._state["two_streams"] = two
    torch.manual_seed(0)
    net = Network(engine.make_cfg(layers=8, init_channels=16)).cuda().train()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 192, 192, generator=gen).cuda()
    pl, par = net(x)
    outs = [t for p in pl + par for t in p]
    cots = [torch.randn(t.shape, generator=gen).cuda() for t in outs]
    sum((t * c).sum() for t, c in zip(outs, cots)).backward()
    torch.cuda.synchronize()
    return [o.detach() for o in outs], {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


def test_two_streams_match_one_stream_fp32(lib_built):
    from npp_b200 import functional as F_
    F_.set_compute_dtype(torch.float32)
    keep = F_._state.get("two_streams")
    try:
        o1, g1 = _fwd_bwd(False)
        o2, g2 = _fwd_bwd(True)
        assert F_.TaskStreams.side_stream_of(torch.cuda.current_device()) is not None      # the side stream was used
        ferr = [rel(a, b) for a, b in zip(o2, o1)]
        gmax = max(g.abs().max().item() for g in g1.values())
        gerr = sorted(rel(g2[k], g) for k, g in g1.items() if g.abs().max().item() > 1e-5 * gmax)
        print("two streams vs one (fp32): forward", ["%.1e" % e for e in ferr], "grads median %.1e max %.1e" % (
            gerr[len(gerr) // 2], gerr[-1]))
        assert set(g1) == set(g2)
        assert max(ferr) < 1e-4, ferr          # same kernels, same inputs: only atomics order differs
        assert gerr[len(gerr) // 2] < 2e-2 and gerr[-1] < 0.2, (gerr[len(gerr) // 2], gerr[-1])   # cf. test_gpu_engine
    finally:
        F_._state["two_streams"] = keep
        F_.set_compute_dtype(torch.bfloat16)


def test_two_streams_graph_step(lib_built):
    """One captured training step with parallel branches: loss equals the single-stream step's to bf16 noise, replicas
    of it train (falling loss), the captured launch count is unchanged."""
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    from npp_b200.models.model_augment import Network
    keep = F_._state.get("two_streams")
    res = {}
    try:
        for two in (False, True):
            F_._state["two_streams"] = two
            F_.set_compute_dtype(torch.bfloat16)
            torch.manual_seed(1)
            model = Network(engine.make_cfg(layers=8, init_channels=16)).cuda().train()
            cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2, min_kept=2000).cuda()
            opt = engine.build_optimizer(model, cpose, cpar)
            step = engine.TrainStep(model, cpose, cpar, opt, 2, 128, use_graph=True, warmup=1)
            step.load(*engine.synthetic_batch(2, 128, seed=7))
            step.prepare()
            losses = [float(step.run()) for _ in range(8)]
            torch.cuda.synchronize()
            res[two] = (losses, step.launches_per_step)
            step.close()
        (l1, n1), (l2, n2) = res[False], res[True]
        print("one stream", l1, "two streams", l2)
        assert n1 == n2
        assert all(v == v for v in l2) and min(l2[-2:]) < l2[0]
        assert abs(l1[0] - l2[0]) < 2e-2 * abs(l1[0]), (l1[0], l2[0])
    finally:
        F_._state["two_streams"] = keep


def test_two_streams_supernet_fp32(lib_built):
    """Search supernet: forward outputs, architecture gradients (first-order in the MixedOp outputs, hence far less
    amplified than weight gradients) with the parsing stream + ParCells on the side stream."""
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.models.model_search_interact import Network
    F_.set_compute_dtype(torch.float32)
    keep = F_._state.get("two_streams")
    res = {}
    try:
        for two in (False, True):
            F_._state["two_streams"] = two
            torch.manual_seed(0)
            net = Network(engine.make_cfg(layers=8, init_channels=16)).cuda().train()
            gen = torch.Generator().manual_seed(3)
            x = torch.randn(2, 3, 128, 128, generator=gen).cuda()
            pl, par = net(x)
            outs = [t for p in pl + par for t in p]
            cots = [torch.randn(t.shape, generator=gen).cuda() for t in outs]
            (sum((t * c).sum() for t, c in zip(outs, cots)) + net.loss_entropy()).backward()
            torch.cuda.synchronize()
            res[two] = ([o.detach() for o in outs], [p.grad.detach().clone() for p in net.arch_parameters()])
        ferr = [rel(a, b) for a, b in zip(res[True][0], res[False][0])]
        aerr = [rel(a, b) for a, b in zip(res[True][1], res[False][1])]
        print("supernet two streams vs one (fp32): forward", ["%.1e" % e for e in ferr], "arch grads", ["%.1e" % e for e in aerr])
        assert max(ferr) < 1e-4, ferr
        assert sorted(aerr)[len(aerr) // 2] < 5e-2 and max(aerr) < 0.3, aerr
    finally:
        F_._state["two_streams"] = keep
        F_.set_compute_dtype(torch.bfloat16)
