"""GPU: FusedAdam (one multi-tensor kernel) follows torch.optim.Adam step for step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch(lib_built):
    from npp_b200.optim import FusedAdam
    gen = torch.Generator().manual_seed(0)
    shapes = [(64, 3, 3, 3), (17,), (128, 64, 1, 1), (40000,), (1,), (256, 128, 3, 3)]
    pa = [torch.randn(s, generator=gen).cuda().requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedAdam([{"params": pa[:3], "lr": 0.2 * 0.0015}, {"params": pa[3:]}], 0.0015)
    ob = torch.optim.Adam([{"params": pb[:3], "lr": 0.2 * 0.0015}, {"params": pb[3:]}], 0.0015)
    for it in range(5):
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, generator=gen).cuda()
            a.grad, b.grad = g.clone(), g.clone()
        pa[1].grad = None if it == 2 else pa[1].grad   # a parameter without gradient is skipped, like torch
        pb[1].grad = None if it == 2 else pb[1].grad
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-7), (a - b).abs().max()


def test_fused_adam_resume_matches_torch(lib_built):
    """save -> load -> step (augment_lip_sync.py:235 optimizer.load_state_dict): the loaded step counters and moments
    are used (bias correction continues), for state dicts written by FusedAdam and by torch.optim.Adam."""
    from npp_b200.optim import FusedAdam
    gen = torch.Generator().manual_seed(1)
    shapes = [(32, 3, 3, 3), (9,), (5000,)]
    pa = [torch.randn(s, generator=gen).cuda().requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa, ob = FusedAdam(pa, 0.01), torch.optim.Adam(pb, 0.01)

    def step_both(oa, ob, pa, pb):
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, generator=gen).cuda()
            a.grad, b.grad = g.clone(), g.clone()
        oa.step()
        ob.step()

    for _ in range(7):
        step_both(oa, ob, pa, pb)
    import copy
    sd_a, sd_b = copy.deepcopy(oa.state_dict()), copy.deepcopy(ob.state_dict())    # as torch.load would hand them over
    for src in (sd_a, sd_b):                        # resume from our own and from torch's checkpoint format
        pa2 = [p.detach().clone().requires_grad_(True) for p in pa]
        pb2 = [p.detach().clone().requires_grad_(True) for p in pb]
        oa2, ob2 = FusedAdam(pa2, 0.01), torch.optim.Adam(pb2, 0.01)
        oa2.load_state_dict(copy.deepcopy(src))
        ob2.load_state_dict(copy.deepcopy(sd_b))
        for _ in range(3):
            step_both(oa2, ob2, pa2, pb2)
        assert int(oa2.state[pa2[0]]["step"]) == 10
        for a, b in zip(pa2, pb2):
            assert torch.allclose(a, b, rtol=2e-5, atol=1e-7), (a - b).abs().max()
    # loading AFTER a step must drop the stale device table (new moment tensors)
    oa.load_state_dict(copy.deepcopy(sd_a))
    ob.load_state_dict(copy.deepcopy(sd_b))
    for a, b in zip(pa, pb):       # same parameters again (the comparison below starts from one point)
        b.data.copy_(a.data)
    step_both(oa, ob, pa, pb)
    assert int(oa.state[pa[0]]["step"]) == 8
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-7), (a - b).abs().max()


def test_fused_adam_lr_change_reaches_the_kernel_in_place(lib_built):
    """param_group['lr'] changes (MultiStepLR, augment_lip_sync.py:213,249) are written into the device table in
    place — same device memory, so a captured CUDA graph sees them on its next replay."""
    from npp_b200.optim import FusedAdam
    p = torch.ones(1000).cuda().requires_grad_(True)
    opt = FusedAdam([p], lr=0.1)
    p.grad = torch.ones_like(p)
    opt.step()
    tbl_ptr = opt._tables[0].data_ptr()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        opt.step()
    before = p.detach().clone()
    graph.replay()
    torch.cuda.synchronize()
    d1 = (before - p.detach()).mean().item()
    opt.param_groups[0]["lr"] = 0.001
    assert opt.sync_hyperparameters() and opt._tables[0].data_ptr() == tbl_ptr
    before = p.detach().clone()
    graph.replay()
    torch.cuda.synchronize()
    d2 = (before - p.detach()).mean().item()
    assert 0.05 < d1 < 0.15 and 0.0005 < d2 < 0.0015, (d1, d2)
