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
