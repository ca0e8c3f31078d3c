"""GPU: the cross-entropy backward as "per-pixel gradient + separable bilinear backward" (npp_par_loss_grad_pixels /
npp_edge_loss_grad_pixels + npp_bilinear_bwd_sep, NPP_CE_BWD_SEP=1 — a round-2 candidate, NOT on the default path)
against the default shared-memory-atomic kernels (checked against the reference fixtures in test_gpu_golden.py):
same logit gradients up to fp32 summation order.  Written after round 1's GPU budget was spent, hence the non-strict
xfail (a pass shows up as XPASS)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="candidate path not yet run on a B200 (written after the round-1 GPU budget)",
                                strict=False)]


@pytest.mark.parametrize("shape", [(2, 24, 24, 96, 96), (3, 20, 28, 80, 112)])
def test_criterion_par_backward_variants_agree(shape, lib_built):
    from npp_b200 import functional as F_
    from npp_b200.core.criterion import Criterion_par
    n, h, w, lh, lw = shape
    gen = torch.Generator(device="cuda").manual_seed(9)
    par = [torch.randn(n, 20, h, w, device="cuda", generator=gen) for _ in range(2)]
    edge = [torch.randn(n, 2, h, w, device="cuda", generator=gen) for _ in range(2)]
    lab = torch.randint(0, 20, (n, lh, lw), device="cuda", generator=gen)
    lab[:, :4, :] = 255
    lab[:, :, -3:] = 255
    elab = (torch.rand(n, lh, lw, device="cuda", generator=gen) < 0.1).long()
    elab[lab == 255] = 255
    crit = Criterion_par(out_len=2, min_kept=500).cuda()
    res = []
    try:
        for sep in (False, True):
            F_._state["ce_bwd_sep"] = sep
            ins = [t.clone().requires_grad_(True) for t in par + edge]
            loss = crit([[ins[0], ins[2]], [ins[1], ins[3]]], [lab, elab])
            res.append((float(loss), torch.autograd.grad(loss, ins)))
    finally:
        F_._state["ce_bwd_sep"] = False
    (l0, g0), (l1, g1) = res
    assert abs(l0 - l1) <= 1e-5 * abs(l0)     # the forward sums use fp32 atomics
    for a, b in zip(g0, g1):
        assert (a - b).abs().max() <= 1e-4 * a.abs().max() + 1e-9
