"""GPU: the pixel-pair form of the 32 -> 32 channel 3x3 convolutions (functional._ConvPairFn, NPP_CONV_PAIR=1 — a
round-2 candidate that is NOT on the default path) against the default kernels on the same inputs: output, fused
BatchNorm sums, data gradient and weight gradient.  The index mapping itself (super-pixel weights, gradient fold,
statistics fold) is checked in float64 on the CPU in tests/test_cpu_conv_pair.py.  Written after round 1's GPU budget
was spent, hence the non-strict xfail (a pass shows up as XPASS)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="candidate path not yet run on a B200 (written after the round-1 GPU budget)",
                                strict=False)]


@pytest.mark.parametrize("shape", [(4, 24, 40), (2, 96, 96), (3, 17, 22)])
def test_pair_conv_matches_default(shape, lib_built):
    from npp_b200 import functional as F_
    F_.set_compute_dtype(torch.bfloat16)
    n, h, w = shape
    torch.manual_seed(1)
    x32 = torch.randn(n, 32, h, w, device="cuda").bfloat16().float()
    wgt = (torch.randn(32, 32, 3, 3, device="cuda") * 0.1).requires_grad_(True)
    gy = torch.randn(n, 32, h, w, device="cuda").bfloat16().float()
    res = []
    try:
        for pair in (False, True):
            F_._state["conv_pair"] = pair
            x = F_.to_internal(x32.clone().requires_grad_(True))
            y, stats = F_.conv2d(x, wgt, None, 1, 1, 1, want_stats=True)
            dx, dw = torch.autograd.grad((y.float() * gy).sum(), [x, wgt])
            res.append((y.float(), stats.clone(), dx.float(), dw.clone()))
    finally:
        F_._state["conv_pair"] = False
    (y0, s0, dx0, dw0), (y1, s1, dx1, dw1) = res
    assert (y1 - y0).abs().max() <= 2e-2 * y0.abs().max()
    assert (s1 - s0).abs().max() <= 2e-3 * s0.abs().max()
    assert (dx1 - dx0).abs().max() <= 2e-2 * dx0.abs().max()
    assert (dw1 - dw0).abs().max() <= 5e-3 * dw0.abs().max()
