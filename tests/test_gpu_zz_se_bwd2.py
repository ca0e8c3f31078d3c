"""GPU: the two-kernel SE bottleneck backward without weight-gradient atomics (npp_se_fc_bwd2, NPP_SE_BWD2=1 — a
round-2 candidate, NOT on the default path) against the default one-kernel backward (checked against torch in
test_gpu_ops.py): same dx and parameter gradients up to fp32 summation order.  Written after round 1's GPU budget was
spent, hence the non-strict xfail (a pass shows up as XPASS)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="candidate kernel not yet run on a B200 (written after the round-1 GPU budget)",
                                strict=False)]


@pytest.mark.parametrize("c", [32, 256])
def test_se_backward_variants_agree(c, lib_built):
    from npp_b200 import functional as F_
    F_.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(2)
    n, h, w = 8, 12, 12
    x32 = torch.randn(n, c, h, w, device="cuda").bfloat16().float()
    params = [(torch.randn(c // 2, c, 1, 1, device="cuda") * 0.2).requires_grad_(True),
              (torch.randn(c // 2, device="cuda") * 0.1).requires_grad_(True),
              (torch.randn(c, c // 2, 1, 1, device="cuda") * 0.2).requires_grad_(True),
              (torch.randn(c, device="cuda") * 0.1).requires_grad_(True)]
    gy = torch.randn(n, c, h, w, device="cuda").bfloat16().float()
    res = []
    try:
        for v2 in (False, True):
            F_._state["se_bwd2"] = v2
            x = F_.to_internal(x32.clone().requires_grad_(True))
            y = F_.se_scale(x, *params)
            res.append([g.float() for g in torch.autograd.grad((y.float() * gy).sum(), [x] + params)])
    finally:
        F_._state["se_bwd2"] = False
    for k, (a, b) in enumerate(zip(*res)):
        tol = 1e-2 if k == 0 else 1e-3     # dx is stored in bf16 (one ulp = 0.8 %), the parameter gradients in fp32
        assert (a - b).abs().max() <= tol * a.abs().max() + 1e-6
