"""GPU: the separable two-pass bilinear backward (npp_bilinear_bwd_sep, csrc/resample.cu — the default since the end
of round 1) against the gather-form kernel (npp_bilinear_bwd, itself checked against torch in test_gpu_ops.py;
NPP_BILINEAR_SEP=0 selects it).  Green on a B200: profiles/r01_pytest_gpu_n1_and_bilinear_sep.log."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale,align", [(2, True), (4, True), (8, True), (2, False), (0.5, True)])   # 0.5: gather both
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_separable_backward_matches_gather_form(scale, align, dtype, lib_built):
    from npp_b200 import functional as F_
    F_.set_compute_dtype(dtype)
    try:
        torch.manual_seed(3)
        x = torch.randn(2, 24, 12, 20, device="cuda")
        grads = []
        for sep in (False, True):
            F_._state["bilinear_sep"] = sep
            xi = F_.to_internal(x.clone().requires_grad_(True))
            y = F_.interpolate(xi, scale_factor=scale, mode="bilinear", align_corners=align)
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
            (dx,) = torch.autograd.grad(y, xi, gy.to(y.dtype).contiguous(memory_format=torch.channels_last))
            grads.append(dx.float())
        ref, got = grads
        tol = 1e-5 if dtype == torch.float32 else 1.6e-2
        assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()
    finally:
        F_._state["bilinear_sep"] = True
        F_.set_compute_dtype(torch.bfloat16)
