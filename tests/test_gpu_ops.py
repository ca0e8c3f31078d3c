"""GPU parity of every operator primitive (fwd + bwd) against the oracle restatement of
models/operations.py — fp32 validation mode within 1e-4, bf16 product mode within 2e-2 (norm-wise
relative error, the tolerances BASELINE.json's north_star states)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _randomize(module, gen):
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("weight") and p.dim() == 1:     # BN gamma
                p.copy_(torch.rand(p.shape, generator=gen) + 0.5)
            elif p.dim() == 1:                                # biases / BN beta
                p.copy_(torch.randn(p.shape, generator=gen) * 0.2)
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=gen) * (1.5 / fan_in ** 0.5))


PRIMS = ["max_pool_3x3", "avg_pool_3x3", "skip_connect", "std_conv_3x3", "std_conv_1x1", "dil_conv_3x3_2",
         "dil_conv_3x3_4", "dil_conv_5x5_4", "se_connect", "sep_conv_3x3", "sep_conv_5x5", "poled_conv_x1",
         "poled_conv_x2", "none"]


def _oracle_run(name, sd_src, x_cpu, go, stride, dt, storage=None):
    """Oracle forward+backward in precision `dt`; `storage` turns on bf16 storage emulation."""
    from oracle import nppnet_ref as O
    sd = {k: (v.detach().clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in sd_src.items()}
    for k, v in sd.items():
        if v.is_floating_point() and v.dim() > 0 and "running" not in k:
            v.requires_grad_(True)
    x = x_cpu.clone().to(dt).requires_grad_(True)
    O.set_storage_dtype(storage)
    try:
        y = O.primitive(name, O.Params(sd, training=True), x, stride)
        (y * go.to(dt)).sum().backward()
    finally:
        O.set_storage_dtype(None)
    return y.detach(), x.grad, sd


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("name", PRIMS)
def test_primitive_parity(name, stride, dtype, lib_built):
    """Reference = the oracle in fp64.  Ours must be within the north-star tolerance (1e-4 fp32 / 2e-2 bf16) OR
    within 3x of the error the oracle itself shows at the same precision (fp32 arithmetic, resp. fp32 arithmetic
    with bf16-rounded storage) — whichever is larger; the second clause only matters for ill-conditioned
    quantities such as BatchNorm gradients over a handful of samples."""
    from npp_b200 import functional as F_
    from npp_b200.models.operations import OPS

    if name.startswith("poled_conv") and stride == 2:
        pytest.skip("Pooled_Conv is only instantiated with stride 1 (shape-inconsistent otherwise)")
    C, N, H, W = 32, 2, 24, 24
    gen = torch.Generator().manual_seed(1234 + stride)
    F_.set_compute_dtype(dtype)
    try:
        op = OPS[name](C, stride, True)
        _randomize(op, gen)
        sd_src = {k: v.detach().clone() for k, v in op.state_dict().items()}
        x_cpu = torch.randn(N, C, H, W, generator=gen)
        if dtype == torch.bfloat16:
            # feed both sides bf16-representable inputs: max-pool ties (frequent at 8 mantissa bits) then break
            # identically (first maximum) instead of being an artefact of rounding only one side
            x_cpu = x_cpu.bfloat16().float()
        # reference (fp64) and yardstick (same precision as the mode under test)
        y_tmp, _, _ = _oracle_run(name, sd_src, x_cpu, torch.zeros(()), stride, torch.float32)
        go = torch.randn(y_tmp.shape, generator=gen)
        y64, dx64, sd64 = _oracle_run(name, sd_src, x_cpu, go, stride, torch.float64)
        yy, dxy, sdy = _oracle_run(name, sd_src, x_cpu, go, stride, torch.float32,
                                   storage=torch.bfloat16 if dtype == torch.bfloat16 else None)
        # ours
        op = op.cuda().train()
        xm = x_cpu.cuda().requires_grad_(True)
        ym = F_.from_internal(op(F_.to_internal(xm, dtype)), C)
        assert ym.shape == y64.shape
        (ym * go.cuda()).sum().backward()
        tol = TOL[dtype]
        if name == "none":
            assert ym.abs().max().item() == 0 and xm.grad.abs().max().item() == 0
            return

        def check(what, mine, ref, yard, slack=1.0):
            e, ey = rel_err(mine, ref), rel_err(yard, ref)
            assert e < max(tol * slack, 3 * ey), "%s: err %.3g (yardstick %.3g)" % (what, e, ey)

        check("forward", ym, y64, yy)
        check("dx", xm.grad, dx64, dxy, 3 if dtype == torch.bfloat16 else 1)
        grads64 = [v.grad for v in sd64.values() if v.grad is not None]
        gmax = max([g.abs().max().item() for g in grads64], default=0.0)
        for k, p in op.named_parameters():
            ref = sd64[k].grad
            if ref is None:  # e.g. SE_Block.bn at stride 1 is never used (reference behaviour)
                assert p.grad is None or p.grad.abs().max().item() == 0
                continue
            if ref.abs().max().item() < 1e-5 * gmax:
                # conv bias in front of a training-mode BN: the true gradient is exactly zero, both sides are noise
                assert p.grad.abs().max().item() < (1e-2 if dtype == torch.bfloat16 else 1e-4) * gmax, \
                    "grad %s should vanish" % k
                continue
            check("grad " + k, p.grad, ref, sdy[k].grad, 5 if dtype == torch.bfloat16 else 2)
        # BatchNorm running statistics were updated identically
        for k, b in op.named_buffers():
            if "running" in k:
                assert rel_err(b, sd64[k]) < max(tol, 1e-5), "buffer %s" % k
    finally:
        F_.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("scale,align,mode", [(2, True, "bilinear"), (4, True, "bilinear"), (0.5, True, "bilinear"),
                                              (1.0, True, "bilinear"), (8, True, "bilinear"), (0.125, True, "bilinear"),
                                              (2, False, "bilinear"), (0.5, False, "bilinear"), (2, None, "nearest"),
                                              (4, None, "nearest"), (0.5, None, "nearest")])
def test_interpolate_parity(scale, align, mode, dtype, lib_built):
    import torch.nn.functional as F
    from npp_b200 import functional as F_
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(2, 16, 24, 16, generator=gen)
    xo = x.clone().requires_grad_(True)
    kw = dict(mode=mode) if mode == "nearest" else dict(mode=mode, align_corners=align)
    yo = F.interpolate(xo, scale_factor=scale, **kw)
    g = torch.randn(yo.shape, generator=gen)
    (yo * g).sum().backward()
    xm = x.cuda().requires_grad_(True)
    ym = F_.from_internal(F_.interpolate(F_.to_internal(xm, dtype), scale_factor=scale, mode=mode, align_corners=align))
    (ym * g.cuda()).sum().backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert ym.shape == yo.shape
    assert rel_err(ym, yo) < tol
    assert rel_err(xm.grad, xo.grad) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_maxpool_ties_route_to_first_max(dtype, lib_built):
    """ATen routes the gradient of tied maxima to the first element in window scan order."""
    import torch.nn.functional as F
    from npp_b200 import functional as F_
    x = torch.zeros(1, 8, 6, 6)
    x[:, :, 2:4, 2:4] = 1.0   # 2x2 plateau of equal maxima
    xo = x.clone().requires_grad_(True)
    yo = F.max_pool2d(xo, 3, 1, 1)
    yo.sum().backward()
    xm = x.cuda().requires_grad_(True)
    ym = F_.from_internal(F_.max_pool3x3(F_.to_internal(xm, dtype), 1))
    ym.sum().backward()
    assert torch.equal(ym.cpu(), yo)
    assert torch.equal(xm.grad.cpu(), xo.grad)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 128, 96, 96, 2, 1), (2, 64, 48, 48, 4, 1), (3, 32, 24, 24, 2, 1),
                                   (2, 256, 12, 12, 4, 1), (2, 64, 48, 48, 2, 2), (2, 40, 23, 17, 2, 1),
                                   (1, 72, 33, 50, 4, 2), (2, 32, 96, 96, 1, 1),
                                   (2, 16, 96, 96, 2, 1), (3, 24, 30, 22, 4, 1), (2, 8, 40, 72, 2, 1)],
                         ids=lambda s: "n%dc%d_%dx%d_d%ds%d" % s)
@pytest.mark.parametrize("relu_in", [0, 1])
def test_depthwise_conv_vs_torch(shape, relu_in, dtype, lib_built):
    """Depthwise dilated 3x3 (DilConvS.net[1], operations.py:213) at the network's real shapes — the shared-memory
    tiled kernels (multi-tile images, partial channel blocks, ragged tiles; 8 / 4 / 2 channel-vector lanes per pixel
    for wide layers / the supernet's 32- and 16-channel MixedOp slices) and the gather fallback (stride 2) —
    forward, input gradient and weight gradient against torch's grouped convolution on the same rounded inputs."""
    import torch.nn.functional as TF
    from npp_b200 import functional as F_
    n, c, h, w, dil, stride = shape
    gen = torch.Generator().manual_seed(h * 131 + c + dil)
    x = torch.randn(n, c, h, w, generator=gen)
    wt = torch.randn(c, 1, 3, 3, generator=gen) * 0.3
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    xr = x.double().requires_grad_(True)
    wr = wt.double().requires_grad_(True)
    yr = TF.conv2d(torch.relu(xr) if relu_in else xr, wr, None, stride, dil, dil, groups=c)
    go = torch.randn(yr.shape, generator=gen)
    if dtype == torch.bfloat16:
        go = go.bfloat16().float()
    (yr * go.double()).sum().backward()
    xm = F_.to_internal(x.cuda(), dtype).detach().requires_grad_(True)
    wm = wt.cuda().requires_grad_(True)
    ym = F_.dwconv2d(xm, wm, stride, dil, dil, relu_in)
    ym.backward(F_.to_internal(go.cuda(), dtype))
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel_err(F_.from_internal(ym), yr) < tol
    assert rel_err(F_.from_internal(xm.grad), xr.grad) < tol
    assert rel_err(wm.grad, wr.grad) < (1e-4 if dtype == torch.float32 else 2e-3)


def test_depthwise_conv_on_channel_slice(lib_built):
    """The supernet's MixedOp feeds depthwise convs a channel-slice view (model_search_interact.py:59)."""
    import torch.nn.functional as TF
    from npp_b200 import functional as F_
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 40, 40, generator=gen).bfloat16().float()
    wt = torch.randn(32, 1, 3, 3, generator=gen) * 0.3
    xm = F_.to_internal(x.cuda(), torch.bfloat16)
    lo = F_.alias(xm, 0, 32)
    hi = F_.alias(xm, 32, 32)
    for view, ref in ((lo, x[:, :32]), (hi, x[:, 32:])):
        y = F_.dwconv2d(view, wt.cuda(), 1, 2, 2, 1)
        want = TF.conv2d(torch.relu(ref), wt, None, 1, 2, 2, groups=32)
        assert rel_err(F_.from_internal(y), want) < 6e-3
