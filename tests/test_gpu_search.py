"""GPU parity of the search supernet path (models/model_search_interact.py): the fused MixedOp / weighted-node
kernels (csrc/mix.cu) against plain torch, MixedOp against the oracle restatement, and the whole supernet —
outputs, architecture gradients, weight gradients — against the fixture generated from the reference itself."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel(a, b):
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


class _FakeBN:
    """The attributes functional._bn_forward_coef reads from a BatchNorm2d(affine=False)."""

    def __init__(self, c):
        from npp_b200.nn import BatchNorm2d
        self.m = BatchNorm2d(c, affine=False).cuda().train()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("interleave", [False, True])
@pytest.mark.parametrize("k", [1, 3, 7, 8])
def test_mix_kernels_vs_torch(k, interleave, dtype, lib_built):
    """out = sum_k w_k * f_k(y_k) (+ interleave with a pass-through half): forward, d w, d pass, d y_k."""
    from npp_b200 import functional as F_
    gen = torch.Generator().manual_seed(10 * k + int(interleave))
    n, c, h, w = 3, 24, 9, 7
    bn_mask = [(j % 2 == 0) for j in range(k)]           # alternate BatchNorm / plain branches
    ys = [(torch.randn(n, c, h, w, generator=gen) * (1 + j) + 0.3 * j) for j in range(k)]
    if dtype == torch.bfloat16:
        ys = [y.bfloat16().float() for y in ys]
    wts = torch.rand(k, generator=gen) + 0.1
    pas = torch.randn(n, c, h, w, generator=gen)
    if dtype == torch.bfloat16:
        pas = pas.bfloat16().float()
    go = torch.randn(n, 2 * c if interleave else c, h, w, generator=gen)
    # torch fp64 reference
    ys_r = [y.double().requires_grad_(True) for y in ys]
    w_r = wts.double().requires_grad_(True)
    p_r = pas.double().requires_grad_(True)
    tot = 0
    for j in range(k):
        t = TF.batch_norm(ys_r[j], None, None, None, None, True, 0.1, 1e-5) if bn_mask[j] else ys_r[j]
        tot = tot + w_r[j] * t
    if interleave:
        cat = torch.cat([tot, p_r], 1)
        tot = cat.view(n, 2, c, h, w).transpose(1, 2).reshape(n, 2 * c, h, w)
    (tot * go.double()).sum().backward()
    # ours
    F_.set_compute_dtype(dtype)
    try:
        ys_m = [F_.to_internal(y.cuda(), dtype).detach().requires_grad_(True) for y in ys]
        w_m = wts.cuda().requires_grad_(True)
        p_m = F_.to_internal(pas.cuda(), dtype).detach().requires_grad_(True)
        branches = []
        for j in range(k):
            if bn_mask[j]:
                branches.append(_FakeBN(c).m.pending(ys_m[j]))
            else:
                branches.append(ys_m[j])
        out = F_.mix(branches, w_m, pass_=p_m if interleave else None)
        o = F_.from_internal(out)
        (o * go.cuda()).sum().backward()
        torch.cuda.synchronize()
        tol = TOL[dtype]
        assert rel(o, tot) < tol, ("fwd", rel(o, tot))
        assert rel(w_m.grad, w_r.grad) < 5 * tol, ("dw", rel(w_m.grad, w_r.grad))
        if interleave:
            assert rel(F_.from_internal(p_m.grad), p_r.grad) < tol
        for j in range(k):
            e = rel(F_.from_internal(ys_m[j].grad), ys_r[j].grad)
            assert e < (5 * tol if bn_mask[j] else tol), ("dy", j, e)
    finally:
        F_.set_compute_dtype(torch.bfloat16)


def test_split_fanout_interleave(lib_built):
    from npp_b200 import functional as F_
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(2, 32, 6, 5, generator=gen).bfloat16().float()
    xm = F_.to_internal(x.cuda(), torch.bfloat16).detach().requires_grad_(True)
    los, hi = F_.split_halves(xm, 3)
    assert torch.equal(F_.from_internal(los[0]), x[:, :16].cuda()) and torch.equal(F_.from_internal(hi), x[:, 16:].cuda())
    fans = F_.fanout(hi, 2)
    y = F_.interleave2(F_.sum_n([los[0], los[1], fans[0]]), F_.sum_n([los[2], fans[1]]))
    ref_lo, ref_hi = x[:, :16], x[:, 16:]
    a, b = (2 * ref_lo + ref_hi).bfloat16().float(), (ref_lo + ref_hi).bfloat16().float()
    want = torch.stack([a, b], 2).reshape(2, 32, 6, 5)
    assert torch.equal(F_.from_internal(y), want.cuda())
    go = torch.randn(2, 32, 6, 5, generator=gen).bfloat16().float()
    (F_.from_internal(y) * go.cuda()).sum().backward()
    ga, gb = go[:, 0::2], go[:, 1::2]
    want_dx = torch.cat([(2 * ga + gb), (ga + gb)], 1)
    assert rel(F_.from_internal(xm.grad), want_dx) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("up_scale,extra", [(None, False), (1.0, False), (2, True), (0.5, True), (0.25, True)])
def test_mixed_op_vs_oracle(up_scale, extra, dtype, lib_built):
    """MixedOp (model_search_interact.py:39-74) forward, input gradient, d alpha and weight gradients."""
    from npp_b200 import functional as F_
    from npp_b200.models.model_search_interact import MixedOp
    from npp_b200.nn import Conv2d
    from oracle import nppnet_ref as O
    gen = torch.Generator().manual_seed(17)
    C, n, h, w = 32, 4, 16, 16
    torch.manual_seed(5)
    op = MixedOp(C, 1, up_scale, Conv2d(C, 48, 1) if extra else None)
    with torch.no_grad():
        for p in op.parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=gen) * (1.5 / p[0].numel() ** 0.5))
            else:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.2)
    x = torch.randn(n, C, h, w, generator=gen)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    alpha = torch.randn(7, generator=gen)
    sd64 = {k: (v.detach().clone().double() if v.is_floating_point() else v.clone()) for k, v in op.state_dict().items()}
    for k, v in sd64.items():
        if v.is_floating_point() and v.dim() > 0 and "running" not in k:
            v.requires_grad_(True)
    xr = x.double().requires_grad_(True)
    ar = alpha.double().requires_grad_(True)
    yr = O.mixed_op(O.Params(sd64, True), xr, torch.softmax(ar, -1), up_scale, extra)
    go = torch.randn(yr.shape, generator=gen)
    (yr * go.double()).sum().backward()
    F_.set_compute_dtype(dtype)
    try:
        op = op.cuda().train()
        xm = x.cuda().requires_grad_(True)
        am = alpha.cuda().requires_grad_(True)
        ym = F_.from_internal(op(F_.to_internal(xm, dtype), torch.softmax(am, -1)), yr.shape[1])
        (ym * go.cuda()).sum().backward()
        torch.cuda.synchronize()
        tol = TOL[dtype]
        assert ym.shape == yr.shape
        assert rel(ym, yr) < tol, ("fwd", rel(ym, yr))
        assert rel(xm.grad, xr.grad) < 5 * tol, ("dx", rel(xm.grad, xr.grad))
        assert rel(am.grad, ar.grad) < 5 * tol, ("dalpha", rel(am.grad, ar.grad))
        worst = 0.0
        for k, p in op.named_parameters():
            ref = sd64[k].grad
            if ref is None or p.grad is None or ref.abs().max() < 1e-6:
                continue
            worst = max(worst, rel(p.grad, ref))
        assert worst < (1e-3 if dtype == torch.float32 else 0.1), ("dparam", worst)
    finally:
        F_.set_compute_dtype(torch.bfloat16)


def _search_cfg(L, C):
    ns = types.SimpleNamespace
    return ns(DATASET=ns(NUM_CLASSES=20, NUM_JOINTS=16), SEARCH=ns(LAYERS=L, INIT_CHANNELS=C),
              MODEL=ns(DECONV_WITH_BIAS=False, HEAD="PSP", REFINE_LAYERS=1))


def test_supernet_golden_fp32(lib_built):
    """Whole supernet in fp32 validation mode vs the reference-generated fixture (tests/golden/make_golden.py
    golden_search): outputs 1e-4, d(alphas, betas) and sampled weight gradients of sum_i <out_i, r_i> + 3*entropy."""
    from npp_b200 import functional as F_
    from npp_b200.models.model_search_interact import Network
    g = np.load(os.path.join(HERE, "golden", "search_golden.npz"), allow_pickle=False)
    F_.set_compute_dtype(torch.float32)
    try:
        torch.manual_seed(int(g["seed"]))
        net = Network(_search_cfg(int(g["layers"]), int(g["channels"])))
        arch = [k[5:] for k in g.files if k.startswith("arch/")]
        with torch.no_grad():
            for k in arch:
                getattr(net, k).copy_(torch.from_numpy(g["arch/" + k]))
        net = net.cuda().train()
        pl, par = net(torch.from_numpy(g["x"]).cuda())
        names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
        gr = torch.Generator().manual_seed(321)
        loss = 0
        for n, t in zip(names, [t for pair in pl + par for t in pair]):
            assert rel(t, g["out/" + n]) < 1e-4, (n, rel(t, g["out/" + n]))
            loss = loss + (t * torch.randn(t.shape, generator=gr).cuda()).sum()
        ent = net.loss_entropy()
        assert abs(float(ent) - float(g["entropy"][0])) < 1e-6
        (loss + 3.0 * ent).backward()
        torch.cuda.synchronize()
        assert abs(float(loss + 3.0 * ent) - float(g["loss"][0])) < 1e-3 * abs(float(g["loss"][0])) + 1e-2
        # gradients: fp32 against the reference's own fp32 run.  This tiny random-init net with 32-sample BatchNorms is
        # ill-conditioned — the fp32 oracle itself sits ~4e-3 from its fp64 run (tests/test_gpu_network.py) — so two
        # correct fp32 evaluations with different summation orders differ at the 1e-2 level in the worst tensor.
        errs = {k: rel(getattr(net, k).grad, g["grad/" + k]) for k in arch}
        sd = dict(net.named_parameters())
        werrs = {k: rel(sd[k].grad, g["wgrad/" + k]) for k in [f[6:] for f in g.files if f.startswith("wgrad/")]}
        print("arch grad errors", errs, "weight grad errors", werrs)
        # observed over runs: 1e-3 .. 1.3e-2 (summation order differs from run to run)
        assert max(errs.values()) < 5e-2 and sorted(errs.values())[len(errs) // 2] < 1.5e-2, errs
        assert max(werrs.values()) < 5e-2, werrs
        gi, gf = net.genotype()
        assert repr((gi, [list(gf.pose), list(gf.par)])) == str(g["genotype"][0])
    finally:
        F_.set_compute_dtype(torch.bfloat16)


def test_supernet_bf16_step(lib_built):
    """bf16 product mode: one forward + backward of the supernet runs on the native kernels, stays finite and
    tracks the fp32 fixture at the accuracy bf16 storage allows for this depth; eval mode runs too."""
    from npp_b200 import _lib
    from npp_b200.models.model_search_interact import Network
    g = np.load(os.path.join(HERE, "golden", "search_golden.npz"), allow_pickle=False)
    torch.manual_seed(int(g["seed"]))
    net = Network(_search_cfg(int(g["layers"]), int(g["channels"])))
    arch = [k[5:] for k in g.files if k.startswith("arch/")]
    with torch.no_grad():
        for k in arch:
            getattr(net, k).copy_(torch.from_numpy(g["arch/" + k]))
    sd_cpu = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda().train()
    c0 = _lib.launch_count()
    pl, par = net(torch.from_numpy(g["x"]).cuda())
    outs = [t for pair in pl + par for t in pair]
    sum(t.float().pow(2).mean() for t in outs).backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - c0 > 2000
    names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
    errs = [rel(t, g["out/" + n]) for n, t in zip(names, outs)]
    # yardstick: the oracle restatement of the supernet with every operator output rounded to bf16 and bf16 conv
    # operands (what bf16 storage costs for this depth / this tiny 32-sample-BatchNorm configuration), same fixture
    from oracle import nppnet_ref as O
    O.set_storage_dtype(torch.bfloat16, weights=True)
    try:
        with torch.no_grad():
            ypl, ypar = O.search_forward(sd_cpu, torch.from_numpy(g["x"]).bfloat16().float(), layers=int(g["layers"]),
                                         training=True)
    finally:
        O.set_storage_dtype(None)
    yerrs = [rel(t, g["out/" + n]) for n, t in zip(names, [t for pair in ypl + ypar for t in pair])]
    print("bf16 supernet output errors vs fp32 reference fixture: ours", errs, "yardstick", yerrs)
    assert all(torch.isfinite(t).all() for t in outs)
    for e, ey in zip(errs, yerrs):
        assert e < max(2e-2, 1.5 * ey), (errs, yerrs)
    for p in net.arch_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0
    with torch.no_grad():
        net.eval()
        pl, par = net(torch.from_numpy(g["x"]).cuda())
    assert all(torch.isfinite(t).all() for pair in pl + par for t in pair)
