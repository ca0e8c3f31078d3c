"""GPU: the CUDA path through the C ABI reproduces the golden fixtures generated from the reference itself
(tests/golden/make_golden.py) — operator primitives, the derived network, the fused losses (values and
gradients) and, bit-exactly, the integer evaluation kernels."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = lambda name: np.load(os.path.join(HERE, "golden", name), allow_pickle=False)


def rel(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)], ids=["fp32", "bf16"])
def test_ops_golden(dtype, tol, lib_built):
    from npp_b200 import functional as F_
    from npp_b200.models.operations import OPS
    g = G("ops_golden.npz")
    F_.set_compute_dtype(dtype)
    try:
        for tag in g["cases"]:
            tag = str(tag)
            name, stride = tag.rsplit("_s", 1)
            op = OPS[name](16, int(stride), True)
            sd = {k[len(tag) + 4:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(tag + "/sd/")}
            op.load_state_dict(sd, strict=True)
            op = op.cuda().train()
            x = torch.from_numpy(g[tag + "/x"]).cuda().requires_grad_(True)
            y = F_.from_internal(op(F_.to_internal(x, dtype)), 16)
            (y * torch.from_numpy(g[tag + "/gy"]).cuda()).sum().backward()
            assert rel(y, g[tag + "/y"]) < tol, (tag, rel(y, g[tag + "/y"]))
            gtol = tol
            if dtype == torch.bfloat16:
                # small-sample BatchNorm backward is ill-conditioned in bf16: the bound is what the oracle restatement of
                # the same primitive costs with bf16-rounded storage and bf16 conv operands (1.5x), never below 2e-2
                from oracle import nppnet_ref as O
                xo = torch.from_numpy(g[tag + "/x"]).bfloat16().float().requires_grad_(True)
                sdo = {k: v.clone() for k, v in sd.items()}
                O.set_storage_dtype(torch.bfloat16, weights=True)
                try:
                    yo = O.primitive(name, O.Params(sdo, True), xo, int(stride))
                    (yo * torch.from_numpy(g[tag + "/gy"])).sum().backward()
                finally:
                    O.set_storage_dtype(None)
                gtol = max(tol, 1.5 * rel(xo.grad, g[tag + "/dx"]))
            assert rel(x.grad, g[tag + "/dx"]) < gtol, (tag, "dx", rel(x.grad, g[tag + "/dx"]), gtol)
            for k, b in op.named_buffers():
                if "running" in k:
                    assert rel(b, g["%s/after/%s" % (tag, k)]) < max(tol, 1e-5), (tag, k)
    finally:
        F_.set_compute_dtype(torch.bfloat16)


def test_network_golden_fp32(lib_built):
    from npp_b200 import functional as F_
    from npp_b200.models.model_augment import Network
    g = G("net_golden.npz")
    ns = types.SimpleNamespace
    cfg = ns(DATASET=ns(NUM_CLASSES=20, NUM_JOINTS=16), TRAIN=ns(LAYERS=int(g["layers"]), INIT_CHANNELS=int(g["channels"])),
             MODEL=ns(DECONV_WITH_BIAS=False, HEAD="PSP", REFINE_LAYERS=1))
    F_.set_compute_dtype(torch.float32)
    try:
        torch.manual_seed(int(g["seed"]))
        net = Network(cfg).cuda().train()
        with torch.no_grad():
            pl, par = net(torch.from_numpy(g["x"]).cuda())
        names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
        for n, t in zip(names, [t for pair in pl + par for t in pair]):
            assert rel(t, g["out/" + n]) < 1e-4, (n, rel(t, g["out/" + n]))
    finally:
        F_.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize("tag,min_kept", [("default", 131072), ("kept300", 300)])
def test_loss_golden(tag, min_kept, lib_built):
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    g = G("loss_golden.npz")
    lab, edge = torch.from_numpy(g["lab"]).cuda(), torch.from_numpy(g["edge"]).cuda()
    gt = [torch.from_numpy(g["gt0"]).cuda(), torch.from_numpy(g["gt1"]).cuda()]
    par = [[torch.from_numpy(g["par%d" % i]).cuda().requires_grad_(True),
            torch.from_numpy(g["edgelogit%d" % i]).cuda().requires_grad_(True)] for i in range(2)]
    pose = [[torch.from_numpy(g["pose%d" % i]).cuda().requires_grad_(True),
             torch.from_numpy(g["poseaux%d" % i]).cuda().requires_grad_(True)] for i in range(2)]
    cp = Criterion_par(out_len=2, ignore_index=255, thres=0.9, min_kept=min_kept).cuda()
    cq = Criterion_pose(out_len=2, use_target_weight=False).cuda()
    lp = cp(par, [lab, edge])
    lq = cq(pose, gt)
    (lp + lq).backward()
    assert rel(lp, g[tag + "/loss_par"]) < 1e-5, (lp.item(), g[tag + "/loss_par"])
    assert rel(lq, g[tag + "/loss_pose"]) < 1e-5
    assert rel(cp.lamda.grad, g[tag + "/dlamda_par"]) < 1e-4
    assert rel(cq.lamda.grad, g[tag + "/dlamda_pose"]) < 1e-4
    for i in range(2):
        assert rel(par[i][0].grad, g["%s/dpar%d" % (tag, i)]) < 1e-4, ("dpar", i, rel(par[i][0].grad, g["%s/dpar%d" % (tag, i)]))
        assert rel(par[i][1].grad, g["%s/dedge%d" % (tag, i)]) < 1e-4, ("dedge", i)
        assert rel(pose[i][0].grad, g["%s/dpose%d" % (tag, i)]) < 1e-5
        assert rel(pose[i][1].grad, g["%s/dposeaux%d" % (tag, i)]) < 1e-5


def test_eval_golden_bit_exact(lib_built):
    from npp_b200.core import evaluate as ev
    from npp_b200.utils import calc_pckh, utils
    g = G("eval_golden.npz")
    cm = utils.get_confusion_matrix(torch.from_numpy(g["cm/label"]).cuda(), torch.from_numpy(g["cm/logits"]).cuda(),
                                    (2, 20, 24, 24), 20, 255)
    assert cm.dtype == np.float64 and np.array_equal(cm, g["cm/matrix"])
    acc, avg, cnt, pred = ev.accuracy(g["acc/hm"], g["acc/gt"])
    assert np.array_equal(acc, g["acc/acc"]) and avg == float(g["acc/avg"]) and cnt == int(g["acc/cnt"])
    assert np.array_equal(pred, g["acc/pred"])
    preds, maxvals = ev.get_max_preds(g["acc/hm"])
    assert np.array_equal(preds, g["acc/pred"])
    hit, valid = calc_pckh.pckh_counts(g["pckh/pred"], g["pckh/gt"])
    assert np.array_equal(calc_pckh.pck_from_counts(hit, valid), g["pckh/pck"])
    m = utils.tta_merge(torch.from_numpy(g["tta/pred"]).cuda(), torch.from_numpy(g["tta/flip"]).cuda(), (48, 48))
    assert rel(m, g["tta/merged"]) < 1e-6


def test_eval_large_properties(lib_built):
    """BASELINE config-5 sizes (512x512, 7 classes / 14 joints): size-independent properties + oracle equality."""
    from npp_b200.core import evaluate as ev
    from npp_b200.utils import utils
    from oracle import eval_ref as E
    gen = torch.Generator().manual_seed(5)
    logits = torch.randn(4, 7, 512, 512, generator=gen)
    label = torch.randint(0, 7, (4, 512, 512), generator=gen)
    label[:, :7, :] = 255
    hist = utils.confusion_hist(label.cuda(), logits.cuda(), 7, 255)
    assert int(hist.sum()) == int((label != 255).sum())                 # every valid pixel counted exactly once
    cm = hist.cpu().numpy().reshape(7, 7)
    assert np.array_equal(cm.sum(1), np.bincount(label[label != 255].numpy(), minlength=7))  # row sums = gt histogram
    assert np.array_equal(cm.astype(np.float64), E.confusion_matrix(label.numpy(), logits.numpy(), (4, 7, 512, 512), 7, 255))
    # accumulating two batches == histogram of the concatenation (linearity)
    h2 = utils.confusion_hist(label.cuda(), logits.cuda(), 7, 255, hist=hist.clone())
    assert torch.equal(h2, 2 * hist)
    hm = torch.rand(8, 14, 128, 128, generator=gen)
    gt = torch.rand(8, 14, 128, 128, generator=gen)
    hit, valid, _ = ev.pck_counts(hm.cuda(), gt.cuda())
    ohit, ovalid = E.pck_counts(hm.numpy(), gt.numpy())
    assert np.array_equal(hit.cpu().numpy(), ohit) and np.array_equal(valid.cpu().numpy(), ovalid)
    # identical prediction and target -> every valid joint is a hit
    hit2, valid2, _ = ev.pck_counts(gt.cuda(), gt.cuda())
    assert torch.equal(hit2, valid2)
