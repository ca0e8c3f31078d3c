"""CPU: the oracle restatement reproduces the golden fixtures generated from the reference itself
(tests/golden/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import os
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = lambda name: np.load(os.path.join(HERE, "golden", name), allow_pickle=False)


def _close(a, b, tol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= tol * (np.linalg.norm(b) + 1e-12)


def test_ops_golden():
    from oracle import nppnet_ref as O
    g = G("ops_golden.npz")
    for tag in g["cases"]:
        tag = str(tag)
        name, stride = tag.rsplit("_s", 1)
        sd = {k[len(tag) + 4:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(tag + "/sd/")}
        for k, v in sd.items():
            if v.is_floating_point() and v.dim() > 0 and "running" not in k:
                v.requires_grad_(True)
        x = torch.from_numpy(g[tag + "/x"]).requires_grad_(True)
        y = O.primitive(name, O.Params(sd, True), x, int(stride))
        (y * torch.from_numpy(g[tag + "/gy"])).sum().backward()
        assert _close(y.detach().numpy(), g[tag + "/y"]), tag
        assert _close(x.grad.numpy(), g[tag + "/dx"], 1e-4), tag
        for k in g.files:
            if k.startswith(tag + "/grad/"):
                assert _close(sd[k[len(tag) + 6:]].grad.numpy(), g[k], 1e-3) or np.abs(g[k]).max() < 1e-5, k
            if k.startswith(tag + "/after/"):
                assert _close(sd[k[len(tag) + 7:]].numpy(), g[k]), k


def _cfg(layers, channels):
    ns = types.SimpleNamespace
    return ns(DATASET=ns(NUM_CLASSES=20, NUM_JOINTS=16), TRAIN=ns(LAYERS=layers, INIT_CHANNELS=channels),
              MODEL=ns(DECONV_WITH_BIAS=False, HEAD="PSP", REFINE_LAYERS=1))


def test_network_golden():
    """Same seed -> same init as the reference (checksum), and the oracle forward reproduces its outputs."""
    from npp_b200.models.model_augment import Network
    from oracle import nppnet_ref as O
    g = G("net_golden.npz")
    torch.manual_seed(int(g["seed"]))
    net = Network(_cfg(int(g["layers"]), int(g["channels"])))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    chk = float(sum(v.double().sum() for v in sd.values() if v.is_floating_point()))
    assert abs(chk - float(g["init_checksum"][0])) < 1e-6 * abs(chk), "seeded init differs from the reference's"
    with torch.no_grad():
        pl, par = O.network_forward(sd, torch.from_numpy(g["x"]), layers=int(g["layers"]), training=True)
    names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
    for n, t in zip(names, [t for pair in pl + par for t in pair]):
        assert _close(t.numpy(), g["out/" + n], 1e-4), n


def test_loss_golden():
    from oracle import nppnet_ref as O
    g = G("loss_golden.npz")
    lab, edge = torch.from_numpy(g["lab"]), torch.from_numpy(g["edge"])
    gt = [torch.from_numpy(g["gt0"]), torch.from_numpy(g["gt1"])]
    w = torch.tensor(O.WEIGHTS_LIP)
    for tag, min_kept in (("default", 131072), ("kept300", 300)):
        par = [[torch.from_numpy(g["par%d" % i]).requires_grad_(True),
                torch.from_numpy(g["edgelogit%d" % i]).requires_grad_(True)] for i in range(2)]
        pose = [[torch.from_numpy(g["pose%d" % i]).requires_grad_(True),
                 torch.from_numpy(g["poseaux%d" % i]).requires_grad_(True)] for i in range(2)]
        lam_p = (2.3 * torch.ones(2)).requires_grad_(True)
        lam_q = (-2.5 * torch.ones(2)).requires_grad_(True)
        lp = O.criterion_par(par, [lab, edge], lam_p, w, min_kept=min_kept)
        lq = O.criterion_pose(pose, gt, lam_q)
        (lp + lq).backward()
        assert _close(lp.detach().numpy(), g[tag + "/loss_par"], 1e-6)
        assert _close(lq.detach().numpy(), g[tag + "/loss_pose"], 1e-6)
        assert _close(lam_p.grad.numpy(), g[tag + "/dlamda_par"], 1e-5)
        assert _close(lam_q.grad.numpy(), g[tag + "/dlamda_pose"], 1e-5)
        for i in range(2):
            assert _close(par[i][0].grad.numpy(), g["%s/dpar%d" % (tag, i)], 1e-5)
            assert _close(par[i][1].grad.numpy(), g["%s/dedge%d" % (tag, i)], 1e-5)
            assert _close(pose[i][0].grad.numpy(), g["%s/dpose%d" % (tag, i)], 1e-5)
            assert _close(pose[i][1].grad.numpy(), g["%s/dposeaux%d" % (tag, i)], 1e-5)


def test_eval_golden():
    from oracle import eval_ref as E
    g = G("eval_golden.npz")
    cm = E.confusion_matrix(g["cm/label"], g["cm/logits"], (2, 20, 24, 24), 20, 255)
    assert np.array_equal(cm, g["cm/matrix"])
    assert cm.sum() == (g["cm/label"][:, :24, :24] != 255).sum()      # checksum: every valid pixel counted once
    acc, avg, cnt, pred = E.accuracy(g["acc/hm"], g["acc/gt"])
    assert np.array_equal(acc, g["acc/acc"]) and avg == float(g["acc/avg"]) and cnt == int(g["acc/cnt"])
    assert np.array_equal(pred, g["acc/pred"])
    hit, valid = E.pckh_counts(g["pckh/pred"], g["pckh/gt"])
    assert np.array_equal(E.pck_from_counts(hit, valid), g["pckh/pck"])
    m = E.tta_merge(torch.from_numpy(g["tta/pred"]), torch.from_numpy(g["tta/flip"]), (48, 48))
    assert np.array_equal(m.numpy(), g["tta/merged"])


def test_pckh_known_answer():
    """SURVEY.md §4 self-consistency KAT: predictions 3 px off every annotated joint score 100 everywhere
    (head sizes in the fixture rows are far larger than 6 px)."""
    from oracle import eval_ref as E
    g = G("eval_golden.npz")
    gt = g["pckh/gt"]
    pred = gt + 3.0
    hit, valid = E.pckh_counts(pred, gt)
    assert (hit == valid).all() and valid.sum() > 0


def test_search_supernet_golden():
    """Oracle restatement of model_search_interact.Network.forward + loss_entropy vs the fixture generated from the
    reference: outputs, and the architecture gradients of sum_i <out_i, r_i> + 3*loss_entropy."""
    from npp_b200.models.model_search_interact import Network
    from oracle import nppnet_ref as O
    g = G("search_golden.npz")
    ns = types.SimpleNamespace
    L, C = int(g["layers"]), int(g["channels"])
    cfg = ns(DATASET=ns(NUM_CLASSES=20, NUM_JOINTS=16), SEARCH=ns(LAYERS=L, INIT_CHANNELS=C),
             MODEL=ns(DECONV_WITH_BIAS=False, HEAD="PSP", REFINE_LAYERS=1))
    torch.manual_seed(int(g["seed"]))
    net = Network(cfg)     # parameters only: same seeded init as the reference (tests/test_cpu_api.py)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    arch = [k[5:] for k in g.files if k.startswith("arch/")]
    assert len(arch) == 12
    for k in arch:
        sd[k] = torch.from_numpy(g["arch/" + k].copy()).requires_grad_(True)
    pl, par = O.search_forward(sd, torch.from_numpy(g["x"]), layers=L, training=True)
    names = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
    gr = torch.Generator().manual_seed(321)
    loss = 0
    for n, t in zip(names, [t for pair in pl + par for t in pair]):
        assert _close(t.detach().numpy(), g["out/" + n], 1e-5), n
        loss = loss + (t * torch.randn(t.shape, generator=gr)).sum()
    ent = O.loss_entropy([sd[k] for k in ("alphas1", "alphas2", "alphas3", "alphas4", "alphas_pose", "alphas_par")])
    assert abs(float(ent) - float(g["entropy"][0])) < 1e-6
    (loss + 3.0 * ent).backward()
    for k in arch:
        assert _close(sd[k].grad.numpy(), g["grad/" + k], 1e-3), k
