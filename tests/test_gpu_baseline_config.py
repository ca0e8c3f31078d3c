"""GPU parity at the BASELINE configuration (BASELINE.json configs[1]): derived NPPNet L=16 / C=64 at 384x384.

The small fixtures (L=8, C=16, 128^2, B=2..4) leave BatchNorm with 32 samples at the coarsest maps, which is why
their bf16 bounds are yardstick-relative.  Here the network is the one the metric is quoted on, BatchNorm sees
>= 1152 samples everywhere, and the product path (bf16, tcgen05 kernels through the C ABI) is compared with the
oracle run ON THE SAME GPU in fp64 (test side only; stock torch ops) on identical seeded weights and inputs:

  * STAGE-WISE (the bound that bf16 storage can attain, asserted at 2e-2): every one of the 56 stages of the network
    (32 encoder cells, 6 decoder cells, 4 layer convs, 6 refinement cells, 8 heads) is fed the oracle's own stage
    inputs (rounded to bf16) and its outputs, input gradients and parameter gradients are compared with the fp64
    oracle of that stage on the same inputs — forward within 2e-2, gradients within max(2e-2, 1.5 x yardstick);
  * END-TO-END: measured on a B200 (profiles/r02_parity_trace_baseline_config.txt): at random init this network
    amplifies ANY perturbation by ~1.3x per cell — the reference's own arithmetic with bf16-rounded storage is 0.5 %
    off after the stems, 2 % after 4 cells, 45 % after 16 cells and 26-75 % at the logits, and its gradients are
    decorrelated (relative error > 1).  No bf16 implementation can meet 2e-2 at the logits of this random-init
    network; what can be asserted end to end is that the product path adds nothing to that: at EVERY traced stage
    its error is within 1.25x of the yardstick (oracle with bf16-rounded activations AND bf16 conv operands, i.e. what
    torch autocast would compute), and fp32 validation mode meets 1e-4 at the logits (test below);
  * the criteria (OHEM select over 4.7 M pixels with min_kept = 131072 < n, edge CE, heat-map MSE) at
    [32, 20, 96, 96] -> 384^2 against the oracle in fp64.

Every number is also written to gpurun_out/parity_baseline_config.json when that directory exists.
Reference: models/model_augment.py:402-574, core/criterion.py:54-72,158-217.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_err(a, b):
    a, b = a.detach().double(), b.detach().double().to(a.device)
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _record(key, value):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    p = os.path.join(d, "parity_baseline_config.json")
    try:
        cur = json.load(open(p))
    except Exception:
        cur = {}
    cur[key] = value
    json.dump(cur, open(p, "w"), indent=1)


NAMES = ["pose0", "poseaux0", "pose1", "poseaux1", "par0", "edge0", "par1", "edge1"]
L, C, B, S = 16, 64, 8, 384


def _median(v):
    v = sorted(v)
    return v[len(v) // 2]


def _setup():
    from npp_b200 import engine
    from npp_b200.models.model_augment import Network
    torch.manual_seed(0)
    net = Network(engine.make_cfg(layers=L, init_channels=C))
    net_sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    return net, net_sd, x


def _oracle_forward(net_sd, x, dt, storage=None, weights=False, trace=False):
    from oracle import nppnet_ref as O
    sd = {k: (v.cuda().to(dt) if v.is_floating_point() else v.cuda()) for k, v in net_sd.items()}
    store = {} if trace else None
    O.set_trace(store)
    O.set_storage_dtype(storage, weights=weights)
    try:
        with torch.no_grad():
            pl, par = O.network_forward(sd, x.cuda().to(dt), layers=L, training=True)
    finally:
        O.set_trace(None)
        O.set_storage_dtype(None)
    outs = {n: t for n, t in zip(NAMES, [t for pair in pl + par for t in pair])}
    if store is not None:
        store.update(outs)
        return {k: v.float() for k, v in store.items()}
    return outs


def test_baseline_config_end_to_end_bf16_vs_yardstick(lib_built):
    """Whole network, bf16 product mode: per-stage error against the fp64 oracle next to the yardstick's."""
    from npp_b200 import functional as F_
    F_.set_compute_dtype(torch.bfloat16)
    net, net_sd, x = _setup()
    mine = {}
    net = net.cuda().train()
    net._trace = lambda n, t: mine.__setitem__(n, t.float())
    with torch.no_grad():
        pl, par = net(x.cuda())
    for n, t in zip(NAMES, [t for pair in pl + par for t in pair]):
        mine[n] = t.float()
    del net, pl, par
    torch.cuda.empty_cache()
    truth = _oracle_forward(net_sd, x, torch.float64, trace=True)
    yard = _oracle_forward(net_sd, x, torch.float32, torch.bfloat16, weights=True, trace=True)
    rows, worst = [], 0.0
    for k, t in truth.items():
        if k not in mine:
            continue
        e, ey = rel_err(mine[k], t), rel_err(yard[k], t)
        rows.append((k, e, ey))
        worst = max(worst, e / ey)
    print("stage                          ours   yardstick(bf16 storage + bf16 conv operands)")
    for k, e, ey in rows:
        print("%-28s %8.5f %8.5f" % (k, e, ey))
    _record("end_to_end_bf16", {"rows": rows, "worst_ratio": worst})
    amp = [r for r in rows if r[0].startswith("relu(cells1.")]
    print("yardstick amplification per encoder cell: %.3f" % ((amp[-1][2] / amp[0][2]) ** (1.0 / (len(amp) - 1))))
    assert len(rows) >= 70
    for k, e, ey in rows:
        assert e < 1.25 * ey + 1e-3, (k, e, ey)
    # where bf16 storage CAN attain the north-star bound end to end (before the amplification takes over) it is met
    early = [r for r in rows if r[0] in ("stem2", "stem5", "relu(cells1.0)", "relu(cells2.0)", "relu(cells1.1)", "relu(cells2.1)")]
    assert early and all(e < 2e-2 for _, e, _ in early), early


def test_baseline_config_end_to_end_fp32(lib_built):
    """fp32 validation mode at the BASELINE network (B=2): logits / heat maps within 1e-4 of the fp64 oracle."""
    from npp_b200 import functional as F_
    F_.set_compute_dtype(torch.float32)
    try:
        net, net_sd, x = _setup()
        x = x[:2]
        net = net.cuda().train()
        with torch.no_grad():
            pl, par = net(x.cuda())
        outs = [t.float() for pair in pl + par for t in pair]
        del net
        torch.cuda.empty_cache()
        truth = _oracle_forward(net_sd, x, torch.float64)
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False            # the yardstick is true fp32 arithmetic (stock torch, cuDNN)
        try:
            yard = _oracle_forward(net_sd, x, torch.float32)
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        errs = {n: (rel_err(a, truth[n]), rel_err(yard[n], truth[n])) for n, a in zip(NAMES, outs)}
        print("fp32 validation mode @L16/C64/384^2 forward vs fp64 oracle (ours, fp32 oracle):",
              {k: "%.2e/%.2e" % v for k, v in errs.items()})
        _record("end_to_end_fp32", errs)
        # stage-0 outputs (after ~60 layers) meet the north-star 1e-4; the refinement outputs sit ~40 layers deeper
        # and fp32 rounding itself is amplified past 1e-4 there: bounded by the reference's own fp32 arithmetic
        for n in ("pose0", "poseaux0", "par0", "edge0"):
            assert errs[n][0] < 1e-4, errs
        for n, (e, ey) in errs.items():
            assert e < max(1e-4, 1.5 * ey), errs
    finally:
        F_.set_compute_dtype(torch.bfloat16)


def _stage_list(net):
    """(name, our callable(*internal inputs) -> tensor(s), oracle callable(P, *inputs) -> tensor(s), input keys,
    parameter prefix)."""
    from npp_b200 import functional as F_
    from oracle import nppnet_ref as O
    reduces = [L // 4, 2 * L // 4, 3 * L // 4]
    st = []
    for s_ in (1, 2):
        for i in range(L):
            red, red_prev = i in reduces, (i - 1) in reduces
            cell = getattr(net, "cells%d" % s_)[i]
            st.append(("cells%d.%d" % (s_, i),
                       lambda a, b, cell=cell: cell(a, b, out_raw=True, out_relu=False),
                       lambda P, a, b, s_=s_, i=i, red=red, rp=red_prev: O.encoder_cell(P.sub("cells%d" % s_).sub(i), a, b, red, rp),
                       ["cells%d.%d.in0" % (s_, i), "cells%d.%d.in1" % (s_, i)], "cells%d.%d." % (s_, i)))
        for d in range(3):
            cell = getattr(net, "upsamples%d" % s_)[d]
            edges = O.DECODER_UP1 if s_ == 1 else O.DECODER_UP2
            st.append(("upsamples%d.%d" % (s_, d), lambda a, b, cell=cell: cell(a, b),
                       lambda P, a, b, s_=s_, d=d, edges=edges: O.upsample_cell(P.sub("upsamples%d" % s_).sub(d), a, b, edges),
                       ["upsamples%d.%d.in0" % (s_, d), "upsamples%d.%d.in1" % (s_, d)], "upsamples%d.%d." % (s_, d)))
    for name, src in (("pose_auxlayer", "x1"), ("edge_layer", "x2"), ("pose_layer", "x1"), ("par_layer", "x2")):
        mod = getattr(net, name)
        st.append((name, lambda a, mod=mod: mod(a), lambda P, a, name=name: O._seq_conv_bn(P.sub(name), a, 1, 2, relu_in=True),
                   [src], name + "."))
    for k in range(3):
        for name, edges in (("pose_net", O.FUSION_POSE), ("par_net", O.FUSION_PAR)):
            cell = getattr(net, name)[k]
            st.append(("%s.%d" % (name, k), lambda a, b, c, cell=cell: cell(a, b, c, out_raw=True, out_relu=False),
                       lambda P, a, b, c, name=name, k=k, edges=edges: O.fusion_cell(P.sub(name).sub(k), a, b, c, edges),
                       ["%s.%d.in%d" % (name, k, q) for q in range(3)], "%s.%d." % (name, k)))
    for i in range(2):
        for q, (name, ksz) in enumerate((("pose_auxnet", 3), ("edge_head", 3), ("pose_head", 1), ("par_head", 1))):
            mod = getattr(net, name)[i]
            st.append(("%s.%d" % (name, i), lambda a, mod=mod: mod(a),
                       lambda P, a, name=name, i=i, ksz=ksz: O.head(P.sub(name).sub(i), a, ksz),
                       ["head_in.%d.%d" % (i, q)], "%s.%d." % (name, i)))
    return st


def test_baseline_config_stagewise_bf16(lib_built):
    """Every stage of the BASELINE network on the oracle's own stage inputs: forward, input gradients and parameter
    gradients of the bf16 product path against the fp64 oracle of that stage."""
    from npp_b200 import functional as F_
    from oracle import nppnet_ref as O
    F_.set_compute_dtype(torch.bfloat16)
    net, net_sd, x = _setup()
    feats = _oracle_forward(net_sd, x, torch.float64, trace=True)        # realistic activations for every stage input
    feats = {k: v.bfloat16().float() for k, v in feats.items() if ".in" in k or k.startswith("head_in") or k in ("x1", "x2")}
    net = net.cuda().train()
    params = dict(net.named_parameters())
    gen = torch.Generator().manual_seed(7)
    table, bad = [], []

    def oracle_stage(fn, keys, prefix, dt, storage, cots):
        sd = {k: (v.cuda().to(dt) if v.is_floating_point() else v.cuda()) for k, v in net_sd.items() if k.startswith(prefix)}
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
        ins = [feats[k].detach().cuda().to(dt).clone().requires_grad_(True) for k in keys]
        O.set_storage_dtype(storage, weights=storage is not None)
        try:
            out = fn(O.Params(sd, True), *ins)
        finally:
            O.set_storage_dtype(None)
        outs = list(out) if isinstance(out, (tuple, list)) else [out]
        sum((o * c.to(dt)).sum() for o, c in zip(outs, cots)).backward()
        return [o.detach() for o in outs], [t.grad for t in ins], {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}

    for name, mine_fn, ora_fn, keys, prefix in _stage_list(net):
        ins = [F_.to_internal(feats[k].cuda()).detach().requires_grad_(True) for k in keys]
        out = mine_fn(*ins)
        outs = list(out) if isinstance(out, (tuple, list)) else [out]
        true_c = {"pose_auxnet": 16, "edge_head": 2, "pose_head": 16, "par_head": 20}.get(name.split(".")[0])
        outs = [F_.from_internal(F_.to_internal(o), true_c) for o in outs]      # heads: drop the channel padding
        cots = [torch.randn(o.shape, generator=gen).cuda() for o in outs]
        for p in params.values():
            p.grad = None
        sum((o * c).sum() for o, c in zip(outs, cots)).backward()
        torch.cuda.synchronize()
        o64, i64, g64 = oracle_stage(ora_fn, keys, prefix, torch.float64, None, cots)
        oy, iy, gy = oracle_stage(ora_fn, keys, prefix, torch.float32, torch.bfloat16, cots)
        fwd = max(rel_err(a, b) for a, b in zip(outs, o64))
        fwd_y = max(rel_err(a, b) for a, b in zip(oy, o64))
        din = max(rel_err(F_.from_internal(t.grad, r.shape[1]), r) for t, r in zip(ins, i64))
        assert all(r is not None for r in i64 + iy), name
        din_y = max(rel_err(a, b) for a, b in zip(iy, i64))
        gmax = max(v.abs().max().item() for v in g64.values())
        ge, gey = [], []
        for k, ref in g64.items():
            if ref.abs().max().item() < 1e-4 * gmax or params[k].grad is None:
                continue
            ge.append(rel_err(params[k].grad, ref))
            gey.append(rel_err(gy[k], ref))
        p90 = lambda v: sorted(v)[int(0.9 * (len(v) - 1))]
        row = (name, fwd, fwd_y, din, din_y, _median(ge), _median(gey), max(ge), max(gey))
        table.append(row)
        # forward: the north-star bound.  Gradients: a training-mode BatchNorm backward subtracts two nearly equal
        # terms, so bf16 storage costs ~7 % on d_in and ~3 % on the typical weight gradient for the reference's own
        # arithmetic too (yardstick columns); a few tensors per reduce cell have almost no gradient signal and are
        # 30-60 % off in the yardstick as well — hence median / p90 against the yardstick, and only a loose cap on the worst
        if not (fwd < 2e-2 and din < max(2e-2, 1.5 * din_y) and _median(ge) < max(2e-2, 1.5 * _median(gey))
                and p90(ge) < max(5e-2, 1.5 * p90(gey)) and max(ge) < max(0.1, 3.0 * max(gey))):
            bad.append(row)
        del out, outs, ins, o64, i64, g64, oy, iy, gy
    print("%-16s %8s %8s | %8s %8s | %8s %8s | %8s %8s" % ("stage", "fwd", "yard", "d_in", "yard", "dW med", "yard", "dW max", "yard"))
    for r in table:
        print("%-16s %8.4f %8.4f | %8.4f %8.4f | %8.4f %8.4f | %8.4f %8.4f" % r)
    _record("stagewise_bf16", {"columns": ["stage", "fwd", "fwd_yard", "din", "din_yard", "dw_median", "dw_median_yard",
                                           "dw_max", "dw_max_yard"], "rows": table})
    assert len(table) == 56
    assert not bad, bad


def _loss_inputs(batch, boost, seed=3):
    from npp_b200 import engine
    gen = torch.Generator().manual_seed(seed)
    _, par, edge, g0, g1 = engine.synthetic_batch(batch, 384, seed=seed)
    low = par[:, 2::4, 2::4].clone()
    low[low == 255] = 0
    preds = []
    for i in range(2):
        par_logit = torch.randn(batch, 20, 96, 96, generator=gen) * 2.0
        par_logit.scatter_add_(1, low.unsqueeze(1), torch.full((batch, 1, 96, 96), float(boost)))
        edge_logit = torch.randn(batch, 2, 96, 96, generator=gen)
        preds.append([par_logit, edge_logit])
    pose = [[torch.rand(batch, 16, 96, 96, generator=gen), torch.rand(batch, 16, 96, 96, generator=gen)] for _ in range(2)]
    return preds, pose, par, edge, [g0, g1]


@pytest.mark.parametrize("boost,gtol", [(0.0, 1e-4), (9.0, 2e-3)], ids=["thr0.9", "thr_min_kept"])
def test_baseline_config_criteria(boost, gtol, lib_built):
    """Criterion_par / Criterion_pose at the bench size.  boost = 0: few confident pixels, threshold = 0.9 (criterion.py:66);
    boost = 9: most pixels confident, threshold = the (min_kept)-th smallest target probability, i.e. exactly the
    element sort() would pick out of 4.7 M (the radix select).  In the second case a pixel whose probability equals
    the threshold to the last bit can fall on either side when the two implementations round the bilinear taps
    differently, hence the wider gradient bound there (membership of a handful of the 131072 kept pixels)."""
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    from oracle import nppnet_ref as O
    batch = 32
    preds, pose, par, edge, gt = _loss_inputs(batch, boost)
    dev = torch.device("cuda")
    mp = [[t.to(dev).requires_grad_(True) for t in pr] for pr in preds]
    mq = [[t.to(dev).requires_grad_(True) for t in pr] for pr in pose]
    cp, cq = Criterion_par(out_len=2).to(dev), Criterion_pose(out_len=2).to(dev)
    lp = cp(mp, [par.to(dev), edge.to(dev)])
    lq = cq(mq, [g.to(dev) for g in gt])
    (lp + lq).backward()
    torch.cuda.synchronize()

    dt = torch.float64
    op = [[t.to(dev).to(dt).requires_grad_(True) for t in pr] for pr in preds]
    oq = [[t.to(dev).to(dt).requires_grad_(True) for t in pr] for pr in pose]
    lam_p = (2.3 * torch.ones(2, dtype=dt, device=dev)).requires_grad_(True)
    lam_q = (-2.5 * torch.ones(2, dtype=dt, device=dev)).requires_grad_(True)
    w = torch.tensor(O.WEIGHTS_LIP, dtype=dt, device=dev)
    # stage by stage so the fp64 384^2 temporaries of one stage are freed before the next
    olp = O.criterion_par(op, [par.to(dev), edge.to(dev)], lam_p, w)
    olq = O.criterion_pose(oq, [g.to(dev).to(dt) for g in gt], lam_q)
    (olp + olq).backward()
    res = {"loss_par": rel_err(lp, olp), "loss_pose": rel_err(lq, olq),
           "dlamda_par": rel_err(cp.lamda.grad, lam_p.grad), "dlamda_pose": rel_err(cq.lamda.grad, lam_q.grad)}
    for i in range(2):
        res["dpar%d" % i] = rel_err(mp[i][0].grad, op[i][0].grad)
        res["dedge%d" % i] = rel_err(mp[i][1].grad, op[i][1].grad)
        res["dpose%d" % i] = rel_err(mq[i][0].grad, oq[i][0].grad)
        res["dposeaux%d" % i] = rel_err(mq[i][1].grad, oq[i][1].grad)
    print("criteria @[32,20,96,96]->384^2, boost %.0f:" % boost, {k: "%.2e" % v for k, v in res.items()})
    _record("criteria_boost%d" % int(boost), res)
    assert res["loss_par"] < 1e-5 and res["loss_pose"] < 1e-5, res
    assert res["dlamda_par"] < 1e-4 and res["dlamda_pose"] < 1e-4, res
    for i in range(2):
        assert res["dpar%d" % i] < gtol, res
        assert res["dedge%d" % i] < 1e-4, res
        assert res["dpose%d" % i] < 1e-5 and res["dposeaux%d" % i] < 1e-5, res
