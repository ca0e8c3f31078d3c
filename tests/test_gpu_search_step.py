"""GPU: engine.SearchStep — the bilevel search step of the reference (core/function.py:485-621
`train_with_alpha`, optimizers of search_lip_sync.py:273-279): a weight step on batch 1, then an architecture step on
batch 2 with loss2 = 2 * mean(par + pose [+ 2 * loss_entropy()]), both inside one CUDA graph."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(use_graph, bilevel=True, entropy=False, seed=0):
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    from npp_b200.models.model_search_interact import Network
    F_.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(seed)
    model = Network(engine.make_cfg(layers=8, init_channels=16)).cuda().train()
    cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2, min_kept=2000).cuda()
    w_opt, a_opt = engine.build_search_optimizers(model, cpose, cpar)
    step = engine.SearchStep(model, cpose, cpar, w_opt, a_opt, 2, 128, use_graph=use_graph, warmup=1, bilevel=bilevel,
                             entropy=entropy)
    return model, step


def test_optimizer_split_matches_reference_setup(lib_built):
    model, step = _make(False)
    arch = set(id(p) for p in model.arch_parameters())
    w_ids = set(id(p) for g in step.opt.param_groups for p in g["params"])
    a_ids = set(id(p) for g in step.a_opt.param_groups for p in g["params"])
    assert a_ids == arch and not (w_ids & arch)
    assert len(arch) == 12
    g = step.a_opt.param_groups[0]
    assert g["lr"] == 0.001 and tuple(g["betas"]) == (0.5, 0.999) and g["weight_decay"] == 0.001
    assert [pg["lr"] for pg in step.opt.param_groups[1:]] == [0.0001, 0.0001]


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_bilevel_step_moves_weights_then_alphas(use_graph, lib_built):
    from npp_b200 import engine
    model, step = _make(use_graph, entropy=True)
    a0 = [p.detach().clone() for p in model.arch_parameters()]
    w0 = model.stem0[0].weight.detach().clone()
    step.load(*engine.synthetic_batch(2, 128, seed=3))
    step.load2(*engine.synthetic_batch(2, 128, seed=4))
    step.prepare()                      # warm-up / capture leave the state untouched
    assert all(torch.equal(a, p) for a, p in zip(a0, model.arch_parameters()))
    assert torch.equal(w0, model.stem0[0].weight)
    step.run()
    torch.cuda.synchronize()
    l1, l2 = float(step.loss), float(step.loss2)
    assert l1 == l1 and l2 == l2 and l2 > l1            # loss2 = 2 * (...) + entropy term on similar data
    assert not torch.equal(w0, model.stem0[0].weight)
    moved = [not torch.equal(a, p) for a, p in zip(a0, model.arch_parameters())]
    assert all(moved), moved
    # first Adam step: |delta| ~ lr (1e-3) for every architecture entry with a gradient
    d = (model.alphas_pose.detach() - a0[4]).abs()
    assert 5e-4 < d.max().item() < 1.5e-3, d.max().item()
    assert step.launches_per_step > 1500


def test_weight_only_schedule_leaves_alphas(lib_built):
    """epochs < 15 (search_lip_sync.py:325-326): `train` — architecture gradients are produced but never applied."""
    from npp_b200 import engine
    model, step = _make(True, bilevel=False)
    a0 = [p.detach().clone() for p in model.arch_parameters()]
    step.load(*engine.synthetic_batch(2, 128, seed=3))
    step.prepare()
    for _ in range(2):
        step.run()
    torch.cuda.synchronize()
    assert all(torch.equal(a, p) for a, p in zip(a0, model.arch_parameters()))
    assert step.a_flat.abs().sum().item() > 0           # the backward did compute d loss / d alpha


def test_bilevel_schedule_order_and_losses(lib_built):
    """The step is exactly `train_with_alpha` (core/function.py:510-528, 555-621): [forward, backward, Adam on the
    weights] on batch 1, then [forward, backward, Adam on the architecture] on batch 2 — checked on the ABI call trace —
    and the two reported losses are the reference's loss1 = mean(par + pose) on batch 1 with the initial weights and
    loss2 = 2 * mean(par + pose) on batch 2 with the UPDATED weights (forward-only re-evaluation, 5e-2: bf16)."""
    from npp_b200 import _lib, engine
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    model, step = _make(False)
    ref0 = copy.deepcopy(model)
    b1, b2 = engine.synthetic_batch(2, 128, seed=3), engine.synthetic_batch(2, 128, seed=4)
    step.load(*b1)
    step.load2(*b2)
    step.prepare()
    lam_pose, lam_par = step.cpose.lamda.detach().clone(), step.cpar.lamda.detach().clone()
    _lib.trace_begin()
    step.run()
    trace = _lib.trace_end()
    torch.cuda.synchronize()
    names = [t[0] for t in trace]
    adam = [i for i, n in enumerate(names) if n == "npp_adam_step"]
    assert len(adam) == 2
    n_w = sum(p.numel() > 0 for g in step.opt.param_groups for p in g["params"])
    n_a = sum(1 for g in step.a_opt.param_groups for p in g["params"])
    assert trace[adam[0]][1][1].value == n_w and trace[adam[1]][1][1].value == n_a == 12   # weights first, then alphas
    fwd = [i for i, n in enumerate(names) if n == "npp_nchw_to_nhwc"]
    assert fwd[0] < adam[0] < min(i for i in fwd if i > adam[0]) < adam[1]                # second forward after the w-step
    half = names[:adam[0]].count("npp_conv2d_fwd")
    assert half > 300 and names[adam[0]:adam[1]].count("npp_conv2d_fwd") == half           # two identical passes

    def loss_of(net, batch, lp, lq):
        cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2, min_kept=2000).cuda()
        with torch.no_grad():
            cpose.lamda.copy_(lp), cpar.lamda.copy_(lq)
            img, par, edge, g0, g1 = [t.cuda() for t in batch]
            pose, parl = net(img)
            return float((cpar(parl, [par, edge]).unsqueeze(0) + cpose(pose, [g0, g1]).unsqueeze(0)).mean())

    l1 = loss_of(ref0.train(), b1, lam_pose, lam_par)
    assert abs(float(step.loss) - l1) < 5e-2 * abs(l1), (float(step.loss), l1)
    # batch 2 on the weights AFTER the weight step (the alphas have moved by 1e-3 since: second-order effect)
    l2 = 2 * loss_of(model, b2, step.cpose.lamda.detach(), step.cpar.lamda.detach())
    assert abs(float(step.loss2) - l2) < 5e-2 * abs(l2), (float(step.loss2), l2)
