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


def test_alpha_update_direction_matches_manual_sequence(lib_built):
    """The graph-captured alpha step takes the same first Adam step as a hand-written eager sequence with
    torch.optim.Adam on a copy of the model (sign agreement: the first Adam update is lr * sign(g) up to eps)."""
    from npp_b200 import engine
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    model, step = _make(True)
    ref = copy.deepcopy(model)
    b1, b2 = engine.synthetic_batch(2, 128, seed=3), engine.synthetic_batch(2, 128, seed=4)
    step.load(*b1)
    step.load2(*b2)
    step.prepare()
    step.run()
    torch.cuda.synchronize()
    cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2, min_kept=2000).cuda()
    arch = set(id(p) for p in ref.arch_parameters())
    w_opt = torch.optim.Adam([p for p in ref.parameters() if id(p) not in arch], 0.0015)
    a_opt = torch.optim.Adam(ref.arch_parameters(), lr=0.001, betas=(0.5, 0.999), weight_decay=0.001)

    def loss_of(batch):
        img, par, edge, g0, g1 = [t.cuda() for t in batch]
        pose, parl = ref(img)
        return (cpar(parl, [par, edge]).unsqueeze(0) + cpose(pose, [g0, g1]).unsqueeze(0)).mean()

    w_opt.zero_grad()
    loss_of(b1).backward()
    w_opt.step()
    a_opt.zero_grad()
    (2 * loss_of(b2)).backward()
    a_opt.step()
    agree = tot = 0
    for p, q, init in zip(model.arch_parameters(), ref.arch_parameters(), [1e-3] * 12):
        dp, dq = (p.detach() - init).flatten(), (q.detach() - init).flatten()
        big = dq.abs() > 2e-4
        agree += int((torch.sign(dp[big]) == torch.sign(dq[big])).sum())
        tot += int(big.sum())
    assert tot > 100 and agree / tot > 0.9, (agree, tot)
