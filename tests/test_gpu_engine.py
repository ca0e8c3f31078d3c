"""GPU: engine.TrainStep (the caller side of the hot path, core/function.py:87-107) — the CUDA-graph replay of a
whole training step must produce the same losses and parameters as the eager launches of the same step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(seed, layers=8, channels=16, batch=2, size=128, use_graph=False):
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    from npp_b200.models.model_augment import Network
    F_.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(seed)
    model = Network(engine.make_cfg(layers=layers, init_channels=channels)).cuda().train()
    cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2, min_kept=2000).cuda()
    opt = engine.build_optimizer(model, cpose, cpar)
    step = engine.TrainStep(model, cpose, cpar, opt, batch, size, use_graph=use_graph, warmup=1)
    return model, step


def test_graph_replay_matches_eager(lib_built):
    """Same parameters, same batch: a graph replay computes the same loss as the eager launches.  The learning rate
    is 0 so both models stay identical.  (Gradients are not compared element-wise: in this tiny random-init
    configuration two EAGER runs of the same step already differ by ~50% in gradient norm — fp32 atomics reorder
    sums, bf16 rounding flips ReLU / max-pool / OHEM decisions, and BatchNorm over 32 samples amplifies it; the
    gradient parity of the kernels is established against the oracle in test_gpu_network.py / test_gpu_ops.py.)"""
    from npp_b200 import engine
    batches = [engine.synthetic_batch(2, 128, seed=10 + i) for i in range(3)]
    ref_model, eager = _make(0, use_graph=False)
    model, graphed = _make(0, use_graph=True)
    for st in (eager, graphed):
        for g in st.opt.param_groups:
            g["lr"] = 0.0
    graphed.load(*batches[0])
    graphed.prepare()   # one eager warm-up step (optimizer state, kernel attributes), then the capture
    assert graphed.launches_per_step > 500
    for b in batches[1:]:
        eager.load(*b)
        graphed.load(*b)
        le, lg = float(eager.run()), float(graphed.run())
        torch.cuda.synchronize()
        assert abs(le - lg) <= 5e-3 * abs(le), (le, lg)
        ge, gg = eager.flat_grads, graphed.flat_grads
        assert torch.isfinite(gg).all() and gg.abs().max() > 0
        assert 0.5 < (gg.norm() / ge.norm()).item() < 2.0
    for p, q in zip(ref_model.parameters(), model.parameters()):
        assert torch.equal(p, q)


def test_train_step_loss_decreases(lib_built):
    from npp_b200 import engine
    model, step = _make(2, use_graph=True)
    step.load(*engine.synthetic_batch(2, 128, seed=5))
    step.prepare()
    losses = [float(step.run()) for _ in range(12)]
    assert all(l == l for l in losses)
    assert min(losses[-3:]) < losses[0], losses


def test_fused_adam_in_train_step_updates_every_used_parameter(lib_built):
    from npp_b200 import engine
    model, step = _make(1, use_graph=False)
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    step.load(*engine.synthetic_batch(2, 128, seed=3))
    step.run()
    torch.cuda.synchronize()
    # TrainStep uses persistent flat gradients: a parameter that takes no part in the forward keeps an all-zero
    # gradient and is not moved.  SE_Block.bn exists but is unused at stride 1 (operations.py:117,126-129): that is
    # the reference's DDP find_unused_parameters set.
    unused = [k for k, p in model.named_parameters() if not p.grad.any()]
    assert unused and all(".bn." in k for k in unused), unused
    for k, p in model.named_parameters():
        if k in unused:
            assert torch.equal(before[k], p.detach()), k
    used = [k for k, p in model.named_parameters() if k not in unused]
    moved = sum(int(not torch.equal(before[k], p.detach())) for k, p in model.named_parameters() if k in used)
    assert moved >= 0.95 * len(used), (moved, len(used))
