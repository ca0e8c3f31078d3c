"""GPU: engine.EvalStep — batched inference + on-GPU evaluation (BASELINE configs[4]; core/function_ppp.py:869-964).
The graph-captured step must produce, bit-exactly, the counters the oracle computes from the same network outputs
(confusion matrix E2, PCK hit / valid counts E3) and the flip merges must match the oracle (E1, pascal heat-map
average); accumulating two batches equals evaluating their union."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(b, size, nc, nj, seed):
    gen = torch.Generator().manual_seed(seed)
    img = torch.randn(b, 3, size, size, generator=gen)
    lab = torch.randint(0, nc, (b, size, size), generator=gen)
    lab[:, :5, :] = 255
    hs = size // 4
    ys = torch.arange(hs).view(1, 1, hs, 1).float()
    xs = torch.arange(hs).view(1, 1, 1, hs).float()
    cy = torch.rand(b, nj, 1, 1, generator=gen) * hs
    cx = torch.rand(b, nj, 1, 1, generator=gen) * hs
    gt = torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * 2.0 * 2.0))
    return img, lab, gt


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_eval_step_counters_bit_exact(use_graph, lib_built):
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.models.model_augment import Network
    from oracle import eval_ref as E
    F_.set_compute_dtype(torch.bfloat16)
    nc, nj, size, b = 7, 14, 128, 3
    torch.manual_seed(0)
    model = Network(engine.make_cfg(num_classes=nc, num_joints=nj, layers=8, init_channels=16)).cuda()
    model.train()
    with torch.no_grad():           # give the running statistics real values
        for s in range(2):
            model(_batch(b, size, nc, nj, 50 + s)[0].cuda())
    model.eval()
    step = engine.EvalStep(model, b, size, use_graph=use_graph)
    batches = [_batch(b, size, nc, nj, 60 + i) for i in range(2)]
    want_cm = np.zeros((nc, nc))
    want_hit, want_valid = np.zeros(nj, dtype=np.int64), np.zeros(nj, dtype=np.int64)
    step.load(*batches[0])
    step.prepare()
    for img, lab, gt in batches:
        step.load(img, lab, gt)
        step.run()
        torch.cuda.synchronize()
        # identical predictions (the step's own last-stage outputs), evaluated by the oracle on the host
        o = {k: t.detach().float().cpu() for k, t in step.outputs.items()}
        merged = E.tta_merge(o["par"], o["flip_par"], (size, size), swap_lr=False)
        assert ((o["merged_par"] - merged).norm() / merged.norm()).item() < 1e-6          # E1 (fp32 rounding only)
        want_cm += E.confusion_matrix(lab.numpy(), o["merged_par"].numpy(), (b, nc, size, size), nc, 255)
        hm = E.flip_average_pascal(o["pose"].numpy(), o["flip_pose"].numpy())
        assert np.array_equal(hm, o["merged_pose"].numpy())
        h, v = E.pck_counts(hm, gt.numpy())
        want_hit += h
        want_valid += v
    cm, miou, pck, hit, valid = step.results()
    # bit-exact given identical predictions (north star): counters of the fused device path == oracle on the host
    assert np.array_equal(cm, want_cm), np.abs(cm - want_cm).sum()
    assert np.array_equal(hit, want_hit) and np.array_equal(valid, want_valid)
    assert int(cm.sum()) == sum(int((lab != 255).sum()) for _, lab, _ in batches)
    assert step.launches_per_step > 300


def test_flip_average_matches_oracle(lib_built):
    from npp_b200.core import evaluate as ev
    from oracle import eval_ref as E
    gen = torch.Generator().manual_seed(2)
    a, f = torch.randn(5, 14, 32, 32, generator=gen), torch.randn(5, 14, 32, 32, generator=gen)
    perm = [0, 1, 8, 9, 10, 11, 12, 13, 2, 3, 4, 5, 6, 7]
    got = ev.flip_average(a.cuda(), f.cuda(), perm).cpu().numpy()
    assert np.array_equal(got, E.flip_average_pascal(a.numpy(), f.numpy(), perm))
