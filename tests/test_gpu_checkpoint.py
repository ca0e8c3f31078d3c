"""GPU: checkpoints in the reference's format around the graph executor (utils/utils.py:60-65,
models/model_augment.py:673-709, augment_lip_sync.py:222-237,268-278): train with the CUDA-graph step, save a
DDP-style (`module.`-prefixed) checkpoint incl. optimizer and criteria, resume into a fresh model + optimizer and get the
same next step; warm-start another head configuration through load_pretrain_backbone (shape mismatches skipped); and
the saved weights evaluated by the oracle give the product path's eval outputs."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(seed, num_classes=20, num_joints=16, use_graph=True):
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    from npp_b200.models.model_augment import Network
    F_.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(seed)
    model = Network(engine.make_cfg(num_classes=num_classes, num_joints=num_joints, layers=8, init_channels=16)).cuda().train()
    cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2, min_kept=2000).cuda()
    opt = engine.build_optimizer(model, cpose, cpar)
    step = engine.TrainStep(model, cpose, cpar, opt, 2, 128, use_graph=use_graph, warmup=1)
    return model, cpose, cpar, opt, step


def test_checkpoint_roundtrip_resume_and_oracle_eval(tmp_path, lib_built):
    from npp_b200 import engine
    from npp_b200 import functional as F_
    from npp_b200.utils import utils as U
    from oracle import nppnet_ref as O
    model, cpose, cpar, opt, step = _build(0)
    batches = [engine.synthetic_batch(2, 128, seed=40 + i) for i in range(4)]
    step.load(*batches[0])
    step.prepare()
    for b in batches[:3]:
        step.load(*b)
        step.run()
    torch.cuda.synchronize()
    U.save_checkpoint({"epoch": 3, "state_dict": U.ddp_state_dict(model), "best_state_dict": U.ddp_state_dict(model),
                       "perf_iou": 0.1, "perf_pck": 0.2, "lr": 0.0015, "optimizer": opt.state_dict(),
                       "cri1": cpose.state_dict(), "cri2": cpar.state_dict()}, True, str(tmp_path))
    assert os.path.isfile(tmp_path / "checkpoint.pth") and os.path.isfile(tmp_path / "model_best.pth")
    ck = torch.load(tmp_path / "checkpoint.pth", map_location="cpu")
    assert all(k.startswith("module.") for k in ck["state_dict"]) and len(ck["state_dict"]) == len(model.state_dict())

    # ---- resume (augment_lip_sync.py:226-237): strip the prefix, strict load, optimizer + criteria state
    model2, cpose2, cpar2, opt2, step2 = _build(123)          # different init: everything must come from the file
    model2.load_state_dict(U.strip_ddp_prefix(ck["state_dict"]), strict=True)
    cpose2.load_state_dict(ck["cri1"])
    cpar2.load_state_dict(ck["cri2"])
    opt2.load_state_dict(ck["optimizer"])
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    step2.load(*batches[3])
    step2.prepare()                                            # warm-up + capture must not disturb the resumed state
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    assert max(int(st["step"]) for st in opt2.state.values() if "step" in st) == 3
    step.load(*batches[3])
    l1, l2 = float(step.run()), float(step2.run())
    torch.cuda.synchronize()
    assert abs(l1 - l2) < 2e-2 * abs(l1), (l1, l2)             # same step from the same state (bf16, atomics order)
    assert max(int(st["step"]) for st in opt2.state.values() if "step" in st) == 4
    moved = [(a - b).abs().max().item() for a, b in zip(model.parameters(), model2.parameters())]
    assert max(moved) < 0.0015 * 2.5                           # both took one Adam step of <= lr from the same point

    # ---- the saved weights in the oracle: eval-mode outputs of the product path (fp32 validation mode) match
    best = U.strip_ddp_prefix(torch.load(tmp_path / "model_best.pth", map_location="cpu"))
    x = torch.randn(2, 3, 128, 128, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        opose, opar = O.network_forward({k: v.clone() for k, v in best.items()}, x, layers=8, training=False)
    F_.set_compute_dtype(torch.float32)
    try:
        from npp_b200.models.model_augment import Network
        m3 = Network(engine.make_cfg(layers=8, init_channels=16))
        m3.load_state_dict(best, strict=True)
        m3 = m3.cuda().eval()
        with torch.no_grad():
            pose, par = m3(x.cuda())
        for a, b in zip([t for p in pose + par for t in p], [t for p in opose + opar for t in p]):
            err = ((a.cpu().double() - b.double()).norm() / b.double().norm()).item()
            assert err < 1e-4, err
    finally:
        F_.set_compute_dtype(torch.bfloat16)


def test_load_pretrain_backbone_skips_mismatched_heads(tmp_path, lib_built, capsys):
    """model_augment.py:673-709: warm start of a 7-class / 14-joint model (pascal) from a 20 / 16 LIP checkpoint with
    DDP-prefixed keys — backbone tensors are taken, the four last-layer head tensors with other shapes are skipped."""
    from npp_b200.utils import utils as U
    model, *_ = _build(1)
    path = str(tmp_path / "encoder.pth")
    torch.save(U.ddp_state_dict(model), path)
    target, *_ = _build(2, num_classes=7, num_joints=14)
    before = {k: v.detach().clone() for k, v in target.state_dict().items()}
    target.load_pretrain_backbone(path)
    out = capsys.readouterr().out
    src = model.state_dict()
    skipped = taken = 0
    for k, v in target.state_dict().items():
        if src[k].shape != v.shape:
            skipped += 1
            assert torch.equal(v, before[k]), k                # kept its own initialisation
            assert ("Skip loading parameter %s," % k) in out
        else:
            taken += 1
            assert torch.equal(v, src[k]), k
    assert skipped == 12 and taken > 1500, (skipped, taken)    # 2 stages x (pose, pose_aux, par) heads x (weight, bias)
    assert "successful load pretrained backbone" in out
    target.load_pretrain_backbone(str(tmp_path / "missing.pth"))   # a missing file is silently ignored, as upstream
