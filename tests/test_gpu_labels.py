"""GPU: on-device label synthesis (csrc/labels.cu, npp_b200/dataset/target_generation.py) against the
fixture generated from the reference and against the oracle at the training shapes (B=32, 384^2 labels, 96^2 maps)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _close_fp32(a, b):
    """`a` fp32 from the device, `b` float64 from numpy: equal after rounding b to fp32, up to 1 ulp where CUDA's double
    exp() (not correctly rounded) lands on the other side of an fp32 rounding boundary."""
    b32 = b.astype(np.float32)
    return np.all(np.abs(a - b32) <= np.spacing(np.abs(b32)).astype(np.float32))


def test_labels_match_reference_fixture(lib_built):
    from npp_b200.dataset import target_generation as T
    g = np.load(os.path.join(HERE, "golden", "labels_golden.npz"))
    maps, aux = T.gen_pose_target(g["joints"], g["vis"], 4, 96, 96, 7, aux=True)
    assert _close_fp32(maps.cpu().numpy(), g["pose"]) and _close_fp32(aux.cpu().numpy(), g["pose_aux"])
    assert (maps.cpu().numpy() == g["pose"].astype(np.float32)).mean() > 0.9999
    edge = T.generate_edge(torch.from_numpy(g["label"].astype(np.int64)))
    assert np.array_equal(edge.cpu().numpy(), g["edge"].astype(np.int64))
    flip = T.flip_parsing(torch.from_numpy(g["label"].astype(np.int64)))
    assert np.array_equal(flip.cpu().numpy(), g["flip"].astype(np.int64))
    assert np.array_equal(T.flip_joints(g["joints"], 384), g["flip_joints"])


def test_labels_at_training_shapes_vs_oracle(lib_built):
    from npp_b200 import engine
    from npp_b200.dataset import target_generation as T
    from oracle import labels_ref as R
    b = 32
    rng = np.random.RandomState(3)
    joints = rng.uniform(0, 384, size=(b, 16, 2))
    vis = (rng.uniform(size=(b, 16)) > 0.15).astype(np.int32)
    _, par, _, _, _ = engine.synthetic_batch(b, 384, seed=9)
    maps, aux = T.gen_pose_target(joints, vis, 4, 96, 96, 7, aux=True)
    edge = T.generate_edge(par)
    flip = T.flip_parsing(par)
    torch.cuda.synchronize()
    for i in (0, 13, 31):
        mo, ao = R.gen_pose_target(joints[i], vis[i], 4, 96, 96, 7, aux=True)
        assert _close_fp32(maps[i].cpu().numpy(), mo) and _close_fp32(aux[i].cpu().numpy(), ao)
        lab = par[i].numpy().astype(np.uint8)
        assert np.array_equal(edge[i].cpu().numpy(), R.generate_edge(lab).astype(np.int64))
        assert np.array_equal(flip[i].cpu().numpy(), R.flip_parsing(lab).astype(np.int64))
    # properties at full size: flipping twice is the identity; background = 1 - max over joints; edges only where
    # labels differ nearby and never on ignored pixels
    assert torch.equal(T.flip_parsing(flip), par.cuda())
    assert torch.allclose(maps[:, 16], 1 - maps[:, :16].max(1).values, atol=1e-6)
    assert int(((edge == 1) & (par.cuda() == 255)).sum()) == 0 and int((edge == 255).sum()) == int((par == 255).sum())
    # the synthesized labels feed the training step's criteria directly
    from npp_b200.core.criterion import Criterion_par, Criterion_pose
    # (blocky labels: with per-pixel random classes every pixel is an edge, the edge class weights become (1, 0) and
    #  the weighted cross-entropy is 0 / 0 in the reference too)
    blocks = torch.randint(0, 20, (4, 48, 48), generator=torch.Generator().manual_seed(1))
    blocks = blocks.repeat_interleave(8, 1).repeat_interleave(8, 2)
    blocks[:, :8, :] = 255
    bedge = T.generate_edge(blocks)
    assert int((bedge == 1).sum()) > 0 and int((bedge == 0).sum()) > 0
    logits = [[torch.randn(4, 20, 96, 96).cuda(), torch.randn(4, 2, 96, 96).cuda()]]
    lp = Criterion_par(out_len=1).cuda()(logits, [blocks.cuda(), bedge])
    lq = Criterion_pose(out_len=1).cuda()([[maps[:4, :16].contiguous(), aux[:4, :16].contiguous()]],
                                          [maps[:4, :16].contiguous(), aux[:4, :16].contiguous()])
    assert torch.isfinite(lp) and abs(float(lq) + 2.5) < 1e-6        # identical prediction and target: loss = lamda
