"""GPU parity of the LIP pose post-process (npp_pose_merge / npp_gaussian_filter /
npp_heatmap_argmax, npp_b200/core/pose_post.py) against the oracle (oracle/pose_post_ref.py, itself pinned to
cv2 / scipy by tests/test_oracle_pose_post.py) on the committed fixture inputs and on fresh seeded maps.

First B200 run (last seconds of round 1's GPU budget): all four cases green — profiles/r01_pytest_gpu_n1_and_bilinear_sep.log.
"""
import os

import numpy as np
import pytest

from oracle import pose_post_ref as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_post_golden.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_pose_postprocess_matches_fixture(tag, lib_built):
    from npp_b200.core import pose_post
    size = tuple(int(v) for v in G[tag + "_size"])
    pose = pose_post.pose_postprocess(G[tag + "_pred"], G[tag + "_flip"], size, G[tag + "_crop"], G[tag + "_scale"])
    ref = G[tag + "_pose"]
    np.testing.assert_array_equal(pose[..., :2], ref[..., :2])           # integer peak positions -> identical x, y
    np.testing.assert_allclose(pose[..., 2], ref[..., 2], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(pose_post.lip_csv_rows(pose), P.lip_csv_rows(ref))


@pytest.mark.gpu
def test_merged_heatmaps_match_oracle(lib_built):
    from npp_b200.core import pose_post
    rng = np.random.RandomState(5)
    pred = rng.rand(2, 16, 24, 24).astype(np.float32)
    flip = rng.rand(2, 16, 24, 24).astype(np.float32)
    hm = pose_post.merged_heatmaps(pred, flip, (96, 96)).cpu().numpy()
    for num, ji in [(0, 0), (0, 5), (1, 11), (1, 15)]:
        np.testing.assert_allclose(hm[num, ji], P.merged_heatmap(pred, flip, num, ji, 96, 96), rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_gaussian_filter_ragged_plane(lib_built):
    import torch
    from npp_b200._lib import call, fptr, i32, f64, stream
    x = torch.from_numpy(G["filter_in"]).cuda().contiguous()
    tmp, out = torch.empty_like(x), torch.empty_like(x)
    call("npp_gaussian_filter", fptr(x), fptr(tmp), fptr(out), i32(1), i32(x.shape[0]), i32(x.shape[1]), f64(3.0), f64(4.0),
         stream())
    np.testing.assert_allclose(out.cpu().numpy(), G["filter_out"], rtol=0, atol=2e-6)
