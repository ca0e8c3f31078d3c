"""GPU parity of the LIP pose post-process (npp_pose_merge / npp_gaussian_filter /
npp_heatmap_argmax, npp_b200/core/pose_post.py) against the oracle (oracle/pose_post_ref.py, itself pinned to
cv2 / scipy by tests/test_oracle_pose_post.py) on the committed fixture inputs and on fresh seeded maps.

These kernels were written after the round's GPU budget was spent: they compile for sm_100a and the oracle is
pinned on the CPU, but they have not run on hardware yet — hence the non-strict xfail (a pass shows up as XPASS).
Remove the marker once a B200 run is green.
"""
import os

import numpy as np
import pytest

from oracle import pose_post_ref as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_post_golden.npz"))
PENDING = pytest.mark.xfail(reason="N1 kernels not yet run on a B200 (written after the GPU budget of round 1)",
                            strict=False)


@pytest.mark.gpu
@PENDING
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
@PENDING
def test_merged_heatmaps_match_oracle(lib_built):
    from npp_b200.core import pose_post
    rng = np.random.RandomState(5)
    pred = rng.rand(2, 16, 24, 24).astype(np.float32)
    flip = rng.rand(2, 16, 24, 24).astype(np.float32)
    hm = pose_post.merged_heatmaps(pred, flip, (96, 96)).cpu().numpy()
    for num, ji in [(0, 0), (0, 5), (1, 11), (1, 15)]:
        np.testing.assert_allclose(hm[num, ji], P.merged_heatmap(pred, flip, num, ji, 96, 96), rtol=0, atol=2e-6)


@pytest.mark.gpu
@PENDING
def test_gaussian_filter_ragged_plane(lib_built):
    import torch
    from npp_b200._lib import call, fptr, i32, f64, stream
    x = torch.from_numpy(G["filter_in"]).cuda().contiguous()
    tmp, out = torch.empty_like(x), torch.empty_like(x)
    call("npp_gaussian_filter", fptr(x), fptr(tmp), fptr(out), i32(1), i32(x.shape[0]), i32(x.shape[1]), f64(3.0), f64(4.0),
         stream())
    np.testing.assert_allclose(out.cpu().numpy(), G["filter_out"], rtol=0, atol=2e-6)
