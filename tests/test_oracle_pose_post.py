"""Oracle of the LIP pose post-process against fixtures generated with the reference's own
third-party arithmetic (cv2.resize / cv2.flip / scipy gaussian_filter, tests/golden/make_golden_pose.py) and, where the
libraries import, against them live on fresh random inputs.  CPU only."""
import os

import numpy as np
import pytest

from oracle import pose_post_ref as P

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_post_golden.npz"))


def test_resize_matches_cv2_fixture():
    out = P.resize_bilinear_cv2(G["resize_in"], 31, 50)
    assert out.shape == G["resize_out"].shape and out.dtype == np.float32
    np.testing.assert_allclose(out, G["resize_out"], rtol=0, atol=2e-6)


def test_gaussian_filter_matches_scipy_fixture():
    out = P.gaussian_filter_reflect(G["filter_in"], 3.0)
    np.testing.assert_allclose(out, G["filter_out"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_pose_postprocess_matches_reference_loop(tag):
    size = tuple(int(v) for v in G[tag + "_size"])
    pose = P.pose_postprocess(G[tag + "_pred"], G[tag + "_flip"], size, G[tag + "_crop"], G[tag + "_scale"])
    ref = G[tag + "_pose"]
    # arg-max positions (hence x, y) must be identical; the peak value is float32 arithmetic
    np.testing.assert_array_equal(pose[..., :2], ref[..., :2])
    np.testing.assert_allclose(pose[..., 2], ref[..., 2], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(P.lip_csv_rows(pose), P.lip_csv_rows(ref))


def test_first_heatmap_matches():
    hm = P.merged_heatmap(G["a_pred"], G["a_flip"], 0, 0, 96, 96)
    np.testing.assert_allclose(hm, G["a_first_heatmap"], rtol=0, atol=2e-6)


def test_lip_csv_rows_truncate_toward_zero():
    pose = np.zeros((1, 16, 3))
    pose[0, :, 0] = np.linspace(-3.7, 400.9, 16)
    pose[0, :, 1] = np.linspace(250.2, -0.4, 16)
    rows = P.lip_csv_rows(pose)
    assert rows.shape == (1, 32)
    for k, j in enumerate(P.IDX_MAP_TO_LIP):
        assert rows[0, 2 * k] == int(pose[0, j, 0]) and rows[0, 2 * k + 1] == int(pose[0, j, 1])
    assert int(-3.7) == -3 and rows[0, 2 * P.IDX_MAP_TO_LIP.index(0)] == -3


def test_live_against_cv2_and_scipy():
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.RandomState(123)
    for (h, w, oh, ow) in [(24, 24, 96, 96), (12, 20, 48, 80), (7, 5, 30, 11)]:
        img = rng.randn(h, w).astype(np.float32)
        np.testing.assert_allclose(P.resize_bilinear_cv2(img, oh, ow),
                                   cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR), rtol=0, atol=2e-6)
    img = rng.randn(64, 48).astype(np.float32)
    np.testing.assert_allclose(P.gaussian_filter_reflect(img, 3.0), ndi.gaussian_filter(img, sigma=3), rtol=0, atol=2e-6)
