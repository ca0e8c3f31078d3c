"""Generates tests/golden/pose_post_golden.npz: the LIP pose post-process of the reference's validate_sync
(core/function.py:962-986) run with the reference's own third-party arithmetic — cv2.resize(INTER_LINEAR), cv2.flip,
scipy.ndimage.gaussian_filter(sigma=3) — on seeded heat maps.  Build container only:

    python tests/golden/make_golden_pose.py

The loop below is the reference's loop with the tensors replaced by seeded numpy arrays (no model involved); it pins
oracle/pose_post_ref.py (tests/test_oracle_pose_post.py).
"""
import os

import cv2
import numpy as np
from scipy.ndimage import gaussian_filter

HERE = os.path.dirname(os.path.abspath(__file__))
FLIPPED = [0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 14, 15]


def make(seed, n, hs, size):
    rng = np.random.RandomState(seed)
    # smooth-ish heat maps: a few Gaussian blobs + noise, like network outputs in [~0, 1]
    ys, xs = np.mgrid[0:hs, 0:hs].astype(np.float32)
    pred = np.zeros((n, 16, hs, hs), np.float32)
    flip = np.zeros((n, 16, hs, hs), np.float32)
    for arr in (pred, flip):
        for i in range(n):
            for j in range(16):
                cy, cx = rng.uniform(0, hs, 2)
                s = rng.uniform(1.0, 2.5)
                arr[i, j] = np.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * s * s)) + 0.05 * rng.randn(hs, hs)
    crop = np.zeros((n, 2, 4), np.float64)
    crop[:, 0, :] = rng.randint(0, 40, (n, 4))
    scale = rng.uniform(0.6, 1.7, n)
    pose = np.zeros((n, 16, 3))
    first = None
    for num in range(n):
        for ji in range(16):
            heatmap = pred[num, ji].copy()
            heatmap = cv2.resize(heatmap, (size, size), interpolation=cv2.INTER_LINEAR)
            flipped = flip[num, FLIPPED[ji]].copy()
            flipped = cv2.resize(flipped, (size, size), interpolation=cv2.INTER_LINEAR)
            flipped = cv2.flip(flipped, 1)
            heatmap += flipped
            heatmap *= 0.5
            heatmap = gaussian_filter(heatmap, sigma=3)
            if first is None:
                first = heatmap.copy()
            pos = np.unravel_index(heatmap.argmax(), np.shape(heatmap))
            pose[num, ji, 0] = (pos[1] - crop[num, 0, 2] + crop[num, 0, 0]) / scale[num]
            pose[num, ji, 1] = (pos[0] - crop[num, 0, 3] + crop[num, 0, 1]) / scale[num]
            pose[num, ji, 2] = heatmap[pos[0], pos[1]]
    return dict(pred=pred, flip=flip, crop=crop, scale=scale, pose=pose, first_heatmap=first, size=np.array([size, size]))


if __name__ == "__main__":
    out = {}
    for tag, (seed, n, hs, size) in {"a": (1, 2, 24, 96), "b": (2, 1, 32, 128)}.items():
        for k, v in make(seed, n, hs, size).items():
            out[tag + "_" + k] = v
    # one plain resize and one plain filter case (odd sizes, non-integer ratio)
    rng = np.random.RandomState(7)
    img = rng.randn(13, 17).astype(np.float32)
    out["resize_in"] = img
    out["resize_out"] = cv2.resize(img, (50, 31), interpolation=cv2.INTER_LINEAR)
    img2 = rng.randn(40, 29).astype(np.float32)
    out["filter_in"] = img2
    out["filter_out"] = gaussian_filter(img2, sigma=3)
    np.savez_compressed(os.path.join(HERE, "pose_post_golden.npz"), **out)
    print("wrote pose_post_golden.npz:", {k: v.shape for k, v in out.items()})
