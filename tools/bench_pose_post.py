"""Timing of the LIP pose post-process: the reference's host loop (cv2.resize / cv2.flip / scipy
gaussian_filter per image and joint, core/function.py:962-986) on the host cores next to the device path
(npp_b200/core/pose_post.py) when a GPU is present.

    python tools/bench_pose_post.py [--batch 32] [--size 384]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=384)
ap.add_argument("--cpu-images", type=int, default=4, help="images timed through the host loop (scaled to the batch)")
args = ap.parse_args()

rng = np.random.RandomState(0)
hs = args.size // 4
pred = rng.rand(args.batch, 16, hs, hs).astype(np.float32)
flip = rng.rand(args.batch, 16, hs, hs).astype(np.float32)
crop = np.zeros((args.batch, 2, 4))
scale = np.ones(args.batch)
out = {"batch": args.batch, "size": args.size}

try:   # the reference's own arithmetic
    import cv2
    from scipy.ndimage import gaussian_filter
    FL = [0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 14, 15]
    n = min(args.cpu_images, args.batch)
    t0 = time.perf_counter()
    for num in range(n):
        for ji in range(16):
            hm = cv2.resize(pred[num, ji].copy(), (args.size, args.size), interpolation=cv2.INTER_LINEAR)
            fh = cv2.flip(cv2.resize(flip[num, FL[ji]].copy(), (args.size, args.size), interpolation=cv2.INTER_LINEAR), 1)
            hm = gaussian_filter((hm + fh) * 0.5, sigma=3)
            np.unravel_index(hm.argmax(), hm.shape)
    dt = (time.perf_counter() - t0) / n
    out["host_loop_ms_per_image"] = 1e3 * dt
    out["host_loop_ms_per_batch"] = 1e3 * dt * args.batch
except ImportError as e:
    out["host_loop"] = "unavailable: %r" % (e,)

try:
    import torch
    if torch.cuda.is_available():
        from npp_b200.core import pose_post
        p, f = torch.from_numpy(pred).cuda(), torch.from_numpy(flip).cuda()
        for _ in range(2):
            pose_post.pose_postprocess(p, f, (args.size, args.size), crop, scale)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            pose_post.pose_postprocess(p, f, (args.size, args.size), crop, scale)
        torch.cuda.synchronize()
        out["device_ms_per_batch"] = 1e3 * (time.perf_counter() - t0) / reps
except Exception as e:  # no GPU here: the host figure stands alone
    out["device"] = "unavailable: %r" % (e,)
print(json.dumps(out))
