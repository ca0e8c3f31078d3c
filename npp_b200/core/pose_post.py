"""LIP pose post-process of the reference's `validate_sync` on the GPU (core/function.py:962-986).

The reference loops over every image and joint on the host: cv2.resize of the 96x96 heat map and of the mirrored
image's heat map (joint-permuted), cv2.flip, average, scipy gaussian_filter(sigma=3), arg-max, crop/scale inverse —
B x 16 Python iterations with CPU filters per batch.  Here the whole batch is three kernel families (csrc/eval.cu:
npp_pose_merge, npp_gaussian_filter, npp_heatmap_argmax); only the [B, 16] peak indices / values come back to the
host, where the coordinate arithmetic is done in float64 exactly as the reference does it.
"""
import ctypes

import numpy as np
import torch

from .._lib import call, fptr, i32, f64, stream

FLIPPED_POSEIDX = [0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 14, 15]   # core/function.py:908
IDX_MAP_TO_LIP = [10, 9, 8, 11, 12, 13, 15, 14, 1, 0, 4, 3, 2, 5, 6, 7]    # utils/utils.py:279


def _cuda_f32(a):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    return a.detach().to(device="cuda", dtype=torch.float32).contiguous()


def merged_heatmaps(pred_pose, flip_pred_pose, size, flipped_poseidx=FLIPPED_POSEIDX):
    """[B, J, H, W] filtered, flip-averaged heat maps at the network-input resolution (function.py:973-980)."""
    pred, flip = _cuda_f32(pred_pose), _cuda_f32(flip_pred_pose)
    if pred.shape != flip.shape or pred.dim() != 4:
        raise AssertionError("pred_pose / flip_pred_pose must be [B, J, h, w] tensors of the same shape")
    b, j, h, w = pred.shape
    if len(flipped_poseidx) != j:
        raise AssertionError("flipped_poseidx must list one source joint per joint")
    oh, ow = int(size[0]), int(size[1])
    merged = torch.empty((b, j, oh, ow), dtype=torch.float32, device=pred.device)
    fidx = (ctypes.c_int * j)(*[int(v) for v in flipped_poseidx])
    call("npp_pose_merge", fptr(pred), fptr(flip), i32(b), i32(j), i32(h), i32(w), fidx, i32(oh), i32(ow), fptr(merged),
         stream())
    tmp = torch.empty_like(merged)
    call("npp_gaussian_filter", fptr(merged), fptr(tmp), fptr(merged), i32(b * j), i32(oh), i32(ow), f64(3.0), f64(4.0),
         stream())
    return merged


def pose_postprocess(pred_pose, flip_pred_pose, size, crop_param, scale, flipped_poseidx=FLIPPED_POSEIDX):
    """pose [B, J, 3] float64 = (x, y, peak value) in original-image coordinates, as the loop at function.py:969-986
    fills `pose`; crop_param [B, >=1, 4] and scale [B] are the dataset's `meta['crop_param']` / `meta['scale']`."""
    hm = merged_heatmaps(pred_pose, flip_pred_pose, size, flipped_poseidx)
    b, j, oh, ow = hm.shape
    idx = torch.empty((b, j), dtype=torch.int32, device=hm.device)
    mx = torch.empty((b, j), dtype=torch.float32, device=hm.device)
    call("npp_heatmap_argmax", fptr(hm), i32(b * j), i32(oh), i32(ow), fptr(idx), fptr(mx), stream())
    idx_np = idx.cpu().numpy().astype(np.int64)
    mx_np = mx.cpu().numpy()
    cp = np.asarray(crop_param, dtype=np.float64)
    sc = np.asarray(scale, dtype=np.float64).reshape(b, 1)
    px, py = (idx_np % ow).astype(np.float64), (idx_np // ow).astype(np.float64)
    pose = np.zeros((b, j, 3))
    pose[:, :, 0] = (px - cp[:, 0, 2][:, None] + cp[:, 0, 0][:, None]) / sc
    pose[:, :, 1] = (py - cp[:, 0, 3][:, None] + cp[:, 0, 1][:, None]) / sc
    pose[:, :, 2] = mx_np
    return pose


def lip_csv_rows(pose):
    """int(x), int(y) per joint in LIP order — the numbers save_hpe_results_to_lip_format writes (utils.py:278-286)."""
    pose = np.asarray(pose)
    cols = []
    for jj in IDX_MAP_TO_LIP:
        cols.append(np.trunc(pose[:, jj, 0]).astype(np.int64))
        cols.append(np.trunc(pose[:, jj, 1]).astype(np.int64))
    return np.stack(cols, axis=1)
