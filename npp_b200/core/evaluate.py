"""Heat-map PCK on the GPU — drop-in for the reference's core/evaluate.py (:13-99).

`get_max_preds` and `accuracy` take numpy arrays or CUDA tensors shaped [B, J, H, W] and return the same
structures as the reference (numpy).  The arg-max and the hit / valid counting run in csrc/eval.cu; the
counters are int64 and bit-exact, the ratios are formed from them in float64 exactly as numpy does.
"""
import numpy as np
import torch

from .._lib import call, fptr, i32, f32, stream


def _as_cuda_f32(a):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if a.dim() != 4:
        raise AssertionError("batch_images should be 4-ndim")
    return a.to(device="cuda", dtype=torch.float32).contiguous()


def heatmap_argmax(hm):
    """first arg-max (flat index, int32) and max value per [B, J] map."""
    hm = _as_cuda_f32(hm)
    b, j, h, w = hm.shape
    idx = torch.empty((b, j), dtype=torch.int32, device=hm.device)
    mx = torch.empty((b, j), dtype=torch.float32, device=hm.device)
    call("npp_heatmap_argmax", fptr(hm), i32(b * j), i32(h), i32(w), fptr(idx), fptr(mx), stream())
    return idx, mx


def get_max_preds(batch_heatmaps):
    """evaluate.py:13-41: preds [B,J,2] float32 (x = idx % W, y = floor(idx / W), zeroed where max <= 0), maxvals [B,J,1]."""
    hm = _as_cuda_f32(batch_heatmaps)
    w = hm.shape[3]
    idx, mx = heatmap_argmax(hm)
    idx_np, mx_np = idx.cpu().numpy().astype(np.int64), mx.cpu().numpy()
    preds = np.stack([idx_np % w, idx_np // w], axis=2).astype(np.float32)
    preds *= (mx_np > 0.0).astype(np.float32)[:, :, None]
    return preds, mx_np[:, :, None]


def flip_average(pred_pose, flip_pred_pose, flipped_poseidx):
    """core/function_ppp.py:957-958: 0.5 * (pred[:, j] + flip_pred[:, flipped_poseidx[j]]) per joint, in heat-map
    space (the reference does not mirror the second set of maps back; neither do we)."""
    import ctypes
    pred, flip = _as_cuda_f32(pred_pose), _as_cuda_f32(flip_pred_pose)
    if pred.shape != flip.shape:
        raise AssertionError("pred_pose / flip_pred_pose must have the same shape")
    b, j, h, w = pred.shape
    if len(flipped_poseidx) != j:
        raise AssertionError("flipped_poseidx must list one source joint per joint")
    out = torch.empty_like(pred)
    fidx = (ctypes.c_int * j)(*[int(v) for v in flipped_poseidx])
    call("npp_heatmap_flip_avg", fptr(pred), fptr(flip), i32(b), i32(j), i32(h), i32(w), fidx, fptr(out), stream())
    return out


def pck_counts(output, target, thr=0.5, hit=None, valid=None):
    """Per-joint (hit, valid) int64 counters of evaluate.py:43-65 for hm_type='gaussian'; pass `hit` / `valid` to
    accumulate across batches on the device."""
    out, tgt = _as_cuda_f32(output), _as_cuda_f32(target)
    b, j, h, w = out.shape
    pi, pm = heatmap_argmax(out)
    gi, gm = heatmap_argmax(tgt)
    if hit is None:
        hit = torch.zeros(j, dtype=torch.int64, device=out.device)
    if valid is None:
        valid = torch.zeros(j, dtype=torch.int64, device=out.device)
    call("npp_pck_counts", fptr(pi), fptr(pm), fptr(gi), fptr(gm), i32(b), i32(j), i32(h), i32(w), f32(thr), fptr(hit),
         fptr(valid), stream())
    return hit, valid, (pi, pm)


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    """evaluate.py:68-99: returns (acc[J+1], avg_acc, cnt, pred) — acc[0] is the mean over joints with acc > 0."""
    if hm_type != "gaussian":
        raise NotImplementedError("only hm_type='gaussian' is used by the reference's loops")
    hit, valid, (pi, pm) = pck_counts(output, target, thr)
    hit, valid = hit.cpu().numpy(), valid.cpu().numpy()
    j = hit.shape[0]
    acc = np.zeros(j + 1)
    avg_acc, cnt = 0, 0
    for i in range(j):
        acc[i + 1] = hit[i] * 1.0 / valid[i] if valid[i] > 0 else 0
        if acc[i + 1] > 0:
            avg_acc = avg_acc + acc[i + 1]
            cnt += 1
    avg_acc = avg_acc / cnt if cnt != 0 else 0
    if cnt != 0:
        acc[0] = avg_acc
    w = output.shape[3]
    idx_np, mx_np = pi.cpu().numpy().astype(np.int64), pm.cpu().numpy()
    pred = np.stack([idx_np % w, idx_np // w], axis=2).astype(np.float32)
    pred *= (mx_np > 0.0).astype(np.float32)[:, :, None]
    return acc, avg_acc, cnt, pred
