"""ORACLE — test infrastructure only.  numpy restatements of the reference's integer evaluation code.

  confusion_matrix      utils/utils.py:192-218   get_confusion_matrix
  tta_merge             core/function.py:927-939 flip-test merge (incl. the aliasing channel copy)
  max_preds / accuracy  core/evaluate.py:13-99   get_max_preds / calc_dists / dist_acc / accuracy
  flip_average_pascal   core/function_ppp.py:905,955-958 heat-map flip average of the pascal validate loop
  pckh                  utils/calc_pckh.py:35-97 get_head_size / get_norm_dist / compute_pck

Pinned against the reference's own functions (imported from /root/reference in the build container,
tests/test_oracle_vs_reference.py) and against tests/golden/eval_*.npz generated from them.
"""
import numpy as np


def confusion_matrix(label, pred, size, num_class, ignore=-1):
    """label int [N,H,W]; pred float [N,C,h,w] (numpy).  Returns float64 [C,C] indexed [gt, pred]."""
    seg_pred = np.argmax(pred.transpose(0, 2, 3, 1), axis=3).astype(np.uint8)
    seg_gt = label[:, :size[-2], :size[-1]].astype(int)
    keep = seg_gt != ignore
    idx = (seg_gt[keep] * num_class + seg_pred[keep]).astype("int32")
    counts = np.bincount(idx)
    cm = np.zeros((num_class, num_class))
    for g in range(num_class):
        for p in range(num_class):
            k = g * num_class + p
            if k < len(counts):
                cm[g, p] = counts[k]
    return cm


def tta_merge(pred, flip_pred, size, swap_lr=True):
    """torch in / torch out: bilinear resize (align_corners=False), aliased channel copy, flip, average."""
    import torch.nn.functional as F
    a = F.interpolate(pred, size=(size[-2], size[-1]), mode="bilinear")
    b = F.interpolate(flip_pred, size=(size[-2], size[-1]), mode="bilinear")
    if swap_lr:
        # the reference assigns through an alias (`tmp = flip_pred_par`), so each pair ends up as two copies of
        # the odd channel: 14,15 <- 15; 16,17 <- 17; 18,19 <- 19
        for lo in (14, 16, 18):
            b[:, lo] = b[:, lo + 1]
    return 0.5 * (a + b.flip(3))


def flip_average_pascal(pred_pose, flip_pred_pose, flipped_poseidx=(0, 1, 8, 9, 10, 11, 12, 13, 2, 3, 4, 5, 6, 7)):
    """core/function_ppp.py:905,955-958: numpy float32 in place, joint by joint; the mirrored image's heat maps are
    joint-permuted but not mirrored back."""
    pred = np.array(pred_pose, dtype=np.float32, copy=True)
    flip = np.asarray(flip_pred_pose, dtype=np.float32)
    for ji in range(len(flipped_poseidx)):
        pred[:, ji, :, :] = 0.5 * (pred[:, ji, :, :].copy() + flip[:, flipped_poseidx[ji], :, :])
    return pred


def max_preds(hm):
    """[B,J,H,W] -> preds [B,J,2] float32 (x, y), maxvals [B,J,1]."""
    b, j, h, w = hm.shape
    flat = hm.reshape(b, j, -1)
    idx = np.argmax(flat, 2)
    mx = np.amax(flat, 2)
    preds = np.zeros((b, j, 2), dtype=np.float32)
    preds[:, :, 0] = idx % w
    preds[:, :, 1] = np.floor(idx / w)
    preds *= (mx > 0.0).astype(np.float32)[:, :, None]
    return preds, mx[:, :, None]


def pck_counts(output, target, thr=0.5):
    """Per-joint hit / valid counts behind accuracy() (evaluate.py:43-65, 68-99)."""
    pred, _ = max_preds(output)
    tgt, _ = max_preds(target)
    b, j = pred.shape[:2]
    h, w = output.shape[2], output.shape[3]
    norm = np.ones((b, 2)) * np.array([h, w]) / 10
    hit = np.zeros(j, dtype=np.int64)
    valid = np.zeros(j, dtype=np.int64)
    for n in range(b):
        for c in range(j):
            if tgt[n, c, 0] < 1 and tgt[n, c, 1] < 1:
                continue
            d = np.linalg.norm(pred[n, c, :] / norm[n] - tgt[n, c, :] / norm[n])
            valid[c] += 1
            hit[c] += d < thr
    return hit, valid


def accuracy(output, target, thr=0.5):
    hit, valid = pck_counts(output, target, thr)
    j = len(hit)
    acc = np.zeros(j + 1)
    tot, cnt = 0, 0
    for i in range(j):
        acc[i + 1] = hit[i] * 1.0 / valid[i] if valid[i] > 0 else 0
        if acc[i + 1] > 0:
            tot += acc[i + 1]
            cnt += 1
    avg = tot / cnt if cnt else 0
    if cnt:
        acc[0] = avg
    return acc, avg, cnt, max_preds(output)[0]


def pckh_counts(pred, gt, thr=0.5):
    """pred, gt float64 [N,P,2]; gt < 0 marks missing joints.  Per-joint hit / valid counts (calc_pckh.py:35-84)."""
    n, p, _ = pred.shape
    head = np.linalg.norm(gt[:, 9, :] - gt[:, 8, :], axis=1)
    for i in range(n):
        if gt[i, 8, 0] < 0 or gt[i, 9, 0] < 0:
            head[i] = 0
    hit = np.zeros(p, dtype=np.int64)
    valid = np.zeros(p, dtype=np.int64)
    for i in range(n):
        if head[i] == 0:
            continue
        d = np.linalg.norm(gt[i] - pred[i], axis=1) / head[i]
        for k in range(p):
            if gt[i, k, 0] < 0 or gt[i, k, 1] < 0:
                continue
            if d[k] >= 0:
                valid[k] += 1
                hit[k] += d[k] <= thr
    return hit, valid


def pck_from_counts(hit, valid):
    p = len(hit)
    out = np.zeros([1, p + 2])
    for k in range(p):
        out[0, k] = 100 * (hit[k] / valid[k])
    ub = list(range(8, 16))
    out[0, p] = 100 * (hit[ub].sum() / valid[ub].sum())
    al = list(range(0, 6)) + list(range(8, 16))
    out[0, p + 1] = 100 * (hit[al].sum() / valid[al].sum())
    return out
