"""GPU versions of the evaluation helpers in the reference's utils/utils.py."""
import numpy as np
import torch

from .._lib import call, fptr, i32, stream


def confusion_hist(label, pred, num_class, ignore=-1, hist=None):
    """int64 [num_class*num_class] histogram of (gt, argmax pred) pairs accumulated on the device
    (utils/utils.py:192-218 without the host round trip); pass `hist` to accumulate across batches."""
    if isinstance(pred, list):
        pred = pred[0]
    pred = pred.contiguous().float()
    label = label.to(pred.device).long().contiguous()
    n, c, h, w = pred.shape
    if c != num_class:
        raise RuntimeError("prediction has %d channels, num_class=%d" % (c, num_class))
    if hist is None:
        hist = torch.zeros(num_class * num_class, dtype=torch.int64, device=pred.device)
    call("npp_confusion_hist", fptr(pred), fptr(label), i32(n), i32(c), i32(h), i32(w), i32(label.shape[1]),
         i32(label.shape[2]), i32(ignore), fptr(hist), stream())
    return hist


def get_confusion_matrix(label, pred, size, num_class, ignore=-1):
    """utils/utils.py:192-218: float64 [num_class, num_class] matrix indexed [gt, pred].
    `size` crops the label to [:size[-2], :size[-1]] like the reference; pred must already be that size."""
    if isinstance(pred, list):
        pred = pred[0]
    label = label[:, :size[-2], :size[-1]]
    hist = confusion_hist(label, pred, num_class, ignore)
    return hist.cpu().numpy().astype(np.float64).reshape(num_class, num_class)


def tta_merge(pred_par, flip_pred_par, size, swap_lr=True):
    """core/function.py:927-939: resize both predictions to `size` (bilinear, align_corners=False), apply the
    reference's (aliasing) left/right channel copy, un-flip and average."""
    pred_par, flip_pred_par = pred_par.contiguous().float(), flip_pred_par.contiguous().float()
    n, c, h, w = pred_par.shape
    out = torch.empty((n, c, size[-2], size[-1]), dtype=torch.float32, device=pred_par.device)
    call("npp_tta_merge", fptr(pred_par), fptr(flip_pred_par), i32(n), i32(c), i32(h), i32(w), i32(size[-2]),
         i32(size[-1]), i32(1 if swap_lr else 0), fptr(out), stream())
    return out
