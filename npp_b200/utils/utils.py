"""GPU versions of the evaluation helpers in the reference's utils/utils.py."""
import os

import numpy as np
import torch

from .._lib import call, fptr, i32, stream


def save_checkpoint(states, is_best, output_dir, filename="checkpoint.pth"):
    """utils/utils.py:60-65: the whole `states` dict to checkpoint.pth; when `is_best`, the bare `best_state_dict` to
    model_best.pth.  Keys as the reference's drivers build them (augment_lip_sync.py:268-278): epoch, state_dict
    (DistributedDataParallel-prefixed `module.` names), best_state_dict, perf_iou, perf_pck, lr, optimizer, cri1, cri2."""
    torch.save(states, os.path.join(output_dir, filename))
    if is_best and "state_dict" in states:
        torch.save(states["best_state_dict"], os.path.join(output_dir, "model_best.pth"))


def ddp_state_dict(model):
    """The state_dict a DistributedDataParallel-wrapped model would save: every key prefixed with `module.`
    (augment_lip_sync.py:270 `model.state_dict()` on the DDP wrapper)."""
    return {"module." + k: v for k, v in model.state_dict().items()}


def strip_ddp_prefix(state_dict):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}


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
