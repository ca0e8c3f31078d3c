"""Losses of NPPNet on libnpp_b200 kernels — drop-in for the reference's core/criterion.py.

`Criterion_pose`, `Criterion_par` and `OhemCrossEntropy` keep the reference's constructor signatures,
parameter names (`lamda`) and call conventions (criterion.py:43-217): predictions are the NCHW fp32
tensors `Network.forward` returns, labels are int64 maps / fp32 heat maps.  The forward and backward
arithmetic runs in the fused kernels of csrc/loss.cu; the uncertainty weighting
`loss * exp(-lamda) + lamda` is scalar torch arithmetic on 0-dim tensors.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from .._lib import call, fptr, i32, i64, f32, stream

# criterion.py:13-21
pascal = [0.82877791, 0.95688253, 0.94921949, 1.00538108, 1.0201687, 1.01665831, 1.05470914]
weights_pascal = torch.from_numpy(np.array(pascal)).float()
lip = [0.7602572, 0.94236198, 0.85644457, 1.04346266, 1.10627293, 0.80980162,
       0.95168713, 0.8403769, 1.05798412, 0.85746254, 1.01274366, 1.05854692,
       1.03430773, 0.84867818, 0.88027721, 0.87580925, 0.98747462, 0.9876475,
       1.00016535, 1.00108882]
weights_lip = torch.from_numpy(np.array(lip)).float()


def _prep_logits(x):
    if x.dim() != 4 or not x.is_cuda:
        raise RuntimeError("expected CUDA NCHW logits, got %s on %s" % (tuple(x.shape), x.device))
    return x.contiguous().float()


def _prep_target(t):
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


def _ce_bwd_sep_enabled():
    from .. import functional as F_
    return F_._state.get("ce_bwd_sep", False)


def _grad_through_upsample(score, lh, lw, align_corners, fill):
    """d loss / d head-logit from the per-pixel gradient at label resolution: `fill(G, cq)` writes G [N, lh, lw, cq]
    (fp32 NHWC), the separable bilinear backward pulls it back to the head resolution, the result is converted to
    the NCHW layout of `score` (round-2 candidate path, see csrc/loss.cu ce_grad_kernel)."""
    from .._lib import NPP_F32, f64, ref, view
    n, c, h, w = score.shape
    cq = (c + 3) // 4 * 4
    dev = score.device
    G = torch.empty((n, lh, lw, cq), dtype=torch.float32, device=dev)
    fill(G, cq)
    dxn = torch.empty((n, h, w, cq), dtype=torch.float32, device=dev)
    tmp = torch.empty(n * lh * w * cq, dtype=torch.float32, device=dev)
    gv, dv = G.permute(0, 3, 1, 2), dxn.permute(0, 3, 1, 2)      # logical NCHW over NHWC memory
    call("npp_bilinear_bwd_sep", ref(view(gv)), ref(view(dv)), fptr(tmp), i32(align_corners), f64(0.0), f64(0.0),
         i32(NPP_F32), stream())
    d = torch.empty_like(score)
    call("npp_nhwc_to_nchw", ref(view(dv)), fptr(d), i32(c), i32(NPP_F32), stream())
    return d


class _OhemCEFn(Function):
    """OhemCrossEntropy.forward (criterion.py:54-72) of bilinear_upsample(score) — fused."""

    @staticmethod
    def forward(ctx, score, target, weight, ignore_index, thresh, min_kept, align_corners):
        n, c, h, w = score.shape
        lh, lw = target.shape[1], target.shape[2]
        if weight.numel() != c:
            raise RuntimeError("weight tensor should be defined either for all %d classes or no classes but got "
                               "weight tensor of shape: [%d]" % (c, weight.numel()))
        dev = score.device
        npix = n * lh * lw
        prob = torch.empty(npix, dtype=torch.float32, device=dev)
        loss = torch.empty(npix, dtype=torch.float32, device=dev)
        n_valid = torch.zeros(1, dtype=torch.int64, device=dev)
        call("npp_par_loss_pixels", fptr(score), i32(n), i32(c), i32(h), i32(w), fptr(target), i32(lh), i32(lw),
             fptr(weight), i32(ignore_index), i32(align_corners), fptr(prob), fptr(loss), fptr(n_valid), stream())
        out3 = torch.zeros(3, dtype=torch.float32, device=dev)
        ws = torch.zeros(1056, dtype=torch.int32, device=dev)
        call("npp_ohem_select", fptr(prob), fptr(loss), i64(npix), fptr(n_valid), i32(min_kept), f32(thresh),
             fptr(out3), fptr(ws), stream())
        ctx.save_for_backward(score, target, weight, prob, out3)
        ctx.cfg = (ignore_index, align_corners)
        return out3[0] / out3[1]

    @staticmethod
    def backward(ctx, g):
        score, target, weight, prob, out3 = ctx.saved_tensors
        ignore_index, align_corners = ctx.cfg
        n, c, h, w = score.shape
        gs = g.reshape(1).float().contiguous()
        if _ce_bwd_sep_enabled():
            lh, lw = target.shape[1], target.shape[2]
            d = _grad_through_upsample(score, lh, lw, align_corners, lambda G, cq: call(
                "npp_par_loss_grad_pixels", fptr(score), i32(n), i32(c), i32(h), i32(w), fptr(target), i32(lh), i32(lw),
                fptr(weight), i32(ignore_index), i32(align_corners), fptr(prob), fptr(out3), fptr(gs), fptr(G), i32(cq),
                stream()))
            return d, None, None, None, None, None, None
        d = torch.zeros_like(score)
        call("npp_par_loss_bwd", fptr(score), i32(n), i32(c), i32(h), i32(w), fptr(target), i32(target.shape[1]),
             i32(target.shape[2]), fptr(weight), i32(ignore_index), i32(align_corners), fptr(prob), fptr(out3),
             fptr(gs), fptr(d), stream())
        return d, None, None, None, None, None, None


class _EdgeCEFn(Function):
    """F.cross_entropy(bilinear_upsample(edge), target, [w_neg, w_pos], ignore_index) (criterion.py:161-166,194-197)."""

    @staticmethod
    def forward(ctx, score, target, posneg, ignore_index, align_corners):
        n, c, h, w = score.shape
        if c != 2:
            raise RuntimeError("edge logits must have 2 channels, got %d" % c)
        out2 = torch.zeros(2, dtype=torch.float32, device=score.device)
        call("npp_edge_loss_fwd", fptr(score), i32(n), i32(h), i32(w), fptr(target), i32(target.shape[1]),
             i32(target.shape[2]), i32(ignore_index), i32(align_corners), fptr(posneg), fptr(out2), stream())
        ctx.save_for_backward(score, target, posneg, out2)
        ctx.cfg = (ignore_index, align_corners)
        return out2[0] / out2[1]

    @staticmethod
    def backward(ctx, g):
        score, target, posneg, out2 = ctx.saved_tensors
        ignore_index, align_corners = ctx.cfg
        n, c, h, w = score.shape
        gs = g.reshape(1).float().contiguous()
        if _ce_bwd_sep_enabled():
            lh, lw = target.shape[1], target.shape[2]
            d = _grad_through_upsample(score, lh, lw, align_corners, lambda G, cq: call(
                "npp_edge_loss_grad_pixels", fptr(score), i32(n), i32(h), i32(w), fptr(target), i32(lh), i32(lw),
                i32(ignore_index), i32(align_corners), fptr(posneg), fptr(out2), fptr(gs), fptr(G), i32(cq), stream()))
            return d, None, None, None, None
        d = torch.zeros_like(score)
        call("npp_edge_loss_bwd", fptr(score), i32(n), i32(h), i32(w), fptr(target), i32(target.shape[1]),
             i32(target.shape[2]), i32(ignore_index), i32(align_corners), fptr(posneg), fptr(out2), fptr(gs), fptr(d),
             stream())
        return d, None, None, None, None


class _SqErrSumFn(Function):
    """sum((w*(pred - target))^2) — the numerator of every nn.MSELoss in Criterion_pose.joint_loss."""

    @staticmethod
    def forward(ctx, pred, target, row_w, row_len):
        out = torch.zeros(1, dtype=torch.float32, device=pred.device)
        call("npp_mse_fwd", fptr(pred), fptr(target), i64(pred.numel()), fptr(row_w), i64(row_len), fptr(out), stream())
        ctx.save_for_backward(pred, target, row_w)
        ctx.row_len = row_len
        return out[0]

    @staticmethod
    def backward(ctx, g):
        pred, target, row_w = ctx.saved_tensors
        d = torch.empty_like(pred)
        gs = g.reshape(1).float().contiguous()
        call("npp_mse_bwd", fptr(pred), fptr(target), i64(pred.numel()), fptr(row_w), i64(ctx.row_len), fptr(gs),
             fptr(d), stream())
        return d, None, None, None


def ohem_cross_entropy(score, target, weight, ignore_index=255, thresh=0.7, min_kept=100000, align_corners=False):
    score, target = _prep_logits(score), _prep_target(target)
    weight = weight.to(device=score.device, dtype=torch.float32).contiguous()
    return _OhemCEFn.apply(score, target, weight, int(ignore_index), float(thresh), int(max(1, min_kept)),
                           int(bool(align_corners)))


class OhemCrossEntropy(nn.Module):
    """criterion.py:43-72.  `forward(score, target)` accepts logits at any resolution <= the label's; the
    bilinear resize the reference applies first (:57-58, align_corners=False) is fused into the kernel."""

    def __init__(self, ignore_index=255, thres=0.7, min_kept=100000, weight=weights_lip):
        super().__init__()
        self.thresh = thres
        self.min_kept = max(1, min_kept)
        self.ignore_index = ignore_index
        self.register_buffer("weight", weight.clone().float(), persistent=False)

    def forward(self, score, target, align_corners=False, **kwargs):
        return ohem_cross_entropy(score, target, self.weight, self.ignore_index, self.thresh, self.min_kept,
                                  align_corners)


class Criterion_pose(nn.Module):
    """criterion.py:74-145: per stage (sum_j MSE(pred_j, gt_j) + sum_j MSE(aux_j, gtaux_j)) / J, combined over
    stages as loss_i * exp(-lamda_i) + lamda_i."""

    def __init__(self, out_len=1, use_target_weight=False):
        super().__init__()
        self.use_target_weight = use_target_weight
        self.lamda = nn.Parameter(-2.5 * torch.ones(out_len))

    def _mse_sum_over_joints(self, pred, gt, target_weight):
        """sum_j nn.MSELoss()(pred[:, j], gt[:, j]) = sum((pred-gt)^2) / (B*H*W)   (:98-108)."""
        if pred.shape != gt.shape:
            raise RuntimeError("heat-map prediction %s and target %s differ in shape; resize is not implemented "
                               "(the reference configs use equal sizes)" % (tuple(pred.shape), tuple(gt.shape)))
        pred, gt = _prep_logits(pred), gt.contiguous().float()
        b, j, h, w = pred.shape
        row_w, row_len = None, 0
        if self.use_target_weight:
            row_w = target_weight.reshape(b * j).contiguous().float()
            row_len = h * w
        return _SqErrSumFn.apply(pred, gt, row_w, row_len) / float(b * h * w)

    def joint_loss(self, output, target, target_weight=None):
        if isinstance(output, list):
            main, aux = output[0], output[1]
            gt, gt_aux = target[0], target[1]
        else:
            main, aux, gt_aux = output, None, None
            gt = target[0] if isinstance(target, list) else target
        num_joints = main.size(1)
        loss = self._mse_sum_over_joints(main, gt, target_weight)
        if aux is not None:
            loss = loss + self._mse_sum_over_joints(aux, gt_aux, target_weight)
        return loss / num_joints

    def forward(self, output, target, target_weight=None):
        if not isinstance(output, list):
            raise RuntimeError("Criterion_pose expects the list of per-stage predictions Network.forward returns "
                               "(the reference's non-list branch is dead code: criterion.py:144 uses an undefined name)")
        loss = 0.
        for i in range(len(output)):
            loss = loss + self.joint_loss(output[i], target, target_weight) * torch.exp(-self.lamda[i]) + self.lamda[i]
        return loss


class Criterion_par(nn.Module):
    """criterion.py:148-217: per stage OHEM parsing CE + class-balanced edge CE on logits bilinearly upsampled
    (align_corners=True) to the label resolution, combined over stages with learned uncertainty weights."""

    def __init__(self, out_len=1, ignore_index=255, thres=0.9, min_kept=131072, weight=weights_lip):
        super().__init__()
        self.ignore_index = ignore_index
        self.criterion = OhemCrossEntropy(ignore_index=ignore_index, thres=thres, min_kept=min_kept, weight=weight)
        self.lamda = nn.Parameter(2.3 * torch.ones(out_len))

    def _ohem(self, logits, label):
        return self.criterion(logits, label, align_corners=True)

    def parsing_loss(self, preds, target):
        par_label, edge_label = _prep_target(target[0]), _prep_target(target[1])
        loss = 0.
        if not isinstance(preds, list):
            return self._ohem(preds, par_label)
        preds_parsing = preds[0]
        if isinstance(preds_parsing, list):  # criterion.py:175-185
            loss = loss + self._ohem(preds_parsing[0], par_label) + self._ohem(preds_parsing[1], par_label) * 0.4
        else:
            loss = loss + self._ohem(preds_parsing, par_label)
        # class-balancing weights from this batch's label counts (:161-166) — counted on the device, no host sync
        posneg = torch.zeros(2, dtype=torch.int64, device=edge_label.device)
        call("npp_edge_count", fptr(edge_label), i64(edge_label.numel()), fptr(posneg), stream())
        preds_edge = preds[1]
        for pe in (preds_edge if isinstance(preds_edge, list) else [preds_edge]):
            loss = loss + _EdgeCEFn.apply(_prep_logits(pe), edge_label, posneg, int(self.ignore_index), 1)
        return loss

    def forward(self, preds, target):
        if not isinstance(preds, list):
            return self.criterion(preds, target) * torch.exp(-self.lamda) + self.lamda
        loss = 0.
        for i in range(len(preds)):
            loss = loss + self.parsing_loss(preds[i], target) * torch.exp(-self.lamda[i]) + self.lamda[i]
        return loss
