"""Search supernet of NPPNet (parsing <-> pose interaction search) on libnpp_b200 kernels.

Drop-in for the reference's models/model_search_interact.py: `Network(cfg)` reads the same config fields (:437-444),
registers the same module tree (identical state_dict keys, identical seeded init), owns the same twelve architecture
tensors (`arch_parameters()`, :772-804) and `forward(x)` returns the same `(pose_list, par_list)` structure (:626-770);
`loss_entropy()` (:881-896), `genotype()` (:913-1051) and `btw()` (:1054-1065) keep their semantics.

What executes differently (same arithmetic, B200-first schedule):
  * MixedOp (:39-74): the seven candidate primitives run at C/2 on a channel-slice VIEW of the input (no split
    copy); their BatchNorm(affine=False) applies, the alpha-weighted sum, the concat with the pass-through half and
    channel_shuffle(2) are ONE kernel (functional.mix -> npp_mix_fwd); backward produces every branch gradient, the
    pass-through gradient and d alpha from two passes (npp_mix_bwd_reduce / npp_mix_bwd_apply).
  * cross-scale MixedOps: bilinear resampling is linear with weights summing to one, so the seven per-branch
    `Interpolate(up_scale)` (:50-51) commute with the weighted sum and are applied once to the mixed half.
  * nodes `s = base + sum_j beta_j * MixedOp_j(h_j, alpha_j)` (:352-356, :648-654): one n-ary weighted-sum kernel,
    d beta_j = <g, MixedOp_j> from the same reduction.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import functional as F_
from ..nn import BatchNorm2d, Conv2d, MaxPool2d, Sequential, call_lazy
from . import genotypes as gt
from .genotypes import PRIMITIVES_INTER, Genotype_fuse, Genotype_inter
from .model_augment import Cell, Interpolate, ParCell1, PoseCell1, Upsample, _head, _layer_1x1, _stem, _StepCell
from .operations import OPS, ReLUConvBN

BN_MOMENTUM = 0.1


def channel_shuffle(x, groups):
    """model_search_interact.py:22-36 on a plain NCHW tensor (kept for API parity; MixedOp fuses the groups=2 case)."""
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).contiguous().view(b, -1, h, w)


class MixedOp(nn.Module):
    """model_search_interact.py:39-74.  `_ops[k]` has the reference's structure (OPS[p](C//2, stride, False), wrapped
    in Sequential(op, BatchNorm2d(affine=False)) for pooling and Sequential(op, Interpolate(up_scale)) when rescaling);
    `_core[k]` is the same module without the Interpolate wrapper, which forward applies once after the sum."""

    def __init__(self, C, stride, up_scale=None, extra_conv=None):
        super().__init__()
        self._ops = nn.ModuleList()
        self.mp = MaxPool2d(2, 2)
        self._core = []
        for primitive in PRIMITIVES_INTER:
            op = OPS[primitive](C // 2, stride, False)
            if "pool" in primitive:
                op = Sequential(op, BatchNorm2d(C // 2, affine=False))
            self._core.append(op)
            if up_scale:
                op = Sequential(op, Interpolate(scale_factor=up_scale))
            self._ops.append(op)
        self.up_scale = up_scale
        self.stride = stride
        self.extra_conv = extra_conv

    def forward(self, x, weights):
        x = F_.check_raw(F_.to_internal(x), "MixedOp")
        kinds = [_StepCell._kind(op) for op in self._core]
        n_relu = sum(k == "relu" for k in kinds)
        # xtemp (:59) is read by all candidates: raw by the pooling/SE/depthwise ones, through nn.ReLU by the dense
        # convolutions (one shared ReLU pass); xtemp2 (:60) passes through
        los, hi = F_.split_halves(x, (len(kinds) - n_relu) + (1 if n_relu else 0))
        relus = F_.fanout(F_.relu(los[-1]), n_relu) if n_relu else []
        branches = []
        for op, kind in zip(self._core, kinds):
            h = relus.pop() if kind == "relu" else los.pop(0)
            branches.append(call_lazy(op, h))
        rescale = bool(self.up_scale) and float(self.up_scale) != 1.0  # scale 1.0 resamples are exact identities
        if not rescale and self.stride == 1:
            ans = F_.mix(branches, weights, pass_=hi)
        else:
            t = F_.mix(branches, weights)
            if rescale:
                t = F_.interpolate(t, scale_factor=self.up_scale, mode="bilinear", align_corners=True)
                hi = F_.interpolate(hi, scale_factor=self.up_scale)   # :63-64 default (nearest) mode
            if t.shape[2] != hi.shape[2]:
                hi = self.mp(hi)                                       # :66-69 reduction MixedOp
            ans = F_.interleave2(t, hi)
        if self.extra_conv is not None:
            return self.extra_conv(ans)
        return ans


def _weighted_node(ops, states, alphas, betas, base=None):
    """base + sum_j betas[j] * ops[j](states[j], alphas[j]) in one pass (:352-356, :648-654)."""
    # the MixedOps of one node are independent of each other: worker streams (functional.parallel_branches)
    terms = F_.parallel_branches([((lambda op=op, h=h, a=a: op(h, a)), (h, a)) for op, h, a in zip(ops, states, alphas)])
    if base is None:
        return F_.mix(terms, betas)
    one = torch.ones(1, dtype=betas.dtype, device=betas.device)
    return F_.mix([base] + terms, torch.cat([one, betas]))


class Upsample1(nn.Module):
    """Searchable decoder cell — model_search_interact.py:124-160 (not instantiated by Network)."""

    def __init__(self, steps, multiplier, C_prev_prev, C_prev):
        super().__init__()
        self.preprocess0 = ReLUConvBN(C_prev_prev, C_prev // 4, 1, 1, 0, affine=True)
        self.preprocess1 = ReLUConvBN(C_prev, C_prev // 4, 1, 1, 0, affine=True)
        self._steps = steps
        self._multiplier = multiplier
        self._ops = nn.ModuleList()
        self._bns = nn.ModuleList()
        for i in range(steps):
            for j in range(2 + i):
                self._ops.append(MixedOp(C_prev // 4, 1, 2 if j == 0 else None))

    def forward(self, s0, s1, weights, weights2):
        states = [self.preprocess0(s0), self.preprocess1(s1)]
        offset = 0
        for _ in range(self._steps):
            n = len(states)
            states.append(_weighted_node(self._ops[offset:offset + n], states, weights[offset:offset + n],
                                         weights2[offset:offset + n]))
            offset += n
        return F_.cat(states[-self._multiplier:])


class _MixedFusionCell(nn.Module):
    """PoseCell / ParCell — model_search_interact.py:332-429: three preprocessed inputs, `steps` nodes each summing
    one MixedOp per earlier state; returns (cat(states[0:3]), cat(states[-multiplier:]))."""

    def __init__(self, steps, multiplier, C_prev_prev, C_prev, C_cur, order):
        super().__init__()
        if order == 0:
            cins = (C_prev_prev, C_prev, C_cur)
        else:
            cins = (3 * C_prev, 4 * C_prev, 4 * C_prev)
        self.preprocess0 = ReLUConvBN(cins[0], C_cur, 1, 1, 0, affine=True)
        self.preprocess1 = ReLUConvBN(cins[1], C_cur, 1, 1, 0, affine=True)
        self.preprocess2 = ReLUConvBN(cins[2], C_cur, 1, 1, 0, affine=True)
        self._steps = steps
        self._multiplier = multiplier
        self.order = order
        self._ops = nn.ModuleList()
        for i in range(steps):
            for j in range(3 + i):
                up_scale = None
                if order == 0:
                    up_scale = 4 if j == 0 else (2 if j == 1 else None)
                self._ops.append(MixedOp(C_cur, 1, up_scale))

    def forward(self, s0, s1, s2, weights, weights2):
        steps, nst = self._steps, 3 + self._steps
        states = [self.preprocess0(s0), self.preprocess1(s1), self.preprocess2(s2)]
        # state k is read by one MixedOp of every later node and by the output concats it belongs to: hand out that
        # many handles so the backward sums their gradients in one pass
        def readers(k):
            later = sum(1 for i in range(steps) if 3 + i > k)
            return later + (1 if k < 3 else 0) + (1 if k >= nst - self._multiplier else 0)

        handles = [F_.fanout(s, readers(k)) for k, s in enumerate(states)]
        offset = 0
        for i in range(steps):
            n = 3 + i
            s = _weighted_node(self._ops[offset:offset + n], [handles[k].pop() for k in range(n)],
                               weights[offset:offset + n], weights2[offset:offset + n])
            offset += n
            handles.append(F_.fanout(s, readers(n)))
        first = [handles[k].pop() for k in range(3)]
        if self.order == 0:   # :372-374 default-mode (nearest) F.interpolate
            first[0] = F_.interpolate(first[0], scale_factor=4)
            first[1] = F_.interpolate(first[1], scale_factor=2)
        last = [handles[k].pop() for k in range(nst - self._multiplier, nst)]
        return F_.cat(first), F_.cat(last)


class PoseCell(_MixedFusionCell):
    pass


class ParCell(_MixedFusionCell):
    pass


class Network(nn.Module):
    """model_search_interact.py:432-1089."""

    def __init__(self, cfg, steps=4, multiplier=4):
        super().__init__()
        self._num_classes = cfg.DATASET.NUM_CLASSES
        self._num_joints = cfg.DATASET.NUM_JOINTS
        self._layers = cfg.SEARCH.LAYERS
        self._steps = steps
        self._multiplier = multiplier
        self.C = cfg.SEARCH.INIT_CHANNELS
        self._head = cfg.MODEL.HEAD
        self.refine_layers = cfg.MODEL.REFINE_LAYERS
        C, L = self.C, self._layers

        self.stem0 = _stem(3, C, 2, True)
        self.stem1 = _stem(C, 2 * C, 2, True)
        self.stem2 = _stem(2 * C, 2 * C, 1, False)
        self.stem3 = _stem(3, C, 2, True)
        self.stem4 = _stem(C, 2 * C, 2, True)
        self.stem5 = _stem(2 * C, 2 * C, 1, False)

        # encoder with the fixed ENCODER genotype (:482-505)
        self._tap_layers = [L // 4 - 1, 2 * L // 4 - 1, 3 * L // 4 - 1, 4 * L // 4 - 1]
        reduce_layers = [L // 4, 2 * L // 4, 3 * L // 4]
        C_pp, C_p, C_cur = 2 * C, 2 * C, int(C / 2)
        self.cells1 = nn.ModuleList()
        self.cells2 = nn.ModuleList()
        widths = []
        reduction_prev = False
        for i in range(L):
            if i in self._tap_layers:
                widths.append(int(C_cur * multiplier))
            reduction = i in reduce_layers
            if reduction:
                C_cur *= 2
            self.cells1.append(Cell(gt.ENCODER, C_pp, C_p, C_cur, reduction, reduction_prev))
            self.cells2.append(Cell(gt.ENCODER, C_pp, C_p, C_cur, reduction, reduction_prev))
            reduction_prev = reduction
            C_pp, C_p = C_p, multiplier * C_cur
        self.num_inchannels = nin = widths[::-1]

        # searchable encoder interactions: stage i mixes the other stream's scales 0..i (:509-528)
        self._ops1 = nn.ModuleList()
        self._ops2 = nn.ModuleList()
        for i in range(len(nin)):
            for j in range(1 + i):
                up_scale = 1 / 2 ** (i - j)
                conv1 = Conv2d(nin[3 - j], nin[3 - i], 1) if i != j else None
                conv2 = Conv2d(nin[3 - j], nin[3 - i], 1) if i != j else None
                self._ops1.append(MixedOp(nin[3 - j], 1, up_scale, conv1))
                self._ops2.append(MixedOp(nin[3 - j], 1, up_scale, conv2))

        # decoder with the fixed DECODER genotype (:530-540)
        self.upsamples1 = nn.ModuleList()
        self.upsamples2 = nn.ModuleList()
        for j in range(len(nin) - 1):
            self.upsamples1.append(Upsample(gt.DECODER.upsample1, gt.DECODER.upsample_concat1, nin[j], nin[j + 1]))
        for j in range(len(nin) - 1):
            self.upsamples2.append(Upsample(gt.DECODER.upsample2, gt.DECODER.upsample_concat2, nin[j], nin[j + 1]))

        # searchable decoder interactions over the 7-entry feature list (:542-565)
        self.up_ops1 = nn.ModuleList()
        self.up_ops2 = nn.ModuleList()
        resolution = [1, 1 / 2, 1 / 4, 1 / 8, 1 / 4, 1 / 2, 1]
        channels = [int(2 * C / r) for r in resolution]
        for i in range(len(resolution) - 4):
            for j in range(4 + 1 + i):
                up_scale = resolution[4 + i] / resolution[j]
                conv1 = Conv2d(channels[j], channels[4 + i], 1) if 4 + i != j else None
                conv2 = Conv2d(channels[j], channels[4 + i], 1) if 4 + i != j else None
                self.up_ops1.append(MixedOp(channels[j], 1, up_scale, conv1))
                self.up_ops2.append(MixedOp(channels[j], 1, up_scale, conv2))

        C3 = nin[3]
        self.pose_layer = _layer_1x1(8 * C3, 4 * C3)
        self.pose_auxlayer = _layer_1x1(8 * C3, 3 * C3)
        self.par_layer = _layer_1x1(8 * C3, 4 * C3)
        self.edge_layer = _layer_1x1(8 * C3, 3 * C3)

        self.pose_net = nn.ModuleList()
        self.par_net = nn.ModuleList()
        for _ in range(3):
            self.pose_net.append(PoseCell(4, 4, C3, C3, C3, 1))
            self.par_net.append(ParCell(4, 4, C3, C3, C3, 1))

        self.pose_head = nn.ModuleList()
        self.pose_auxnet = nn.ModuleList()
        self.par_head = nn.ModuleList()
        self.edge_head = nn.ModuleList()
        for _ in range(self.refine_layers + 1):
            self.pose_head.append(_head(4 * C3, 256, self._num_joints, 1))
            self.pose_auxnet.append(_head(3 * C3, 128, self._num_joints, 3))
            self.par_head.append(_head(4 * C3, 256, self._num_classes, 1))
            self.edge_head.append(_head(3 * C3, 6, 2, 3, first_bias=False))
        self.init_weights()
        self._initialize_alphas()

    # ------------------------------------------------------------------ forward
    def _interact(self, ops, offset, feats, alphas, betas, base):
        n = len(feats)
        wa = torch.softmax(alphas[offset:offset + n], dim=-1)
        wb = torch.softmax(betas[offset:offset + n], dim=-1)
        return _weighted_node(ops[offset:offset + n], feats, wa, wb, base=base)

    def forward(self, x):
        x = F_.to_internal(x)
        # pose stream on the current CUDA stream, parsing stream on the side stream (functional.TaskStreams; see
        # models/model_augment.py Network.forward): joined at every interaction point and at the heads
        ts = F_.TaskStreams(x)
        if ts.on and x.dtype == torch.bfloat16 and x.shape[1] == 8 and not x.requires_grad and F_._state.get("stem_im2col", True):
            F_.im2col3x3_c3(x, 2, 1)
        ts.fork(x)
        s0 = self.stem1(self.stem0(x))
        s1 = self.stem2(s0)
        with ts.side():
            s2 = self.stem4(self.stem3(x))
            s3 = self.stem5(s2)
        f1, f2 = [], []
        offset = 0
        for i, (cell1, cell2) in enumerate(zip(self.cells1, self.cells2)):
            tap = i in self._tap_layers
            s0, s1 = s1, cell1(s0, s1, out_raw=tap, out_relu=True)
            with ts.side():
                s2, s3 = s3, cell2(s2, s3, out_raw=tap, out_relu=True)
            if tap:
                ts.join(s3)
                f1.append(s1)
                f2.append(s3)
                n1 = self._interact(self._ops1, offset, f2, self.alphas1, self.betas1, s1)   # :646-651
                n3 = self._interact(self._ops2, offset, f1, self.alphas2, self.betas2, s3)   # :652-653 (f1 still old)
                s1, s3 = n1, n3
                f1[-1], f2[-1] = s1, s3
                offset += len(f1)
                ts.fork(s3)

        cont = 0
        prev1, prev2 = f1[3], f2[3]
        for d in range(3):                                                                  # :665-729
            o1 = self.upsamples1[d](prev1, f1[2 - d])
            with ts.side():
                o2 = self.upsamples2[d](prev2, f2[2 - d])
            ts.join(o2)
            f1.append(o1)
            f2.append(o2)
            n1 = self._interact(self.up_ops1, cont, f2, self.alphas3, self.betas3, o1)
            n2 = self._interact(self.up_ops2, cont, f1, self.alphas4, self.betas4, o2)
            f1[-1], f2[-1] = n1, n2
            prev1, prev2 = n1, n2
            cont += len(f1)
            ts.fork(n2)

        def pyramid(f):                                                                     # :733-738
            return F_.cat_relu([f[0], f[6],
                                F_.interpolate(f[5], scale_factor=2, mode="bilinear", align_corners=True),
                                F_.interpolate(f[4], scale_factor=4, mode="bilinear", align_corners=True)])

        x1 = pyramid(f1)
        in1 = self.pose_auxlayer(x1)
        in3 = self.pose_layer(x1)
        with ts.side():
            x2 = pyramid(f2)
            in2 = self.edge_layer(x2)
            in4 = self.par_layer(x2)
        pose_list, par_list = [], []
        side_outputs = []

        def emit(k):
            pose_aux = self.pose_auxnet[k](in1)
            pose_map = self.pose_head[k](in3)
            pose_list.append([F_.from_internal(pose_map, self._num_joints), F_.from_internal(pose_aux, self._num_joints)])
            with ts.side():
                edge = self.edge_head[k](in2)
                par_map = self.par_head[k](in4)
                par_list.append([F_.from_internal(par_map, self._num_classes), F_.from_internal(edge, 2)])
            side_outputs.extend(par_list[-1])

        emit(0)
        w_pose = torch.softmax(self.alphas_pose, dim=-1)
        w_pose2 = self.btw(3, self._steps, self.betas_pose)
        w_par = torch.softmax(self.alphas_par, dim=-1)
        w_par2 = self.btw(3, self._steps, self.betas_par)
        for i in range(1, self.refine_layers + 1):
            for j in range(3):
                # PoseCell on the main stream, ParCell on the side stream; each reads the other's fea2: barrier first
                ts.join(in4)
                ts.fork(in3, w_par, w_par2)
                in1, tmp = self.pose_net[2 * (i - 1) + j](in1, in3, in4, w_pose, w_pose2)
                with ts.side():
                    in2, in4 = self.par_net[2 * (i - 1) + j](in2, in3, in4, w_par, w_par2)
                in3 = tmp
            emit(i)
        ts.join(side_outputs)
        return pose_list, par_list

    # ------------------------------------------------------------------ architecture parameters
    def _initialize_alphas(self):
        """:772-804 — 1e-3 * ones; rows = one MixedOp each, columns = PRIMITIVES_INTER."""
        k = sum(3 + i for i in range(self._steps))
        num_ops = len(PRIMITIVES_INTER)

        def p(*shape):
            return nn.Parameter(1e-3 * torch.ones(*shape))

        self.alphas1, self.alphas2 = p(10, num_ops), p(10, num_ops)
        self.alphas3, self.alphas4 = p(18, num_ops), p(18, num_ops)
        self.betas1, self.betas2, self.betas3, self.betas4 = p(10), p(10), p(18), p(18)
        self.alphas_pose, self.alphas_par = p(k, num_ops), p(k, num_ops)
        self.betas_pose, self.betas_par = p(k), p(k)
        self._arch_parameters = [self.alphas1, self.alphas2, self.alphas3, self.alphas4, self.alphas_pose,
                                 self.alphas_par, self.betas1, self.betas2, self.betas3, self.betas4, self.betas_pose,
                                 self.betas_par]

    def arch_parameters(self):
        return self._arch_parameters

    def btw(self, n_input, steps, betas):
        """Segment-wise softmax of betas: node i normalises over its n_input + i incoming edges (:1054-1065)."""
        out, start = [], 0
        for i in range(steps):
            n = n_input + i
            out.append(torch.softmax(betas[start:start + n], dim=-1))
            start += n
        return torch.cat(out, dim=0)

    def loss_entropy(self):
        """0.5 * sum_i mean_rows(H(softmax(alpha_i)) / ln(#ops)) / 12 (:881-896)."""
        params = self._arch_parameters
        alphas = params[:len(params) // 2]
        total = 0.
        for a in alphas:
            w = torch.softmax(a, dim=-1)
            # Categorical(probs=w).entropy(): probabilities re-normalised and clamped to [eps, 1-eps] before the log
            pr = w / w.sum(-1, keepdim=True)
            eps = torch.finfo(pr.dtype).eps
            logits = torch.log(pr.clamp(min=eps, max=1 - eps))
            ent = -(logits * pr).sum(-1)
            total = total + (ent / math.log(w.shape[1])).mean(dim=0)
        return 0.25 * 2 * total / len(params)

    def genotype(self):
        """CPU decode of the architecture tensors (:913-1051)."""
        def scaled(alpha, w2, start, n):
            W = torch.softmax(alpha, dim=-1).data.cpu().numpy()[start:start + n].copy()
            return W * w2.data.cpu().numpy()[start:start + n, None]

        def parse_cumulative(alpha, w2, n_input, step):
            """Per node: take the largest alpha*beta entries until their sum reaches 0.7 or four are taken (:958-993)."""
            gene, start, n = [], 0, n_input
            for _ in range(step):
                W = scaled(alpha, w2, start, n)
                prob, picked = 0., []
                while prob < 0.7 and len(picked) < 4:
                    m = np.max(W)
                    prob += m
                    where = np.where(W == m)
                    W[where] = 0
                    picked.append((PRIMITIVES_INTER[where[1][0]], where[0][0]))
                gene.append(picked)
                start += n
                n += 1
            return gene

        def parse_top2(alpha, w2):
            """Per node: the two incoming edges with the largest best-op weight, each with its best op (:995-1015)."""
            gene, start, n = [], 0, 3
            for i in range(self._steps):
                W = scaled(alpha, w2, start, n)
                edges = sorted(range(i + 3), key=lambda e: -max(W[e]))[:2]
                for j in edges:
                    k_best = 0
                    for k in range(len(W[j])):
                        if W[j][k] > W[j][k_best]:
                            k_best = k
                    gene.append((PRIMITIVES_INTER[k_best], j))
                start += n
                n += 1
            return gene

        genotype_inter = Genotype_inter(
            task1=parse_cumulative(self.alphas1, self.btw(1, 4, self.betas1), 1, 4),
            task2=parse_cumulative(self.alphas2, self.btw(1, 4, self.betas2), 1, 4),
            task3=parse_cumulative(self.alphas3, self.btw(5, 3, self.betas3), 5, 3),
            task4=parse_cumulative(self.alphas4, self.btw(5, 3, self.betas4), 5, 3))
        genotype_fuse = Genotype_fuse(
            pose=parse_top2(self.alphas_pose, self.btw(3, self._steps, self.betas_pose)), pose_concat=range(3, 7),
            par=parse_top2(self.alphas_par, self.btw(3, self._steps, self.betas_par)), par_concat=range(3, 7))
        return genotype_inter, genotype_fuse

    def init_weights(self, pretrained=""):
        """xavier_normal on conv weights, zero conv bias, BN weight 1 / bias 0 (:1067-1089), in modules() order."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, BatchNorm2d):
                if m.affine:
                    m.weight.data.fill_(1)
                    m.bias.data.zero_()
