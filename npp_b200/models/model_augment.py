"""Derived NPPNet (joint human parsing + pose) on libnpp_b200 kernels.

Drop-in for the reference's models/model_augment.py: `Network(cfg)` takes the same config fields
(model_augment.py:236-242), builds the same module tree in the same order (identical state_dict
keys and identical xavier initialisation under a given torch seed) and `forward(x)` returns the
same `(pose_list, par_list)` structure of NCHW fp32 tensors (model_augment.py:402-574).  Internally
activations are NHWC bf16 (or fp32 in validation mode) and every op is one of our kernels.
"""
import os

import torch
import torch.nn as nn

from .. import functional as F_
from ..nn import BatchNorm2d, Conv2d, ReLU, Sequential, call_lazy
from . import genotypes as gt
from .operations import OPS, FactorizedReduce, ReLUConvBN

# NPP_CELL_BRANCHES=0: one primitive after the other.  Default: the primitives of a cell whose inputs exist run
# concurrently on worker streams (train 88.34 -> 86.94 ms, infer512 40.99 -> 40.25 ms, profiles/r02_cell_branches_ab.txt);
# without effect under SyncBN (functional.parallel_branches keeps the exchange order of a stream fixed)
_CELL_BRANCHES = os.environ.get("NPP_CELL_BRANCHES", "1") != "0"


def _preprocess(layers, inputs):
    """The 1x1 preprocess layers of a cell (independent of each other)."""
    if not _CELL_BRANCHES:
        return [call_lazy(l, x) for l, x in zip(layers, inputs)]
    return F_.parallel_branches([((lambda l=l, x=x: call_lazy(l, x)), x) for l, x in zip(layers, inputs)])

BN_MOMENTUM = 0.1


class Interpolate(nn.Module):
    """F.interpolate(scale_factor, 'bilinear', align_corners=True) — model_augment.py:109-116."""

    def __init__(self, scale_factor, mode="bilinear"):
        super().__init__()
        self.s = scale_factor
        self.mode = mode

    def forward(self, x):
        return F_.interpolate(F_.to_internal(x), scale_factor=self.s, mode=self.mode, align_corners=True)


class _RescaleProject(Sequential):
    """`nn.Sequential(Interpolate(s), nn.Conv2d(C_in, C_out, 1))` of model_augment.py:592-596 / 642-646.
    Bilinear weights sum to one, so the 1x1 convolution (with bias) commutes with the resampling;
    when upsampling we project first and resample the narrower/smaller tensor.
    Child indices (0: Interpolate, 1: Conv2d) match the reference for state_dict parity."""

    def forward(self, x):
        resample, proj = self[0], self[1]
        if resample.s > 1:
            return resample(proj(x))
        return proj(resample(x))


def _edge_lists(edges):
    names, idx = zip(*edges)
    return list(names), list(idx)


class _StepCell(nn.Module):
    """Shared evaluation of a DARTS-style cell: every step adds the outputs of two primitives applied to
    earlier states (model_augment.py:48-62, 90-106, 153-174, 210-229)."""

    def _build_ops(self, C, edges, wrap=None):
        names, self._indices = _edge_lists(edges)
        assert len(names) % 2 == 0
        self._steps = len(names) // 2
        self._ops = nn.ModuleList()
        for k, (name, index) in enumerate(zip(names, self._indices)):
            stride = self._stride_for(index)
            op = OPS[name](C, stride, True)
            if wrap is not None:
                op = wrap(op, index)
            self._ops.append(op)

    def _stride_for(self, index):
        return 1

    # Which of {state, relu(state)} the cell's output concat(s) carry is chosen per call: `cell(s0, s1)` returns the
    # raw concat like the reference; Network.forward passes out_raw / out_relu when it knows that the consumers start
    # with nn.ReLU (cells feeding only preprocess layers / heads), so the raw concat is never written.

    def _run_steps(self, states):
        """Plain evaluation (one tensor per state)."""
        for i in range(self._steps):
            a = self._ops[2 * i](states[self._indices[2 * i]])
            b = self._ops[2 * i + 1](states[self._indices[2 * i + 1]])
            states.append(F_.add(a, b))
        return states

    @staticmethod
    def _kind(op):
        k = getattr(op, "input_kind", None)
        if k is None and isinstance(op, nn.Sequential) and len(op) > 0:
            return _StepCell._kind(op[0])
        return k or "raw"

    def _needs(self, nstates, concats, out_raw=True, out_relu=False):
        """Per state: (raw needed, relu needed, (concat id, slot) or None) from the primitives that read it and
        from its concat membership."""
        kinds = [set() for _ in range(nstates)]
        for op, idx in zip(self._ops, self._indices):
            kinds[idx].add(self._kind(op))
        member = {}
        for cid, cat in enumerate(concats):
            for slot, k in enumerate(cat):
                if k in member:
                    return None  # a state in two concats: not expressible as slices, use the plain path
                member[k] = (cid, slot)
        out = []
        for k in range(nstates):
            raw = "raw" in kinds[k] or (k in member and out_raw)
            rel = "relu" in kinds[k] or (k in member and out_relu)
            if "relu_ok" in kinds[k] and not rel:
                raw = True  # the depthwise kernel applies the ReLU inside its own loads
            if not (raw or rel):
                raw = True
            out.append((raw, rel, member.get(k)))
        return out

    def _run_fused(self, pre, concats, out_raw=True, out_relu=False, n_out=None):
        """Evaluates the cell with one fused pass per state (functional.node): BatchNorm-apply of both operands,
        the add, the ReLU of the consumers and the write into the output concat buffer(s).
        pre: outputs of the preprocess layers (Pending BatchNorm outputs); returns one handle per concat — or, with
        n_out = [k per concat], a list of k handles per concat, one per CONSUMER of that cell output (the gradients
        of the consumers then meet inside the member nodes' backward kernels, not in autograd add kernels)."""
        nstates = len(pre) + self._steps
        n_out = [max(1, int(k)) for k in n_out] if n_out is not None else None
        if not (out_raw or out_relu):
            raise ValueError("a cell must produce its output raw, through ReLU, or both")
        needs = self._needs(nstates, concats, out_raw, out_relu)
        if needs is None:
            states = self._run_steps([F_.finish(p) for p in pre])
            outs = [F_.cat([states[i] for i in cat]) for cat in concats]
            return [[o] * n_out[cid] for cid, o in enumerate(outs)] if n_out is not None else outs
        bufs = {}      # (concat id, "raw" | "relu") -> buffer
        slices = {}    # same key -> list of slice tensors in slot order
        handles = []   # per state: {op index: the handle that primitive reads}
        # which of {state, relu(state)} every primitive reads (decides the per-consumer handles below)
        readers = [[] for _ in range(nstates)]
        for j, (op, idx) in enumerate(zip(self._ops, self._indices)):
            readers[idx].append((j, self._kind(op)))

        def emit(k, a, b):
            raw_w, rel_w, mem = needs[k]
            ya = a.y if isinstance(a, F_.Pending) else a
            n, c, h, w = ya.shape
            o_raw = o_rel = None
            if mem is not None:
                cid, slot = mem
                for key, wanted in (("raw", raw_w and out_raw), ("relu", rel_w and out_relu)):
                    if not wanted:
                        continue
                    if (cid, key) not in bufs:
                        bufs[(cid, key)] = F_.empty_internal(n, c * len(concats[cid]), h, w, ya.dtype, ya.device)
                        slices[(cid, key)] = [[None] * len(concats[cid]) for _ in range(n_out[cid] if n_out else 1)]
                    t = F_.alias(bufs[(cid, key)], slot * c, c)
                    if key == "raw":
                        o_raw = t
                    else:
                        o_rel = t
            # One handle per consumer (primitives of this cell reading the state raw / through nn.ReLU, plus the concat
            # route): each consumer's gradient reaches the node's backward kernel on its own and is summed there
            # instead of by autograd's add kernels.
            reads = []
            for j, kind in readers[k]:
                if kind == "relu" or (kind == "relu_ok" and rel_w):
                    reads.append((j, "relu"))
                else:
                    reads.append((j, "raw"))
            n_cat = (n_out[mem[0]] if n_out else 1) if mem is not None else 0
            n_raw = sum(1 for _, t in reads if t == "raw") + (n_cat if o_raw is not None else 0)
            n_rel = sum(1 for _, t in reads if t == "relu") + (n_cat if o_rel is not None else 0)
            raws, rels = F_.node(a, b, want_raw=raw_w, want_relu=rel_w, out_raw=o_raw, out_relu=o_rel,
                                 fan=(n_raw, n_rel))
            raws, rels = list(raws), list(rels)
            for q in range(n_cat):                                  # the concat route(s) take the last handles
                if o_raw is not None:
                    slices[(mem[0], "raw")][q][mem[1]] = raws.pop()
                if o_rel is not None:
                    slices[(mem[0], "relu")][q][mem[1]] = rels.pop()
            per_op = {}
            for j, t in reads:
                if t == "relu":
                    per_op[j] = F_.state_handle(None, rels.pop(0))
                else:
                    per_op[j] = F_.state_handle(raws.pop(0), None)
            handles.append(per_op)

        for k, p in enumerate(pre):
            emit(k, p, None)
        if not _CELL_BRANCHES:
            for i in range(self._steps):
                ja, jb = 2 * i, 2 * i + 1
                a = call_lazy(self._ops[ja], handles[self._indices[ja]][ja])
                b = call_lazy(self._ops[jb], handles[self._indices[jb]][jb])
                emit(len(pre) + i, a, b)
        else:
            # Wave schedule: every primitive whose input state exists runs now, each on its own worker stream
            # (functional.parallel_branches); then the nodes whose two primitives are done are emitted in order.
            done, nxt = {}, 0
            while nxt < self._steps:
                wave = [j for j in range(2 * self._steps) if j not in done and self._indices[j] < len(handles)]
                res = F_.parallel_branches([((lambda j=j: call_lazy(self._ops[j], handles[self._indices[j]][j])),
                                             handles[self._indices[j]][j]) for j in wave])
                done.update(zip(wave, res))
                while nxt < self._steps and 2 * nxt in done and 2 * nxt + 1 in done:
                    emit(len(pre) + nxt, done[2 * nxt], done[2 * nxt + 1])
                    nxt += 1
        outs = []
        for cid in range(len(concats)):
            hs = []
            for q in range(n_out[cid] if n_out else 1):
                raw = F_.assemble(bufs[(cid, "raw")], slices[(cid, "raw")][q]) if (cid, "raw") in bufs else None
                rel = F_.assemble(bufs[(cid, "relu")], slices[(cid, "relu")][q]) if (cid, "relu") in bufs else None
                if raw is not None and rel is not None:
                    raw._npp_relu = rel
                hs.append(F_.state_handle(raw, rel))
            outs.append(hs if n_out else hs[0])
        return outs


class Cell(_StepCell):
    """Encoder cell — model_augment.py:16-62."""

    def __init__(self, genotype, C_prev_prev, C_prev, C, reduction, reduction_prev):
        super().__init__()
        if reduction_prev:
            self.preprocess0 = FactorizedReduce(C_prev_prev, C)
        else:
            self.preprocess0 = ReLUConvBN(C_prev_prev, C, 1, 1, 0, affine=True)
        self.preprocess1 = ReLUConvBN(C_prev, C, 1, 1, 0, affine=True)
        self._reduction = reduction
        if reduction:
            edges, concat = genotype.reduce, genotype.reduce_concat
        else:
            edges, concat = genotype.normal, genotype.normal_concat
        self._concat = concat
        self.multiplier = len(concat)
        self._build_ops(C, edges)

    def _stride_for(self, index):
        return 2 if self._reduction and index < 2 else 1

    def forward(self, s0, s1, out_raw=True, out_relu=False, n_out=None):
        """n_out = k: returns k handles on the cell output, one per consumer (see _run_fused)."""
        return self._run_fused(_preprocess((self.preprocess0, self.preprocess1), (s0, s1)),
                               [list(self._concat)], out_raw, out_relu, [n_out] if n_out else None)[0]


class Upsample(_StepCell):
    """Decoder cell: primitives reading state 0 (the coarser input) are followed by a x2 bilinear
    upsample — model_augment.py:64-106."""

    def __init__(self, upsample, upsample_concat, C_prev_prev, C_prev):
        super().__init__()
        self.preprocess0 = ReLUConvBN(C_prev_prev, C_prev // 4, 1, 1, 0, affine=True)
        self.preprocess1 = ReLUConvBN(C_prev, C_prev // 4, 1, 1, 0, affine=True)
        self._concat = upsample_concat
        self.multiplier = len(upsample_concat)
        self._build_ops(C_prev // 4, upsample,
                        wrap=lambda op, index: Sequential(op, Interpolate(scale_factor=2)) if index == 0 else op)

    def forward(self, s0, s1, out_raw=True, out_relu=False):
        return self._run_fused(_preprocess((self.preprocess0, self.preprocess1), (s0, s1)),
                               [list(self._concat)], out_raw, out_relu)[0]


class _FusionCell(_StepCell):
    """PoseCell1 / ParCell1 — model_augment.py:119-229: three preprocessed inputs, four steps, returns
    (cat(states[0:3]), cat(states[concat]))."""

    def __init__(self, edges, concat, C_prev_prev, C_prev, C_cur, order):
        super().__init__()
        self.order = order
        if order == 0:
            cins = (C_prev_prev, C_prev, C_cur)
        else:
            cins = (3 * C_cur, 4 * C_cur, 4 * C_cur)
        self.preprocess0 = ReLUConvBN(cins[0], C_cur, 1, 1, 0, affine=True)
        self.preprocess1 = ReLUConvBN(cins[1], C_cur, 1, 1, 0, affine=True)
        self.preprocess2 = ReLUConvBN(cins[2], C_cur, 1, 1, 0, affine=True)
        self._concat = concat
        self.multiplier = len(concat)

        def wrap(op, index):
            if order == 0 and index == 0:
                return Sequential(op, Interpolate(scale_factor=4))
            if order == 0 and index == 1:
                return Sequential(op, Interpolate(scale_factor=2))
            return op

        self._build_ops(C_cur, edges, wrap=wrap)

    def forward(self, s0, s1, s2, out_raw=True, out_relu=False, n_out=None):
        """n_out = (k1, k2): k1 handles on fea1 and k2 on fea2, one per consumer (see _run_fused)."""
        if self.order == 0:  # model_augment.py:164-166: default-mode (nearest) F.interpolate; unused by Network
            states = self._run_steps([self.preprocess0(s0), self.preprocess1(s1), self.preprocess2(s2)])
            states[0] = F_.interpolate(states[0], scale_factor=4)
            states[1] = F_.interpolate(states[1], scale_factor=2)
            return F_.cat(states[0:3]), F_.cat([states[i] for i in self._concat])
        fea1, fea2 = self._run_fused(_preprocess((self.preprocess0, self.preprocess1, self.preprocess2),
                                                  (s0, s1, s2)), [[0, 1, 2], list(self._concat)],
                                     out_raw, out_relu, list(n_out) if n_out else None)
        return fea1, fea2


class PoseCell1(_FusionCell):
    def __init__(self, pose, pose_concat, C_prev_prev, C_prev, C_cur, order):
        super().__init__(pose, pose_concat, C_prev_prev, C_prev, C_cur, order)


class ParCell1(_FusionCell):
    def __init__(self, par, par_concat, C_prev_prev, C_prev, C_cur, order):
        super().__init__(par, par_concat, C_prev_prev, C_prev, C_cur, order)


def _stem(cin, cout, stride, relu):
    layers = [Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False), BatchNorm2d(cout, momentum=BN_MOMENTUM)]
    if relu:
        layers.append(ReLU(inplace=True))
    return Sequential(*layers)


def _layer_1x1(cin, cout):
    """ReLU -> Conv1x1(bias) -> BN — model_augment.py:332-351."""
    return Sequential(ReLU(), Conv2d(cin, cout, kernel_size=1, padding=0, dilation=1),
                      BatchNorm2d(cout, momentum=BN_MOMENTUM))


def _head(cin, mid, cout, k, first_bias=True):
    """ReLU -> Conv(k) -> BN -> ReLU -> Conv1x1(bias) — model_augment.py:371-398."""
    return Sequential(ReLU(), Conv2d(cin, mid, kernel_size=k, padding=k // 2, dilation=1, bias=first_bias),
                      BatchNorm2d(mid, momentum=BN_MOMENTUM), ReLU(inplace=True),
                      Conv2d(mid, cout, kernel_size=1, padding=0, dilation=1, bias=True))


class Network(nn.Module):
    """model_augment.py:231-709."""

    def __init__(self, cfg, steps=4, multiplier=4, stem_multiplier=4):
        super().__init__()
        self._num_classes = cfg.DATASET.NUM_CLASSES
        self._num_joints = cfg.DATASET.NUM_JOINTS
        self._layers = cfg.TRAIN.LAYERS
        self.C = cfg.TRAIN.INIT_CHANNELS
        self.deconv_with_bias = cfg.MODEL.DECONV_WITH_BIAS
        self._head = cfg.MODEL.HEAD
        self.refine_layers = cfg.MODEL.REFINE_LAYERS
        C, L = self.C, self._layers

        # two independent stems, one per task stream (model_augment.py:244-272)
        self.stem0 = _stem(3, C, 2, True)
        self.stem1 = _stem(C, 2 * C, 2, True)
        self.stem2 = _stem(2 * C, 2 * C, 1, False)
        self.stem3 = _stem(3, C, 2, True)
        self.stem4 = _stem(C, 2 * C, 2, True)
        self.stem5 = _stem(2 * C, 2 * C, 1, False)

        # encoder: L cells per stream, channel doubling + reduction at L/4, L/2, 3L/4 (:274-295)
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
        self.num_inchannels = widths[::-1]  # coarse -> fine, as in the reference (:297)

        # cross-task interaction ops in the encoder (:299-306) and decoder (:308-317)
        self._indices1, ops = self._compile(gt.INTER.task1, widths)
        self._ops1 = nn.ModuleList(ops)
        self._indices2, ops = self._compile(gt.INTER.task2, widths)
        self._ops2 = nn.ModuleList(ops)
        resolution = [1, 1 / 2, 1 / 4, 1 / 8, 1 / 4, 1 / 2, 1]
        channels = [int(2 * C / r) for r in resolution]
        self.up_indices1, ops = self._compile3(gt.INTER.task3, resolution, channels)
        self.up_ops1 = nn.ModuleList(ops)
        self.up_indices2, ops = self._compile3(gt.INTER.task4, resolution, channels)
        self.up_ops2 = nn.ModuleList(ops)

        # decoder (:319-330)
        nin = self.num_inchannels
        self.upsamples1 = nn.ModuleList()
        self.upsamples2 = nn.ModuleList()
        for j in range(len(nin) - 1):
            self.upsamples1.append(Upsample(gt.DECODER.upsample1, gt.DECODER.upsample_concat1, nin[j], nin[j + 1]))
        for j in range(len(nin) - 1):
            self.upsamples2.append(Upsample(gt.DECODER.upsample2, gt.DECODER.upsample_concat2, nin[j], nin[j + 1]))

        # 1x1 projections of the 8*C3-channel multi-scale concat (:332-351)
        C3 = nin[3]
        self.pose_layer = _layer_1x1(8 * C3, 4 * C3)
        self.pose_auxlayer = _layer_1x1(8 * C3, 3 * C3)
        self.par_layer = _layer_1x1(8 * C3, 4 * C3)
        self.edge_layer = _layer_1x1(8 * C3, 3 * C3)

        # refinement cells (:354-364)
        self.pose_net = nn.ModuleList()
        self.par_net = nn.ModuleList()
        for _ in range(3):
            self.pose_net.append(PoseCell1(gt.FUSION.pose, gt.FUSION.pose_concat, C3, C3, C3, 1))
            self.par_net.append(ParCell1(gt.FUSION.par, gt.FUSION.par_concat, C3, C3, C3, 1))

        # heads, one set per refinement stage (:366-398)
        self.pose_head = nn.ModuleList()
        self.pose_auxnet = nn.ModuleList()
        self.par_head = nn.ModuleList()
        self.edge_head = nn.ModuleList()
        for _ in range(self.refine_layers + 1):
            self.pose_head.append(_head(4 * C3, 256, self._num_joints, 1))
            self.pose_auxnet.append(_head(3 * C3, 128, self._num_joints, 3))
            self.par_head.append(_head(4 * C3, 256, self._num_classes, 1))
            self.edge_head.append(_head(3 * C3, 6, 2, 3, first_bias=False))

        self._init_params()

    # ------------------------------------------------------------------ construction helpers
    @staticmethod
    def _interaction_op(name, c_src, c_dst, scale, same):
        op = OPS[name](c_src, 1, True)
        if not same:
            op = Sequential(op, _RescaleProject(Interpolate(scale), Conv2d(c_src, c_dst, 1)))
        return op

    def _compile(self, geno, C_list):
        """Encoder interaction ops: target scale `cont`, source scale `ind` (:576-598)."""
        indices, ops = [], []
        for cont, edges in enumerate(geno):
            names, idx = zip(*edges)
            indices.append(idx)
            for n, ind in zip(names, idx):
                ops.append(self._interaction_op(n, C_list[ind], C_list[cont], 1 / 2 ** (cont - ind), ind == cont))
        return indices, ops

    def _compile3(self, geno, resolutions, C_list):
        """Decoder interaction ops over the 7-entry feature list (:626-649)."""
        indices, ops = [], []
        for cont, edges in enumerate(geno):
            names, idx = zip(*edges)
            indices.append(idx)
            for n, ind in zip(names, idx):
                ops.append(self._interaction_op(n, C_list[ind], C_list[4 + cont],
                                                resolutions[4 + cont] / resolutions[ind], ind == 4 + cont))
        return indices, ops

    def _init_params(self):
        """xavier_normal on every conv weight, zero conv bias, BN weight 1 / bias 0 (:651-671); iterates
        self.modules() in the same order as the reference so equal seeds give equal weights."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, BatchNorm2d):
                if m.affine:
                    m.weight.data.fill_(1)
                    m.bias.data.zero_()

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _exchange(ops, cursor, idx, feats):
        """sum_j ops[cursor+j](feats[idx[j]]) — the cross-task message (:429-436)."""
        z = None
        for j, src in enumerate(idx):
            y = call_lazy(ops[cursor + j], feats[src])
            z = y if z is None else F_.add(z, y)
        return z, cursor + len(idx)

    # optional per-stage trace: `net._trace = fn` makes forward call fn(name, NCHW fp32 tensor) at the stage boundaries
    # the oracle records under the same names (oracle/nppnet_ref.py set_trace; tools/parity_trace.py).  Off by default.
    _trace = None

    def _tr(self, name, t, relu=False):
        if self._trace is not None:
            with torch.no_grad():
                h = F_.relu(t) if relu else F_.to_internal(t)
                self._trace(name, F_.from_internal(h))

    def forward(self, x):
        x = F_.to_internal(x)
        # The pose stream (stem0-2, cells1, upsamples1) runs on the current CUDA stream, the parsing stream (stem3-5,
        # cells2, upsamples2) on a side stream when F_._state["two_streams"] is set: they only meet at the interaction
        # points, where the side stream is joined, the cross-task messages are computed and the side stream is forked
        # again.  ts.fork / ts.join also mark the handles that cross for the caching allocator.
        ts = F_.TaskStreams(x, enabled=self._trace is None)
        if ts.on and x.dtype == torch.bfloat16 and x.shape[1] == 8 and not x.requires_grad and F_._state.get("stem_im2col", True):
            F_.im2col3x3_c3(x, 2, 1)        # the tap-folded image both stems read: computed once, before the fork
        ts.fork(x)
        s0 = self.stem1(self.stem0(x))
        s1 = self.stem2(s0)
        with ts.side():
            s2 = self.stem4(self.stem3(x))
            s3 = self.stem5(s2)
        self._tr("stem2", s1)
        self._tr("stem5", s3)
        f1, f2 = [], []           # per-stream feature pyramids (fine -> coarse, then decoder outputs)
        c1 = c2 = stage = 0
        n1 = n3 = None            # second handle on the previous cell output (read as `s0` by the cell after next)
        for i, (cell1, cell2) in enumerate(zip(self.cells1, self.cells2)):
            # every encoder cell feeds the next two cells' preprocess layers (nn.ReLU first); only the tapped ones
            # are also read raw (interaction ops, decoder, multi-scale concat).  A cell output that is read by exactly
            # those two consumers is handed out as two handles (n_out=2): their gradients are then summed inside the
            # backward kernels of the member nodes instead of by an autograd add over the whole 4C-channel tensor.
            tap = i in self._tap_layers
            two = (not tap) and i + 2 < len(self.cells1) and self._trace is None
            o1 = cell1(s0, s1, out_raw=tap, out_relu=True, n_out=2 if two else None)
            with ts.side():
                o3 = cell2(s2, s3, out_raw=tap, out_relu=True, n_out=2 if two else None)
            s0, s2 = (n1 if n1 is not None else s1), (n3 if n3 is not None else s3)
            (s1, n1), (s3, n3) = (o1, o3) if two else ((o1, None), (o3, None))
            self._tr("relu(cells1.%d)" % i, s1, relu=True)
            self._tr("relu(cells2.%d)" % i, s3, relu=True)
            if i in self._tap_layers:
                # interaction point (:422-446): each stream's message is computed from the OTHER stream's features,
                # the pose stream's on the main stream, the parsing stream's on the side stream, after a barrier that
                # hands the feature pyramids across (f1 still holds the pre-interaction state when z2 is formed)
                f1.append(s1)
                f2.append(s3)
                ts.join(f2)
                ts.fork(f1)
                z1, c1 = self._exchange(self._ops1, c1, self._indices1[stage], f2)
                n1 = F_.node(s1, z1, want_raw=True, want_relu=True)[0]  # read raw (f1) and through nn.ReLU (cells)
                with ts.side():
                    z2, c2 = self._exchange(self._ops2, c2, self._indices2[stage], f1)
                    n3 = F_.node(s3, z2, want_raw=True, want_relu=True)[0]
                s1, s3 = n1, n3
                self._tr("f1.%d" % stage, s1)
                self._tr("f2.%d" % stage, s3)
                stage += 1
                f1[-1], f2[-1] = s1, s3

        # decoder: three upsample cells per stream with interaction after each (:453-533)
        c1 = c2 = 0
        prev1, prev2 = f1[3], f2[3]
        for d in range(3):
            o1 = self.upsamples1[d](prev1, f1[2 - d])
            with ts.side():
                o2 = self.upsamples2[d](prev2, f2[2 - d])
            f1.append(o1)
            f2.append(o2)
            ts.join(f2)
            ts.fork(f1)
            z1, c1 = self._exchange(self.up_ops1, c1, self.up_indices1[d], f2)
            n1 = F_.node(o1, z1, want_raw=True, want_relu=True)[0]
            with ts.side():
                z2, c2 = self._exchange(self.up_ops2, c2, self.up_indices2[d], f1)
                n2 = F_.node(o2, z2, want_raw=True, want_relu=True)[0]
            o1, o2 = n1, n2
            self._tr("f1.%d" % (4 + d), o1)
            self._tr("f2.%d" % (4 + d), o2)
            f1[-1], f2[-1] = o1, o2
            prev1, prev2 = o1, o2

        def pyramid(f):  # (:538-543); only read through the nn.ReLU of the four 1x1 layers: relu(cat) is written directly
            return F_.cat_relu([f[0], f[6],
                                F_.interpolate(f[5], scale_factor=2, mode="bilinear", align_corners=True),
                                F_.interpolate(f[4], scale_factor=4, mode="bilinear", align_corners=True)])

        # Pending BatchNorm outputs: the heads and refinement cells all start with nn.ReLU, so relu(bn(.)) is
        # produced once per tensor and the raw values are never materialised.  The pose half stays on the main stream,
        # the parsing half (pyramid of f2, edge / parsing layers, heads and ParCell) runs on the side stream.
        split = self.refine_layers == 1 and self._trace is None
        x1 = pyramid(f1)
        in1 = self.pose_auxlayer.lazy(x1)
        in3 = self.pose_layer.lazy(x1)
        if split:   # relu(bn(.)) of the four layer outputs is read by a head and by one / two refinement cells each
            F_.set_fanout(in1, 2), F_.set_fanout(in3, 3)
        if ts.on:
            F_.materialize(in1), F_.materialize(in3)
        with ts.side():
            x2 = pyramid(f2)
            in2 = self.edge_layer.lazy(x2)
            in4 = self.par_layer.lazy(x2)
            if split:
                F_.set_fanout(in2, 2), F_.set_fanout(in4, 3)
            if ts.on:
                F_.materialize(in2), F_.materialize(in4)
        for nm, t in (("relu(pose_auxlayer)", in1), ("relu(edge_layer)", in2), ("relu(pose_layer)", in3),
                      ("relu(par_layer)", in4)):
            self._tr(nm, t, relu=True)

        pose_list, par_list = [], []
        side_outputs = []

        def emit(stage_idx, p1, p2, p3, p4):
            pose_aux = self.pose_auxnet[stage_idx](p1)
            pose_map = self.pose_head[stage_idx](p3)
            pose_list.append([F_.from_internal(pose_map, self._num_joints), F_.from_internal(pose_aux, self._num_joints)])
            with ts.side():
                edge = self.edge_head[stage_idx](p2)
                par_map = self.par_head[stage_idx](p4)
                par_list.append([F_.from_internal(par_map, self._num_classes), F_.from_internal(edge, 2)])
            side_outputs.extend(par_list[-1])

        emit(0, in1, in2, in3, in4)
        in3b, in4b = in3, in4     # the handles the parsing cell reads (own ones when the outputs were split per consumer)
        for i in range(1, self.refine_layers + 1):
            for j in range(3):
                # refinement-cell outputs are read by preprocess layers and heads only (nn.ReLU first).  fea2 of either
                # cell feeds BOTH cells of the next step: two handles (n_out), so the two gradients of these 4C-channel
                # tensors are summed inside the member nodes' backward kernels, not by an autograd add.
                # PoseCell on the main stream, ParCell on the side stream; each reads the other's fea2: barrier first.
                ts.join(in4)
                ts.fork(in3b)
                n_out = (1, 2) if (split and j < 2) else None
                o1, t3 = self.pose_net[2 * (i - 1) + j](in1, in3, in4, out_raw=False, out_relu=True, n_out=n_out)
                with ts.side():
                    o2, t4 = self.par_net[2 * (i - 1) + j](in2, in3b, in4b, out_raw=False, out_relu=True, n_out=n_out)
                if n_out:
                    in1, in2, (in3, in3b), (in4b, in4) = o1[0], o2[0], t3, t4
                else:
                    in1, in2, in3, in4 = o1, o2, t3, t4
                    in3b, in4b = in3, in4
                k = 2 * (i - 1) + j
                for nm, t in (("relu(pose_net.%d.fea1)" % k, in1), ("relu(pose_net.%d.fea2)" % k, in3),
                              ("relu(par_net.%d.fea1)" % k, in2), ("relu(par_net.%d.fea2)" % k, in4)):
                    self._tr(nm, t, relu=True)
            emit(i, in1, in2, in3, in4b)
        ts.join(side_outputs)      # the parsing / edge logits were produced on the side stream
        return pose_list, par_list

    # ------------------------------------------------------------------ checkpoints
    def load_pretrain_backbone(self, path=""):
        """Loads a reference checkpoint: strips DDP's `module.` prefix, skips shape mismatches and
        missing keys (strict=False) — model_augment.py:673-709."""
        if not os.path.isfile(path):
            return
        loaded = torch.load(path, map_location="cpu")
        own = self.state_dict()
        merged = {}
        for k, v in loaded.items():
            k = k[7:] if k.startswith("module") else k
            if k in own and tuple(own[k].shape) != tuple(v.shape):
                print("Skip loading parameter {}, required shape{}, loaded shape{}.".format(k, own[k].shape, v.shape))
                v = own[k]
            merged[k] = v
        for k, v in own.items():
            merged.setdefault(k, v)
        msg = self.load_state_dict(merged, strict=False)
        print("=> loading information:", msg)
        print("successful load pretrained backbone from {}".format(path))
