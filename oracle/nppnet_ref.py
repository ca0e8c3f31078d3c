"""ORACLE — test infrastructure only (imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package npp_b200).

A functional, plain-PyTorch fp32 restatement of the reference's hot path: the operator primitives
(/root/reference/models/operations.py), the derived network (models/model_augment.py) and the
search supernet pieces, driven directly by a reference-format state_dict (same key names) instead
of nn.Module objects.  Each function cites the reference lines it restates.

Pinning: the reference publishes no golden vectors, so this restatement is pinned
against the reference's own modules imported from /root/reference in the build container
(tests/test_oracle_vs_reference.py, skipped where /root/reference is absent) and against the
fixtures under tests/golden/ that were generated from the reference by tests/golden/make_golden.py.
"""
import math

import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.1
BN_EPS = 1e-5

# ---- genotype data (restated from models/genotypes.py:30-54) ------------------------------------
_E = lambda s: [(t.rsplit("@", 1)[0], int(t.rsplit("@", 1)[1])) for t in s.split()]
ENCODER_NORMAL = _E("std_conv_3x3@0 se_connect@1 se_connect@1 std_conv_3x3@0 max_pool_3x3@1 std_conv_3x3@2 "
                    "std_conv_3x3@3 std_conv_3x3@0")
ENCODER_REDUCE = _E("std_conv_3x3@0 se_connect@1 se_connect@1 std_conv_3x3@2 dil_conv_3x3_4@3 dil_conv_3x3_4@2 "
                    "max_pool_3x3@3 dil_conv_3x3_2@0")
DECODER_UP1 = _E("std_conv_1x1@1 std_conv_1x1@0 std_conv_1x1@1 std_conv_3x3@0 std_conv_1x1@0 dil_conv_3x3_2@1 "
                 "std_conv_3x3@3 std_conv_1x1@1")
DECODER_UP2 = _E("std_conv_3x3@1 se_connect@0 dil_conv_3x3_2@2 std_conv_1x1@1 poled_conv_x1@3 std_conv_1x1@2 "
                 "std_conv_3x3@1 std_conv_1x1@2")
INTER_TASK1 = [_E("dil_conv_3x3_2@0"), _E("std_conv_3x3@1"), _E("std_conv_1x1@1 std_conv_3x3@2"),
               _E("std_conv_1x1@2 std_conv_3x3@3")]
INTER_TASK2 = [_E("dil_conv_3x3_2@0"), _E("poled_conv_x1@1"), _E("std_conv_1x1@2"), _E("std_conv_3x3@1 std_conv_3x3@3")]
INTER_TASK3 = [_E("dil_conv_3x3_2@4 dil_conv_3x3_2@2 dil_conv_3x3_2@1"),
               _E("std_conv_3x3@1 std_conv_3x3@2 dil_conv_3x3_2@5 dil_conv_3x3_2@0"),
               _E("std_conv_3x3@1 dil_conv_3x3_2@2 dil_conv_3x3_4@5 dil_conv_3x3_2@3")]
INTER_TASK4 = [_E("std_conv_3x3@0"), _E("std_conv_3x3@1"), _E("std_conv_1x1@2 std_conv_3x3@1")]
FUSION_POSE = _E("std_conv_3x3@1 std_conv_3x3@2 std_conv_3x3@0 max_pool_3x3@2 std_conv_3x3@4 std_conv_3x3@2 "
                 "std_conv_3x3@4 std_conv_3x3@3")
FUSION_PAR = _E("dil_conv_3x3_2@2 se_connect@1 dil_conv_3x3_2@2 dil_conv_3x3_2@3 max_pool_3x3@3 std_conv_3x3@2 "
                "dil_conv_3x3_2@5 std_conv_3x3@2")


# ---- optional storage-precision emulation -----------------------------------------------------------
# The product path stores every activation (and activation gradient) as bf16 and does its arithmetic in
# fp32.  set_storage_dtype(torch.bfloat16) makes this oracle round the output of every operator (and the
# gradient flowing back through it) the same way, which gives the error level that is inherent to bf16
# storage for a given network/input — the yardstick the bf16 parity tests compare against.
_STORAGE = [None, False]


def set_storage_dtype(dtype, weights=False):
    """dtype: round every operator output (and its gradient) to this dtype; weights=True: also round the dense /
    depthwise convolution weights on use (what a bf16 tensor-core GEMM — or torch autocast — feeds the multiplier)."""
    _STORAGE[0] = dtype
    _STORAGE[1] = bool(weights) and dtype is not None


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(_STORAGE[0]).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(_STORAGE[0]).to(g.dtype)


def rnd(x):
    return x if _STORAGE[0] is None else _Round.apply(x)


# ---- optional per-stage trace (tools/parity_trace.py, tests/test_gpu_baseline_config.py) ---------------------------
# set_trace(dict) makes network_forward record named intermediate tensors (detached) so that a parity failure can be
# located to a stage instead of being judged at the outputs only.
_TRACE = [None]


def set_trace(store):
    _TRACE[0] = store


def _tr(name, t):
    if _TRACE[0] is not None:
        _TRACE[0][name] = t.detach()
    return t


class Params:
    """state_dict accessor with a key prefix; `training` selects batch vs running statistics."""

    def __init__(self, sd, training=True, prefix=""):
        self.sd, self.training, self.prefix = sd, training, prefix

    def sub(self, name):
        return Params(self.sd, self.training, self.prefix + str(name) + ".")

    def get(self, name):
        return self.sd.get(self.prefix + name)

    def __getitem__(self, name):
        return self.sd[self.prefix + name]


def bn(p, x):
    """nn.BatchNorm2d(momentum=0.1): batch statistics + running update in training mode."""
    return rnd(F.batch_norm(x, p["running_mean"], p["running_var"], p.get("weight"), p.get("bias"), p.training,
                          BN_MOMENTUM, BN_EPS))


def conv(p, x, stride=1, padding=0, dilation=1, groups=1):
    w = p["weight"]
    if _STORAGE[1] and groups == 1:       # dense convs run on the tensor cores with bf16 operands
        w = w + (w.to(_STORAGE[0]).to(w.dtype) - w).detach()      # straight-through rounding
    return rnd(F.conv2d(x, w, p.get("bias"), stride, padding, dilation, groups))


def relu_conv_bn(p, x, k, stride, pad, dil=1):
    """operations.py:69-82 ReLUConvBN / :85-101 DilConv: net.0 ReLU, net.1 Conv, net.2 BN."""
    n = p.sub("net")
    return bn(n.sub(2), conv(n.sub(1), F.relu(x), stride, pad, dil))


def dil_conv_s(p, x, k, stride, pad, dil):
    """operations.py:202-220: ReLU, depthwise (net.1), pointwise (net.2), BN (net.3)."""
    n = p.sub("net")
    y = conv(n.sub(1), F.relu(x), stride, pad, dil, groups=x.shape[1])
    return bn(n.sub(3), conv(n.sub(2), y))


def sep_conv(p, x, k, stride, pad):
    """operations.py:190-200."""
    n = p.sub("net")
    return dil_conv_s(n.sub(1), dil_conv_s(n.sub(0), x, k, stride, pad, 1), k, 1, pad, 1)


def pool_bn(p, x, kind, stride):
    """operations.py:44-66."""
    if kind == "max":
        y = F.max_pool2d(x, 3, stride, 1)
    else:
        y = rnd(F.avg_pool2d(x, 3, stride, 1, count_include_pad=False))
    return bn(p.sub("bn"), y)


def se_block(p, x, stride):
    """operations.py:105-129."""
    w = F.adaptive_avg_pool2d(x, 1)
    w = F.relu(conv(p.sub("conv1"), w))
    w = torch.sigmoid(conv(p.sub("conv2"), w))
    out = rnd(x * w)
    if stride == 1:
        return out
    return bn(p.sub("bn"), rnd(F.avg_pool2d(out, 2)))


def factorized_reduce(p, x):
    """operations.py:142-157."""
    x = F.relu(x)
    out = torch.cat([conv(p.sub("conv1"), x, 2), conv(p.sub("conv2"), x[:, :, 1:, 1:], 2)], dim=1)
    return bn(p.sub("bn"), out)


def pooled_conv(p, x, stride, conv_nums):
    """operations.py:222-251: net.0 AvgPool, then (ReLU, Conv(bias), BN) x n at net.{1+3i..3+3i}, upsample."""
    n = p.sub("net")
    y = rnd(F.avg_pool2d(x, 2, 2))
    for i in range(conv_nums):
        y = bn(n.sub(3 + 3 * i), conv(n.sub(2 + 3 * i), F.relu(y), stride, 1))
    y = up(y, 2)
    if conv_nums == 2 and stride == 2:
        y = up(y, 2)
    return y


def primitive(name, p, x, stride):
    """OPS[name](C, stride, affine).forward(x) — operations.py:9-25."""
    if name == "none":
        return x * 0. if stride == 1 else x[:, :, ::stride, ::stride] * 0
    if name == "avg_pool_3x3":
        return pool_bn(p, x, "avg", stride)
    if name == "max_pool_3x3":
        return pool_bn(p, x, "max", stride)
    if name == "skip_connect":
        return x if stride == 1 else factorized_reduce(p, x)
    if name == "std_conv_3x3":
        return relu_conv_bn(p, x, 3, stride, 1)
    if name == "std_conv_1x1":
        return relu_conv_bn(p, x, 1, stride, 0)
    if name == "dil_conv_3x3_2":
        return dil_conv_s(p, x, 3, stride, 2, 2)
    if name == "dil_conv_3x3_4":
        return dil_conv_s(p, x, 3, stride, 4, 4)
    if name == "dil_conv_5x5_4":
        return dil_conv_s(p, x, 5, stride, 4, 2)
    if name == "se_connect":
        return se_block(p, x, stride)
    if name == "sep_conv_3x3":
        return sep_conv(p, x, 3, stride, 1)
    if name == "sep_conv_5x5":
        return sep_conv(p, x, 5, stride, 2)
    if name == "poled_conv_x1":
        return pooled_conv(p, x, stride, 1)
    if name == "poled_conv_x2":
        return pooled_conv(p, x, stride, 2)
    raise KeyError(name)


def up(x, s):
    return rnd(F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=True))


# ---- cells (model_augment.py:16-229) -------------------------------------------------------------
def _steps(p, edges, states, stride_of=lambda idx: 1, post=lambda y, idx, k: y, wrapped=lambda idx: False):
    for i in range(len(edges) // 2):
        hs = []
        for k in (2 * i, 2 * i + 1):
            name, idx = edges[k]
            q = p.sub("_ops").sub(k)
            if wrapped(idx):
                q = q.sub(0)  # nn.Sequential(op, Interpolate)
            hs.append(post(primitive(name, q, states[idx], stride_of(idx)), idx, k))
        states.append(rnd(hs[0] + hs[1]))
    return states


def encoder_cell(p, s0, s1, reduction, reduction_prev):
    """model_augment.py:16-62."""
    s0 = factorized_reduce(p.sub("preprocess0"), s0) if reduction_prev else relu_conv_bn(p.sub("preprocess0"), s0, 1, 1, 0)
    s1 = relu_conv_bn(p.sub("preprocess1"), s1, 1, 1, 0)
    edges = ENCODER_REDUCE if reduction else ENCODER_NORMAL
    st = _steps(p, edges, [s0, s1], stride_of=lambda idx: 2 if reduction and idx < 2 else 1)
    return torch.cat(st[2:6], dim=1)


def upsample_cell(p, s0, s1, edges):
    """model_augment.py:64-106: ops on state 0 are followed by a x2 bilinear upsample."""
    s0 = relu_conv_bn(p.sub("preprocess0"), s0, 1, 1, 0)
    s1 = relu_conv_bn(p.sub("preprocess1"), s1, 1, 1, 0)
    st = _steps(p, edges, [s0, s1], post=lambda y, idx, k: up(y, 2) if idx == 0 else y, wrapped=lambda idx: idx == 0)
    return torch.cat(st[2:6], dim=1)


def fusion_cell(p, s0, s1, s2, edges):
    """model_augment.py:119-229 with order == 1 (the only order Network builds, :357-363)."""
    st = [relu_conv_bn(p.sub("preprocess%d" % i), s, 1, 1, 0) for i, s in enumerate((s0, s1, s2))]
    st = _steps(p, edges, st)
    return torch.cat(st[0:3], dim=1), torch.cat(st[3:7], dim=1)


def _seq_conv_bn(p, x, conv_idx, bn_idx, pad=0, stride=1, relu_in=False, relu_out=False):
    if relu_in:
        x = F.relu(x)
    y = bn(p.sub(bn_idx), conv(p.sub(conv_idx), x, stride, pad))
    return F.relu(y) if relu_out else y


def head(p, x, k):
    """model_augment.py:371-398: ReLU, Conv(k), BN, ReLU, Conv1x1."""
    y = _seq_conv_bn(p, x, 1, 2, pad=k // 2, relu_in=True, relu_out=True)
    return conv(p.sub(4), y)


def _interaction(p, list_name, cursor, edges_for_target, feats, scale_of, same_of):
    z = 0
    for j, (name, ind) in enumerate(edges_for_target):
        q = p.sub(list_name).sub(cursor + j)
        if same_of(ind):
            y = primitive(name, q, feats[ind], 1)
        else:
            y = primitive(name, q.sub(0), feats[ind], 1)
            y = conv(q.sub(1).sub(1), up(y, scale_of(ind)))
        z = y if isinstance(z, int) else rnd(z + y)
    return z, cursor + len(edges_for_target)


def network_forward(sd, x, layers=16, refine_layers=1, training=True):
    """models/model_augment.py:402-574 Network.forward on a reference-format state_dict."""
    p = Params(sd, training)
    L = layers
    taps = [L // 4 - 1, 2 * L // 4 - 1, 3 * L // 4 - 1, 4 * L // 4 - 1]
    reduces = [L // 4, 2 * L // 4, 3 * L // 4]

    def stem(name, x, stride, relu_out):
        return _seq_conv_bn(p.sub(name), x, 0, 1, pad=1, stride=stride, relu_out=relu_out)

    s0 = stem("stem1", stem("stem0", x, 2, True), 2, True)
    s1 = _tr("stem2", stem("stem2", s0, 1, False))
    s2 = stem("stem4", stem("stem3", x, 2, True), 2, True)
    s3 = _tr("stem5", stem("stem5", s2, 1, False))
    f1, f2 = [], []
    c1 = c2 = stage = 0
    red_prev = False
    for i in range(L):
        red = i in reduces
        if _TRACE[0] is not None:      # stage inputs, for stage-wise (teacher-forced) parity checks
            _tr("cells1.%d.in0" % i, s0), _tr("cells1.%d.in1" % i, s1)
            _tr("cells2.%d.in0" % i, s2), _tr("cells2.%d.in1" % i, s3)
        s0, s1 = s1, encoder_cell(p.sub("cells1").sub(i), s0, s1, red, red_prev)
        s2, s3 = s3, encoder_cell(p.sub("cells2").sub(i), s2, s3, red, red_prev)
        if _TRACE[0] is not None:
            _tr("relu(cells1.%d)" % i, F.relu(s1))
            _tr("relu(cells2.%d)" % i, F.relu(s3))
        red_prev = red
        if i in taps:
            f1.append(s1)
            f2.append(s3)
            sc = lambda ind, cont=stage: 1 / 2 ** (cont - ind)
            same = lambda ind, cont=stage: ind == cont
            z1, c1 = _interaction(p, "_ops1", c1, INTER_TASK1[stage], f2, sc, same)
            z2, c2 = _interaction(p, "_ops2", c2, INTER_TASK2[stage], f1, sc, same)
            s1 = _tr("f1.%d" % stage, rnd(s1 + z1))
            s3 = _tr("f2.%d" % stage, rnd(s3 + z2))
            stage += 1
            f1[-1], f2[-1] = s1, s3

    res = [1, 1 / 2, 1 / 4, 1 / 8, 1 / 4, 1 / 2, 1]
    c1 = c2 = 0
    prev1, prev2 = f1[3], f2[3]
    for d in range(3):
        if _TRACE[0] is not None:
            _tr("upsamples1.%d.in0" % d, prev1), _tr("upsamples1.%d.in1" % d, f1[2 - d])
            _tr("upsamples2.%d.in0" % d, prev2), _tr("upsamples2.%d.in1" % d, f2[2 - d])
        o1 = upsample_cell(p.sub("upsamples1").sub(d), prev1, f1[2 - d], DECODER_UP1)
        o2 = upsample_cell(p.sub("upsamples2").sub(d), prev2, f2[2 - d], DECODER_UP2)
        f1.append(o1)
        f2.append(o2)
        sc = lambda ind, d=d: res[4 + d] / res[ind]
        same = lambda ind, d=d: ind == 4 + d
        z1, c1 = _interaction(p, "up_ops1", c1, INTER_TASK3[d], f2, sc, same)
        z2, c2 = _interaction(p, "up_ops2", c2, INTER_TASK4[d], f1, sc, same)
        o1, o2 = _tr("f1.%d" % (4 + d), rnd(o1 + z1)), _tr("f2.%d" % (4 + d), rnd(o2 + z2))
        f1[-1], f2[-1] = o1, o2
        prev1, prev2 = o1, o2

    x1 = torch.cat((f1[0], f1[6], up(f1[5], 2), up(f1[4], 4)), dim=1)
    x2 = torch.cat((f2[0], f2[6], up(f2[5], 2), up(f2[4], 4)), dim=1)
    _tr("x1", x1), _tr("x2", x2)
    in1 = _seq_conv_bn(p.sub("pose_auxlayer"), x1, 1, 2, relu_in=True)
    in2 = _seq_conv_bn(p.sub("edge_layer"), x2, 1, 2, relu_in=True)
    in3 = _seq_conv_bn(p.sub("pose_layer"), x1, 1, 2, relu_in=True)
    in4 = _seq_conv_bn(p.sub("par_layer"), x2, 1, 2, relu_in=True)
    if _TRACE[0] is not None:
        for nm, t in (("relu(pose_auxlayer)", in1), ("relu(edge_layer)", in2), ("relu(pose_layer)", in3),
                      ("relu(par_layer)", in4)):
            _tr(nm, F.relu(t))
    pose_list, par_list = [], []

    def emit(i):
        if _TRACE[0] is not None:
            for q, t in enumerate((in1, in2, in3, in4)):
                _tr("head_in.%d.%d" % (i, q), t)
        edge = head(p.sub("edge_head").sub(i), in2, 3)
        pose_aux = head(p.sub("pose_auxnet").sub(i), in1, 3)
        pose_map = head(p.sub("pose_head").sub(i), in3, 1)
        par_map = head(p.sub("par_head").sub(i), in4, 1)
        pose_list.append([pose_map, pose_aux])
        par_list.append([par_map, edge])

    emit(0)
    for i in range(1, refine_layers + 1):
        for j in range(3):
            k = 2 * (i - 1) + j
            if _TRACE[0] is not None:
                _tr("pose_net.%d.in0" % k, in1), _tr("pose_net.%d.in1" % k, in3), _tr("pose_net.%d.in2" % k, in4)
                _tr("par_net.%d.in0" % k, in2), _tr("par_net.%d.in1" % k, in3), _tr("par_net.%d.in2" % k, in4)
            in1, tmp = fusion_cell(p.sub("pose_net").sub(k), in1, in3, in4, FUSION_POSE)
            in2, in4 = fusion_cell(p.sub("par_net").sub(k), in2, in3, in4, FUSION_PAR)
            in3 = tmp
            if _TRACE[0] is not None:
                for nm, t in (("relu(pose_net.%d.fea1)" % k, in1), ("relu(pose_net.%d.fea2)" % k, in3),
                              ("relu(par_net.%d.fea1)" % k, in2), ("relu(par_net.%d.fea2)" % k, in4)):
                    _tr(nm, F.relu(t))
        emit(i)
    return pose_list, par_list


# ---- search supernet (models/model_search_interact.py) ----------------------------------------------
PRIMITIVES_INTER = ["std_conv_3x3", "dil_conv_3x3_4", "se_connect", "max_pool_3x3", "dil_conv_3x3_2", "std_conv_1x1",
                    "poled_conv_x1"]  # genotypes.py:20-28


def mixed_op(p, x, weights, up_scale=None, has_extra=False):
    """MixedOp.forward, model_search_interact.py:56-74 (stride 1): candidates at C/2 with affine=False BatchNorm,
    pooling gets a second BatchNorm (:48-49), every candidate is followed by Interpolate(up_scale) when rescaling
    (:50-51, bilinear align_corners=True) while the pass-through half is resampled in nearest mode (:63-64);
    cat + channel_shuffle(groups=2) (:70-71); optional 1x1 extra_conv."""
    c = x.shape[1]
    lo, hi = x[:, :c // 2], x[:, c // 2:]
    total = 0
    for k, name in enumerate(PRIMITIVES_INTER):
        q = p.sub("_ops").sub(k)
        if up_scale:
            q = q.sub(0)
        if "pool" in name:
            y = bn(q.sub(1), primitive(name, q.sub(0), lo, 1))
        else:
            y = primitive(name, q, lo, 1)
        if up_scale:
            y = up(y, up_scale)
        total = total + weights[k] * y
    total = rnd(total)
    if up_scale:
        hi = F.interpolate(hi, scale_factor=up_scale)
    ans = torch.cat([total, hi], dim=1)
    b, ch, h, w = ans.shape
    ans = ans.view(b, 2, ch // 2, h, w).transpose(1, 2).contiguous().view(b, ch, h, w)
    if has_extra:
        ans = conv(p.sub("extra_conv"), ans)
    return ans


def beta_weights(n_input, steps, betas):
    """Network.btw, model_search_interact.py:1054-1065."""
    out, start = [], 0
    for i in range(steps):
        out.append(torch.softmax(betas[start:start + n_input + i], dim=-1))
        start += n_input + i
    return torch.cat(out)


def search_fusion_cell(p, s0, s1, s2, weights, weights2, steps=4, multiplier=4):
    """PoseCell / ParCell with order == 1, model_search_interact.py:332-429."""
    st = [relu_conv_bn(p.sub("preprocess%d" % i), s, 1, 1, 0) for i, s in enumerate((s0, s1, s2))]
    offset = 0
    for _ in range(steps):
        s = 0
        for j, h in enumerate(st):
            s = s + weights2[offset + j] * mixed_op(p.sub("_ops").sub(offset + j), h, weights[offset + j])
        offset += len(st)
        st.append(rnd(s))
    return torch.cat(st[0:3], dim=1), torch.cat(st[-multiplier:], dim=1)


def search_forward(sd, x, layers=16, refine_layers=1, training=True, steps=4):
    """models/model_search_interact.py:626-770 Network.forward on a reference-format state_dict (which holds the
    twelve architecture tensors under alphas1.. / betas1.. as well)."""
    p = Params(sd, training)
    L = layers
    taps = [L // 4 - 1, 2 * L // 4 - 1, 3 * L // 4 - 1, 4 * L // 4 - 1]
    reduces = [L // 4, 2 * L // 4, 3 * L // 4]

    def stem(name, x, stride, relu_out):
        return _seq_conv_bn(p.sub(name), x, 0, 1, pad=1, stride=stride, relu_out=relu_out)

    def interact(list_name, offset, feats, alphas, betas, scale_of, extra_of):
        wa = torch.softmax(sd[alphas][offset:offset + len(feats)], dim=-1)
        wb = torch.softmax(sd[betas][offset:offset + len(feats)], dim=-1)
        z = 0
        for j, h in enumerate(feats):
            z = z + wb[j] * mixed_op(p.sub(list_name).sub(offset + j), h, wa[j], scale_of(j), extra_of(j))
        return z

    s0 = stem("stem1", stem("stem0", x, 2, True), 2, True)
    s1 = stem("stem2", s0, 1, False)
    s2 = stem("stem4", stem("stem3", x, 2, True), 2, True)
    s3 = stem("stem5", s2, 1, False)
    f1, f2 = [], []
    offset = stage = 0
    red_prev = False
    for i in range(L):
        red = i in reduces
        s0, s1 = s1, encoder_cell(p.sub("cells1").sub(i), s0, s1, red, red_prev)
        s2, s3 = s3, encoder_cell(p.sub("cells2").sub(i), s2, s3, red, red_prev)
        red_prev = red
        if i in taps:
            f1.append(s1)
            f2.append(s3)
            sc = lambda j, i_=stage: 1 / 2 ** (i_ - j)      # :513 (1.0 on the diagonal: a truthy identity resample)
            ex = lambda j, i_=stage: j != i_
            z1 = interact("_ops1", offset, f2, "alphas1", "betas1", sc, ex)
            z2 = interact("_ops2", offset, f1, "alphas2", "betas2", sc, ex)
            s1, s3 = rnd(s1 + z1), rnd(s3 + z2)
            f1[-1], f2[-1] = s1, s3
            offset += len(f1)
            stage += 1

    res = [1, 1 / 2, 1 / 4, 1 / 8, 1 / 4, 1 / 2, 1]
    cont = 0
    prev1, prev2 = f1[3], f2[3]
    for d in range(3):
        o1 = upsample_cell(p.sub("upsamples1").sub(d), prev1, f1[2 - d], DECODER_UP1)
        o2 = upsample_cell(p.sub("upsamples2").sub(d), prev2, f2[2 - d], DECODER_UP2)
        f1.append(o1)
        f2.append(o2)
        sc = lambda j, d=d: res[4 + d] / res[j]
        ex = lambda j, d=d: j != 4 + d
        z1 = interact("up_ops1", cont, f2, "alphas3", "betas3", sc, ex)
        z2 = interact("up_ops2", cont, f1, "alphas4", "betas4", sc, ex)
        o1, o2 = rnd(o1 + z1), rnd(o2 + z2)
        f1[-1], f2[-1] = o1, o2
        prev1, prev2 = o1, o2
        cont += len(f1)

    x1 = torch.cat((f1[0], f1[6], up(f1[5], 2), up(f1[4], 4)), dim=1)
    x2 = torch.cat((f2[0], f2[6], up(f2[5], 2), up(f2[4], 4)), dim=1)
    in1 = _seq_conv_bn(p.sub("pose_auxlayer"), x1, 1, 2, relu_in=True)
    in2 = _seq_conv_bn(p.sub("edge_layer"), x2, 1, 2, relu_in=True)
    in3 = _seq_conv_bn(p.sub("pose_layer"), x1, 1, 2, relu_in=True)
    in4 = _seq_conv_bn(p.sub("par_layer"), x2, 1, 2, relu_in=True)
    pose_list, par_list = [], []

    def emit(i):
        edge = head(p.sub("edge_head").sub(i), in2, 3)
        pose_aux = head(p.sub("pose_auxnet").sub(i), in1, 3)
        pose_map = head(p.sub("pose_head").sub(i), in3, 1)
        par_map = head(p.sub("par_head").sub(i), in4, 1)
        pose_list.append([pose_map, pose_aux])
        par_list.append([par_map, edge])

    emit(0)
    w_pose = torch.softmax(sd["alphas_pose"], dim=-1)
    w_pose2 = beta_weights(3, steps, sd["betas_pose"])
    w_par = torch.softmax(sd["alphas_par"], dim=-1)
    w_par2 = beta_weights(3, steps, sd["betas_par"])
    for i in range(1, refine_layers + 1):
        for j in range(3):
            k = 2 * (i - 1) + j
            in1, tmp = search_fusion_cell(p.sub("pose_net").sub(k), in1, in3, in4, w_pose, w_pose2, steps)
            in2, in4 = search_fusion_cell(p.sub("par_net").sub(k), in2, in3, in4, w_par, w_par2, steps)
            in3 = tmp
        emit(i)
    return pose_list, par_list


def loss_entropy(alphas, n_arch=12):
    """Network.loss_entropy, model_search_interact.py:881-896: normalised row entropy of softmax(alpha)."""
    total = 0.
    for a in alphas:
        w = torch.softmax(a, dim=-1)
        ent = torch.distributions.Categorical(probs=w).entropy() / math.log(w.shape[1])
        total = total + ent.mean(dim=0)
    return 0.25 * 2 * total / n_arch


# ---- losses (core/criterion.py) ---------------------------------------------------------------------
WEIGHTS_LIP = [0.7602572, 0.94236198, 0.85644457, 1.04346266, 1.10627293, 0.80980162, 0.95168713, 0.8403769,
               1.05798412, 0.85746254, 1.01274366, 1.05854692, 1.03430773, 0.84867818, 0.88027721, 0.87580925,
               0.98747462, 0.9876475, 1.00016535, 1.00108882]  # criterion.py:17-21
WEIGHTS_PASCAL = [0.82877791, 0.95688253, 0.94921949, 1.00538108, 1.0201687, 1.01665831, 1.05470914]  # :13-14


def ohem_ce(score, target, weight, ignore_index=255, thresh=0.9, min_kept=131072):
    """criterion.py:54-72 OhemCrossEntropy.forward on already-upsampled logits."""
    pred = F.softmax(score, dim=1)
    pixel_losses = F.cross_entropy(score, target, weight=weight, ignore_index=ignore_index, reduction="none").reshape(-1)
    mask = target.reshape(-1) != ignore_index
    tmp = target.clone()
    tmp[tmp == ignore_index] = 0
    pred = pred.gather(1, tmp.unsqueeze(1)).reshape(-1)[mask]
    pred, ind = pred.sort()
    min_value = pred[min(max(1, min_kept), pred.numel() - 1)]
    threshold = max(min_value, thresh)
    pixel_losses = pixel_losses[mask][ind]
    return pixel_losses[pred < threshold].mean()


def parsing_loss(preds, target, weight, ignore_index=255, thresh=0.9, min_kept=131072):
    """criterion.py:158-202 Criterion_par.parsing_loss for preds = [par_map, edge]."""
    h, w = target[0].shape[1:]
    pos = torch.sum(target[1] == 1, dtype=torch.float)
    neg = torch.sum(target[1] == 0, dtype=torch.float)
    ew = torch.stack([pos / (pos + neg), neg / (pos + neg)]).to(preds[1].device, preds[1].dtype)
    sp = F.interpolate(preds[0], size=(h, w), mode="bilinear", align_corners=True)
    loss = ohem_ce(sp, target[0], weight, ignore_index, thresh, min_kept)
    se = F.interpolate(preds[1], size=(h, w), mode="bilinear", align_corners=True)
    return loss + F.cross_entropy(se, target[1], ew, ignore_index=ignore_index)


def criterion_par(preds, target, lamda, weight, **kw):
    """criterion.py:204-217."""
    loss = 0.
    for i in range(len(preds)):
        loss = loss + parsing_loss(preds[i], target, weight, **kw) * torch.exp(-lamda[i]) + lamda[i]
    return loss


def joint_loss(output, target):
    """criterion.py:82-128 with use_target_weight=False: (sum_j MSE_j(main) + sum_j MSE_j(aux)) / J."""
    loss = 0.
    J = output[0].shape[1]
    for o, t in zip(output, target):
        for j in range(J):
            loss = loss + F.mse_loss(o[:, j].reshape(o.shape[0], -1), t[:, j].reshape(t.shape[0], -1))
    return loss / J


def criterion_pose(output, target, lamda):
    """criterion.py:130-145."""
    loss = 0.
    for i in range(len(output)):
        loss = loss + joint_loss(output[i], target) * torch.exp(-lamda[i]) + lamda[i]
    return loss
