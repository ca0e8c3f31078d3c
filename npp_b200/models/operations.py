"""NPPNet operator primitives on libnpp_b200 kernels.

Same public surface as the reference's models/operations.py: `OPS[name](C, stride, affine)` returns
an nn.Module whose `forward(x)` maps [N, C, H, W] -> [N, C, H/stride, W/stride]; class names, child
module names/indices and parameter shapes are identical so that state_dicts are interchangeable
(SURVEY.md §5 "state_dict key names are part of the drop-in contract").  The arithmetic of each
primitive follows the cited reference lines; the execution is ours (npp_b200.nn / functional).
"""
import torch.nn as nn

from .. import functional as F_
from ..nn import AvgPool2d, BatchNorm2d, Conv2d, MaxPool2d, ReLU, Sequential, UpsamplingBilinear2d, call_lazy

# What a primitive reads from the state it is applied to (used by the cells to decide which of {state, relu(state)}
# the producing node kernel must write):  "relu" = starts with nn.ReLU feeding a dense conv (operations.py:76,95,146),
# "relu_ok" = starts with nn.ReLU feeding a depthwise conv, which can also apply the ReLU inside its own loads,
# "raw" = reads the state itself.

BN_MOMENTUM = 0.1  # operations.py:27


def _unsupported(name):
    def make(C, stride, affine):
        raise NotImplementedError(
            "primitive %r is listed in the reference's OPS table (models/operations.py:9-25) but is in none of its "
            "PRIMITIVES_* candidate sets nor shipped genotypes; no sm_100a kernel is provided for it" % name)
    return make


# operations.py:9-25 — same keys, same (C, stride, affine) signature
OPS = {
    "none": lambda C, stride, affine: Zero(stride),
    "avg_pool_3x3": lambda C, stride, affine: PoolBN("avg", C, 3, stride, 1, affine=affine),
    "max_pool_3x3": lambda C, stride, affine: PoolBN("max", C, 3, stride, 1, affine=affine),
    "skip_connect": lambda C, stride, affine: Identity() if stride == 1 else FactorizedReduce(C, C, affine=affine),
    "std_conv_3x3": lambda C, stride, affine: ReLUConvBN(C, C, 3, stride, 1, affine=affine),
    "std_conv_1x1": lambda C, stride, affine: ReLUConvBN(C, C, 1, stride, 0, affine=affine),
    "dil_conv_3x3_2": lambda C, stride, affine: DilConvS(C, C, 3, stride, 2, 2, affine=affine),
    "dil_conv_3x3_4": lambda C, stride, affine: DilConvS(C, C, 3, stride, 4, 4, affine=affine),
    "dil_conv_5x5_4": lambda C, stride, affine: DilConvS(C, C, 5, stride, 4, 2, affine=affine),
    "se_connect": lambda C, stride, affine: SE_Block(C, stride, affine=affine),
    "conv_7x1_1x7": _unsupported("conv_7x1_1x7"),
    "sep_conv_3x3": lambda C, stride, affine: Sep_Conv(C, C, 3, stride, 1, affine=affine),
    "sep_conv_5x5": lambda C, stride, affine: Sep_Conv(C, C, 5, stride, 2, affine=affine),
    "poled_conv_x1": lambda C, stride, affine: Pooled_Conv(C, C, 3, stride, 1, 1, affine=affine),
    "poled_conv_x2": lambda C, stride, affine: Pooled_Conv(C, C, 3, stride, 1, 2, affine=affine),
}


def _bn(C, affine=True):
    return BatchNorm2d(C, affine=affine, momentum=BN_MOMENTUM)


class Zero(nn.Module):
    """x * 0 (strided subsample first when stride > 1) — operations.py:31-41."""
    input_kind = "raw"

    def __init__(self, stride):
        super().__init__()
        self.stride = stride

    def forward(self, x):
        x = F_.to_internal(x)
        if self.stride != 1:
            x = x[:, :, ::self.stride, ::self.stride]
        return F_.zeros_like_internal(x)


class Identity(nn.Module):
    input_kind = "raw"

    def forward(self, x):
        return x


class PoolBN(nn.Module):
    """pool(3x3, stride, pad 1) -> BN — operations.py:44-66 (avg uses count_include_pad=False)."""

    def __init__(self, pool_type, C, kernel_size, stride, padding, affine=True):
        super().__init__()
        kind = pool_type.lower()
        if kind == "max":
            self.pool = MaxPool2d(kernel_size, stride, padding)
        elif kind == "avg":
            self.pool = AvgPool2d(kernel_size, stride, padding, count_include_pad=False)
        else:
            raise ValueError(pool_type)
        self.bn = _bn(C, affine)

    input_kind = "raw"

    def lazy(self, x):
        return self.bn.pending(self.pool(F_.check_raw(F_.to_internal(x), "pool")))

    def forward(self, x):
        return F_.finish(self.lazy(x))


class _ReLUConvBNBase(nn.Module):
    input_kind = "relu"

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation=1, affine=True):
        super().__init__()
        self.net = Sequential(
            ReLU(),
            Conv2d(C_in, C_out, kernel_size, stride, padding, dilation=dilation, bias=False),
            _bn(C_out, affine))

    def lazy(self, x):
        return self.net.lazy(x)

    def forward(self, x):
        return self.net(x)


class ReLUConvBN(_ReLUConvBNBase):
    """ReLU -> Conv(k, stride, pad, bias=False) -> BN — operations.py:69-82."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, affine=True):
        super().__init__(C_in, C_out, kernel_size, stride, padding, 1, affine)


class StdConv(ReLUConvBN):
    """operations.py:159-172 (same computation as ReLUConvBN)."""


class DilConv(_ReLUConvBNBase):
    """ReLU -> dense dilated Conv -> BN — operations.py:85-101."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation, affine=True):
        super().__init__(C_in, C_out, kernel_size, stride, padding, dilation, affine)


class DilConvS(nn.Module):
    """ReLU -> depthwise dilated k x k -> pointwise 1x1 -> BN — operations.py:202-220."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation, affine=True):
        super().__init__()
        self.net = Sequential(
            ReLU(),
            Conv2d(C_in, C_in, kernel_size, stride, padding, dilation=dilation, groups=C_in, bias=False),
            Conv2d(C_in, C_out, 1, stride=1, padding=0, bias=False),
            _bn(C_out, affine))

    input_kind = "relu_ok"

    def lazy(self, x):
        return self.net.lazy(x)

    def forward(self, x):
        return self.net(x)


class Sep_Conv(nn.Module):
    """Two stacked DilConvS(dilation=1) — operations.py:190-200."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, affine=True):
        super().__init__()
        self.net = Sequential(
            DilConvS(C_in, C_in, kernel_size, stride, padding, dilation=1, affine=affine),
            DilConvS(C_in, C_out, kernel_size, 1, padding, dilation=1, affine=affine))

    input_kind = "relu_ok"

    def lazy(self, x):
        return self.net.lazy(x)

    def forward(self, x):
        return self.net(x)


class SE_Block(nn.Module):
    """x * sigmoid(conv2(relu(conv1(gap(x))))); stride != 1 adds AvgPool2d(2) -> BN — operations.py:105-129.
    (`bn` exists even when stride == 1, exactly like the reference: it is one of the never-used
    parameters DDP must tolerate.)"""

    def __init__(self, C_in, stride, affine=True):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d(1)  # parameter-free placeholder; squeeze runs in our gap kernel
        self.conv1 = Conv2d(C_in, C_in // 2, 1, 1, 0)
        self.conv2 = Conv2d(C_in // 2, C_in, 1, 1, 0)
        self.relu = ReLU()
        self.stride = stride
        self.pool2 = AvgPool2d(2)
        self.bn = BatchNorm2d(C_in, momentum=BN_MOMENTUM)

    input_kind = "raw"

    def lazy(self, x):
        x = F_.check_raw(F_.to_internal(x), "se_connect")
        out = F_.se_scale(x, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias)
        if self.stride == 1:
            return out
        return self.bn.pending(self.pool2(out))

    def forward(self, x):
        return F_.finish(self.lazy(x))


class FactorizedReduce(nn.Module):
    """ReLU -> cat(conv1x1 s2 (x), conv1x1 s2 (x[:, :, 1:, 1:])) -> BN — operations.py:142-157.
    The shifted branch is the same stride-2 kernel reading the odd/odd TMA phase map; nothing is
    sliced or copied."""

    def __init__(self, C_in, C_out, affine=True):
        super().__init__()
        self.relu = ReLU()
        self.conv1 = Conv2d(C_in, C_out // 2, 1, stride=2, padding=0, bias=False)
        self.conv2 = Conv2d(C_in, C_out // 2, 1, stride=2, padding=0, bias=False)
        self.bn = _bn(C_out, affine)

    input_kind = "relu"

    def lazy(self, x):
        x = self.relu(x)
        a, _ = self.conv1.run(x)
        b, _ = self.conv2.run(x, hoff=1, woff=1)
        return self.bn.pending(F_.cat([a, b]))

    def forward(self, x):
        return F_.finish(self.lazy(x))


class Pooled_Conv(nn.Module):
    """AvgPool2d(2,2) -> n x [ReLU -> Conv3x3(bias) -> BN] -> bilinear x2 (align_corners=True)
    [x2 again if n == 2 and stride == 2] — operations.py:222-251."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, conv_nums, affine=True):
        super().__init__()
        layers = [AvgPool2d(2, 2)]
        for _ in range(conv_nums):
            layers += [ReLU(), Conv2d(C_in, C_out, kernel_size, stride, padding), _bn(C_out, affine)]
        layers.append(UpsamplingBilinear2d(scale_factor=2))
        if conv_nums == 2 and stride == 2:
            layers.append(UpsamplingBilinear2d(scale_factor=2))
        self.net = Sequential(*layers)

    input_kind = "raw"

    def lazy(self, x):
        return self.net.lazy(F_.check_raw(F_.to_internal(x), "poled_conv"))

    def forward(self, x):
        return self.net(x)


class FacConv(nn.Module):
    """operations.py:174-188 — present in the table only; see `_unsupported`."""

    def __init__(self, *a, **k):
        super().__init__()
        _unsupported("conv_7x1_1x7")(0, 0, 0)
