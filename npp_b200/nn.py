"""Leaf modules with torch.nn's names and parameter layout, executing on libnpp_b200 kernels.

The reference builds every primitive as an nn.Sequential of nn.ReLU / nn.Conv2d / nn.BatchNorm2d
(models/operations.py:69-251).  To stay state_dict- and init-compatible the same module tree is
kept here (same child indices, same parameter names and shapes); the `Sequential` below walks its
children and fuses neighbouring leaves into single kernels where the arithmetic allows it:

  ReLU -> depthwise Conv2d           : ReLU applied inside the depthwise kernel's loads
  Conv2d -> BatchNorm2d              : batch statistics accumulated in the conv epilogue
  BatchNorm2d -> ReLU                : ReLU fused into the normalise pass

Modules accept internal tensors (functional.py) or plain NCHW fp32 CUDA tensors (converted once).
"""
import torch
import torch.nn as nn

from . import functional as F_


class ReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace  # kept for repr parity; kernels are out-of-place

    def forward(self, x):
        return F_.relu(x if isinstance(x, F_.Pending) else F_.to_internal(x))


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameters (OIHW fp32 master weights) with the forward on our kernels.
    groups == 1: tcgen05 implicit GEMM (bf16) / direct fp32 kernel (validation mode);
    groups == in_channels == out_channels: depthwise kernel."""

    def _check(self):
        if self.padding_mode != "zeros" or self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1] \
                or self.dilation[0] != self.dilation[1]:
            raise RuntimeError("Conv2d configuration not supported by npp_b200 kernels: %s" % self)

    @property
    def is_depthwise(self):
        return self.groups > 1 and self.groups == self.in_channels == self.out_channels

    def run(self, x, relu_in=False, want_stats=False, hoff=0, woff=0):
        self._check()
        if isinstance(x, F_.Pending) and relu_in:
            x, relu_in = F_.relu(x), False   # one pass writes relu(bn(.)); nothing else reads the raw value
        x = F_.to_internal(x)
        if self.is_depthwise:
            if self.bias is not None or hoff or woff:
                raise RuntimeError("depthwise conv with bias/offset is not implemented")
            r = F_.relu_of(x)
            if relu_in and r is not None:   # the producer already wrote relu(x): read that, no mask needed
                x, relu_in = r, False
            elif not relu_in:
                F_.check_raw(x, "depthwise conv")
            y = F_.dwconv2d(x, self.weight, self.stride[0], self.padding[0], self.dilation[0], relu_in)
            return y, None
        if self.groups != 1:
            raise RuntimeError("grouped convolution (groups=%d) is not implemented" % self.groups)
        if relu_in:
            x = F_.relu(x)
        else:
            F_.check_raw(x, "conv")
        if (self.in_channels == 3 and self.kernel_size == (3, 3) and self.dilation[0] == 1 and not hoff and not woff
                and x.dtype == torch.bfloat16 and x.shape[1] == 8 and not x.requires_grad
                and F_._state.get("stem_im2col", True)):
            # image stem: taps folded into the channel axis (K = 27 -> 32), then a 1x1 convolution whose weight
            # matrix [Cout, 27] is the OIHW master weight itself (same memory, so dW lands in the parameter's slot)
            xc = F_.im2col3x3_c3(x, self.stride[0], self.padding[0])
            y, stats = F_.conv2d(xc, self.weight.view(self.out_channels, 27, 1, 1), self.bias, 1, 0, 1, 0, 0, want_stats,
                                 slot_of=self.weight)
            return y, (stats if stats.numel() else None)
        y, stats = F_.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0], self.dilation[0], hoff, woff,
                             want_stats)
        return y, (stats if stats.numel() else None)

    def forward(self, x):
        return self.run(x)[0]


class BatchNorm2d(nn.Module):
    """nn.BatchNorm2d state (weight, bias, running_mean, running_var, num_batches_tracked) on our
    kernels.  Deliberately NOT an nn.modules.batchnorm._BatchNorm subclass: the reference driver's
    nn.SyncBatchNorm.convert_sync_batchnorm (augment_lip_sync.py:191) must leave it alone — cross-rank
    statistics are handled here (npp_b200.distributed.enable_sync_bn) with one all-reduce of the
    raw (sum, sum-of-squares) vector."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.track_running_stats = affine, track_running_stats
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        if track_running_stats:
            self.register_buffer("running_mean", torch.zeros(num_features))
            self.register_buffer("running_var", torch.ones(num_features))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            self.register_buffer("running_mean", None)
            self.register_buffer("running_var", None)
            self.register_buffer("num_batches_tracked", None)

    def extra_repr(self):
        return "{num_features}, eps={eps}, momentum={momentum}, affine={affine}".format(**self.__dict__)

    def pending(self, x, stats=None):
        """The not-yet-normalised output: consumers (functional.node) fold scale/shift into their own pass."""
        return F_.Pending(F_.check_raw(F_.to_internal(x), "BatchNorm2d"), stats, self)

    def run(self, x, stats=None, relu=False):
        p = self.pending(x, stats)
        return F_.relu(p) if relu else F_.finish(p)

    def forward(self, x):
        return self.run(x)


class MaxPool2d(nn.Module):
    def __init__(self, kernel_size, stride=None, padding=0):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride if stride is not None else kernel_size, padding

    def forward(self, x):
        x = F_.to_internal(x)
        if self.kernel_size == 3 and self.padding == 1 and self.stride in (1, 2):
            return F_.max_pool3x3(x, self.stride)
        raise RuntimeError("MaxPool2d(%s, %s, %s) is not implemented" % (self.kernel_size, self.stride, self.padding))


class AvgPool2d(nn.Module):
    def __init__(self, kernel_size, stride=None, padding=0, count_include_pad=True):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride if stride is not None else kernel_size, padding
        self.count_include_pad = count_include_pad

    def forward(self, x):
        x = F_.to_internal(x)
        if self.kernel_size == 2 and self.stride == 2 and self.padding == 0:
            return F_.avg_pool2x2(x)
        if self.kernel_size == 3 and self.padding == 1 and self.stride in (1, 2) and not self.count_include_pad:
            return F_.avg_pool3x3(x, self.stride)
        raise RuntimeError("AvgPool2d(%s, %s, %s) is not implemented" % (self.kernel_size, self.stride, self.padding))


class UpsamplingBilinear2d(nn.Module):
    """nn.UpsamplingBilinear2d == bilinear, align_corners=True (operations.py:241)."""

    def __init__(self, scale_factor):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x):
        return F_.interpolate(F_.to_internal(x), scale_factor=self.scale_factor, mode="bilinear", align_corners=True)


class Sequential(nn.Sequential):
    """nn.Sequential whose forward fuses neighbouring leaves (see module docstring).  `lazy(x)` is the same
    chain but may return a functional.Pending when it ends in a BatchNorm2d, so that the caller (a cell node)
    applies the normalisation inside its own pass; `forward(x)` always returns a tensor."""

    def forward(self, x):
        return F_.finish(self.lazy(x))

    def lazy(self, x):
        mods = list(self)
        i, n = 0, len(mods)
        while i < n:
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < n else None
            if isinstance(m, ReLU) and isinstance(nxt, Conv2d):
                x, stats = nxt.run(x, relu_in=True, want_stats=_bn_follows(mods, i + 2))
                i += 2
                x, i = _maybe_bn(mods, i, x, stats)
            elif isinstance(m, Conv2d):
                x, stats = m.run(x, want_stats=_bn_follows(mods, i + 1))
                i += 1
                x, i = _maybe_bn(mods, i, x, stats)
            elif isinstance(m, BatchNorm2d):
                x, i = _maybe_bn(mods, i, x, None)
            elif isinstance(m, ReLU):
                x = F_.relu(x)
                i += 1
            else:
                x = call_lazy(m, x)
                i += 1
        return x


def call_lazy(m, x):
    """m(x), allowing a Pending BatchNorm output where the module supports it."""
    f = getattr(m, "lazy", None)
    return f(x) if f is not None else m(x)


def _bn_follows(mods, i):
    return i < len(mods) and isinstance(mods[i], BatchNorm2d) and mods[i].training


def _maybe_bn(mods, i, x, stats):
    if i < len(mods) and isinstance(mods[i], BatchNorm2d):
        p = mods[i].pending(x, stats)
        if i + 1 < len(mods) and isinstance(mods[i + 1], ReLU) and i + 2 < len(mods):
            return F_.relu(p), i + 2   # BN -> ReLU -> more layers: one pass writes relu(bn(x)) only
        return p, i + 1
    return x, i
