from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ..._compat import get_norm
from ...functional import conv1x1_autograd, sepconv_eval


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _same_pad(kernel, stride):
    p = max(kernel - stride, 0)
    return (p // 2, p - p // 2, p // 2, p - p // 2)          # (left, right, top, bottom), reference wrappers.py:52-54


class Conv2d(nn.Conv2d):
    """reference wrappers.py:40-164: nn.Conv2d with 'static_same' padding (explicit F.pad, then an unpadded conv), optional norm and activation."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, padding_mode="zeros",
                 norm=None, activation=None, precision=None):
        k, s = _pair(kernel_size), _pair(stride)
        self._static_same = padding_mode == "static_same"
        self._explicit_pad = _same_pad(k[0], s[0]) if self._static_same else None
        super().__init__(in_channels, out_channels, k, s, 0 if self._static_same else padding, dilation, groups, bias)
        if self._static_same:
            self.padding = self._explicit_pad                 # what the reference's module attribute holds (used by SeparableConv2d.padding)
        self.norm, self.activation, self.precision = norm, activation, precision

    def _native_1x1(self, x):
        return (x.is_cuda and self.kernel_size == (1, 1) and self.stride == (1, 1) and self.groups == 1 and self.in_channels % 32 == 0
                and self.out_channels % 32 == 0 and x.dim() == 4 and x.numel() > 0)

    def forward(self, x):
        if self._native_1x1(x):
            x = conv1x1_autograd(x, self.weight, self.bias, self.precision)
        else:
            if self._static_same:
                x = F.pad(x, self._explicit_pad)
                x = F.conv2d(x, self.weight, self.bias, self.stride, 0, self.dilation, self.groups)
            else:
                x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class SeparableConv2d(nn.Module):
    """reference wrappers.py:166-206: depthwise k x k (no bias) -> pointwise 1x1 (+bias) -> norm (eps / momentum overridden) -> activation."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True, padding_mode="zeros", norm=None,
                 eps=1e-05, momentum=0.1, activation=None, precision=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation, self.groups = _pair(kernel_size), _pair(stride), _pair(dilation), in_channels
        self.bias, self.padding_mode = bias, padding_mode
        self.depthwise = Conv2d(in_channels, in_channels, kernel_size, stride, padding, dilation, groups=in_channels, bias=False,
                                padding_mode=padding_mode)
        self.pointwise = Conv2d(in_channels, out_channels, 1, 1, 0, 1, 1, bias=bias, padding_mode=padding_mode, precision=precision)
        self.padding = self.depthwise.padding
        self.norm = None if norm == "" else norm
        if self.norm is not None:
            self.norm = get_norm(norm, out_channels)
            assert self.norm is not None
            self.norm.eps, self.norm.momentum = eps, momentum
        self.activation = activation

    def _folded(self):
        """Pointwise weight / bias with the eval-mode norm folded in: y = s * (W d + b - mean) + beta, s = gamma / sqrt(var + eps).  Cached on the
        parameter / buffer versions (weights are static at inference)."""
        n = self.norm
        ts = [self.pointwise.weight] + ([self.pointwise.bias] if self.pointwise.bias is not None else []) + \
             ([n.weight, n.bias, n.running_mean, n.running_var] if n is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in ts)
        if getattr(self, "_fold_key", None) != key:
            with torch.no_grad():
                w = self.pointwise.weight.float()
                b = self.pointwise.bias.float() if self.pointwise.bias is not None else torch.zeros(w.shape[0], device=w.device)
                if n is not None:
                    s = n.weight.float() * torch.rsqrt(n.running_var.float() + n.eps)
                    w, b = w * s.view(-1, 1, 1, 1), s * (b - n.running_mean.float()) + n.bias.float()
                self._fold_w, self._fold_b, self._fold_key = w.contiguous(), b.contiguous(), key
        return self._fold_w, self._fold_b

    def _native_eval(self, x):
        """Inference on CUDA with a BatchNorm-family norm in eval mode (or none): the whole block is one library call."""
        n = self.norm
        ok_norm = n is None or (isinstance(n, torch.nn.modules.batchnorm._BatchNorm) and not n.training and n.track_running_stats and n.affine)
        return (not torch.is_grad_enabled() and x.is_cuda and ok_norm and self.activation is None and self.kernel_size == (3, 3)
                and self.stride == (1, 1) and self.dilation == (1, 1) and self.padding_mode == "static_same" and x.dim() == 4
                and self.in_channels % 32 == 0 and self.out_channels % 32 == 0 and x.size(0) * x.size(2) <= 65535)

    def forward(self, x, pre_swish: bool = False):
        """pre_swish: apply x * sigmoid(x) in front (the neck's `conv(self._swish(fused))`), fused into the depthwise pass on the native path."""
        if self._native_eval(x):
            w, b = self._folded()
            return sepconv_eval(x, self.depthwise.weight, w, b, pre_swish, self.pointwise.precision)
        if pre_swish:
            x = x * torch.sigmoid(x)
        x = self.pointwise(self.depthwise(x))
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class MaxPool2d(nn.Module):
    """reference wrappers.py:209-252: ZERO-pads (right / bottom for k3 s2) and then max-pools without padding -- the padding takes part in the max."""

    def __init__(self, kernel_size, stride=None, padding=0, dilation=1, return_indices=False, ceil_mode=False, padding_mode="static_same"):
        super().__init__()
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride) if stride is not None else self.kernel_size
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.return_indices, self.ceil_mode, self.padding_mode = return_indices, ceil_mode, padding_mode
        if padding_mode == "static_same":
            self.padding = _same_pad(self.kernel_size[0], self.stride[0])
        elif padding_mode == "dynamic_same":
            self.padding = (0, 0)

    def forward(self, x):
        return F.max_pool2d(F.pad(x, self.padding), self.kernel_size, self.stride, 0, self.dilation, self.ceil_mode, self.return_indices)


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


class MemoryEfficientSwish(Swish):
    """reference activations.py:17-36: same function (the reference's custom autograd Function only saves memory)."""
