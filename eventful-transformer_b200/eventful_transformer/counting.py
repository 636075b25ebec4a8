"""
Counted operators: they own the parameters (state-dict keys `weight`, `bias`) and the MAC
counters, and run on libeventful_b200.  API mirror of the reference's counting.py.

Counter semantics (MACs, not FLOPs) follow counting.py:16-19,43-47,98-110,119-124,146-162,171-175.
Parameters are zero-initialised like the reference's (counting.py:143-144): callers load or
initialise weights explicitly.
"""

from math import prod

import torch
import torch.nn as nn

from eventful_transformer import _native as native
from eventful_transformer.base import ExtendedModule, numeric_tuple


class CountedAdd(ExtendedModule):
    """result = a + b (optionally in place on a); counts result.numel() adds."""

    def forward(self, a, b, inplace=False):
        if a.shape != b.shape:
            b = b.expand_as(a).contiguous()  # broadcast operand materialised once (e.g. position encoding)
        result = native.add(a, b, out=a if inplace else None)
        if self.count_mode:
            self.counts["add_flops"] += result.numel()
        return result


class CountedBias(ExtendedModule):
    """x + bias over a channel dimension followed by `spatial_dims` trailing dims."""

    def __init__(self, features, spatial_dims=0, device=None, dtype=None):
        super().__init__()
        self.features = features
        self.spatial_dims = spatial_dims
        self.bias = nn.Parameter(torch.zeros(features, device=device, dtype=dtype))

    def forward(self, x):
        bias = self.bias.view((self.features,) + (1,) * self.spatial_dims).expand_as(x).contiguous()
        result = native.add(x.contiguous(), bias)
        if self.count_mode:
            self.counts["bias_flops"] += result.numel()
        return result


class CountedConv(ExtendedModule):
    """
    Convolution used by the patch / tubelet embeddings in models/ (outside the gated path,
    SURVEY.md 8(f4)).  Keeps the reference's parameter layout and counter; the convolution
    itself is delegated to the cuDNN library call, as a plain library op off the hot path.
    """

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, device=None, dtype=None):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = numeric_tuple(kernel_size, length=spatial_dims)
        self.stride = numeric_tuple(stride, length=spatial_dims)
        self.padding = numeric_tuple(padding, length=spatial_dims) if isinstance(padding, int) else padding
        self.dilation = numeric_tuple(dilation, length=spatial_dims)
        self.groups = groups
        self.conv_function = getattr(torch.nn.functional, f"conv{spatial_dims}d")
        shape = (out_channels, in_channels // groups) + self.kernel_size
        self.weight = nn.Parameter(torch.zeros(shape, device=device, dtype=dtype))

    def forward(self, x):
        result = self.conv_function(x, self.weight, stride=self.stride, padding=self.padding,
                                    dilation=self.dilation, groups=self.groups)
        if self.count_mode:
            fan_in = (self.in_channels // self.groups) * prod(self.kernel_size)
            self.counts[f"conv{self.spatial_dims}d_flops"] += result.numel() * fan_in
        return result


class CountedEinsum(ExtendedModule):
    """
    Einsum with a MAC counter.  Only the rel-pos einsums sit on the gated path and those are fused into
    the attention kernels; this generic operator remains for API compatibility (library call).
    """

    def forward(self, equation, *operands):
        if self.count_mode:
            ones = [torch.ones_like(x) for x in operands]
            self.counts["einsum_flops"] += int(torch.einsum(equation, *ones).sum())
        return torch.einsum(equation, *operands)


class CountedLinear(ExtendedModule):
    """y = x W^T + b on the tcgen05 GEMM; weight (out, in), bias (out)."""

    def __init__(self, in_features, out_features, device=None, dtype=None):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = nn.Parameter(torch.zeros((out_features, in_features), device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(out_features, device=device, dtype=dtype))

    def count_linear(self, n_rows, with_bias=True):
        """Adds the counters of a linear applied to n_rows input rows (used by the fused block path)."""
        if self.count_mode:
            if with_bias:
                self.counts["bias_flops"] += n_rows * self.out_features
            self.counts["linear_flops"] += n_rows * self.in_features * self.out_features

    def forward_bias(self, x):
        result = native.add(x.contiguous(), self.bias.detach().expand_as(x).contiguous())
        if self.count_mode:
            self.counts["bias_flops"] += result.numel()
        return result

    def forward_linear(self, x):
        if self.count_mode:
            self.counts["linear_flops"] += x.numel() * self.out_features
        return native.linear(x.contiguous(), self.weight.detach(), None)

    def forward(self, x, act=native.ACT_NONE, out=None, idx=None):
        result = native.linear(x.contiguous(), self.weight.detach(), self.bias.detach(), act=act, out=out, idx=idx)
        self.count_linear(x.numel() // self.in_features)
        return result


class CountedMatmul(ExtendedModule):
    """Batched a @ b with a MAC counter (strided operands accepted)."""

    def forward(self, a, b):
        result = native.bmm(a, b)
        if self.count_mode:
            self.counts["matmul_flops"] += result.numel() * a.shape[-1]
        return result
