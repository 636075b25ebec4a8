"""
Operators that own parameters (state-dict keys `weight`, `bias`) and MAC counters, executed by
libeventful_b200.  Same class names, constructor arguments and counter keys as the reference's
counting.py, so its models build on top of this module unchanged.

What is counted (multiply-accumulates, not FLOPs; reference counting.py:16-19,43-47,98-110,119-124,
146-162,171-175):
    add_flops / bias_flops   one per output element
    linear_flops             rows x in_features x out_features
    matmul_flops             output elements x inner dimension
    conv{n}d_flops           output elements x (in_channels / groups) x prod(kernel)
    einsum_flops             the einsum of all-ones operands
Parameters start at zero, as in the reference (counting.py:143-144): weights are loaded or initialised by
the caller.
"""

from math import prod

import torch
import torch.nn as nn
import torch.nn.functional as F

from eventful_transformer import _native as native
from eventful_transformer.base import ExtendedModule, numeric_tuple


class _Counted(ExtendedModule):
    def _tally(self, key, macs):
        if self.count_mode:
            self.counts[key] += int(macs)


def _zeros_parameter(shape, device, dtype):
    return nn.Parameter(torch.zeros(shape, device=device, dtype=dtype))


class CountedAdd(_Counted):
    """Elementwise sum (residual connections, position encoding); `inplace` reuses the first operand."""

    def forward(self, a, b, inplace=False):
        rhs = b if b.shape == a.shape else b.expand_as(a).contiguous()  # e.g. a (1, N, D) position encoding
        total = native.add(a, rhs, out=a if inplace else None)
        self._tally("add_flops", total.numel())
        return total


class CountedBias(_Counted):
    """Adds a per-channel bias; the channel axis is followed by `spatial_dims` trailing axes."""

    def __init__(self, features, spatial_dims=0, device=None, dtype=None):
        super().__init__()
        self.features, self.spatial_dims = features, spatial_dims
        self.bias = _zeros_parameter(features, device, dtype)

    def forward(self, x):
        shaped = self.bias.view((self.features,) + (1,) * self.spatial_dims)
        total = native.add(x.contiguous(), shaped.expand_as(x).contiguous())
        self._tally("bias_flops", total.numel())
        return total


class CountedConv(_Counted):
    """
    N-d convolution of the patch / tubelet embeddings in models/ (outside the gated path).
    Parameter layout and counter follow the reference; the convolution itself is the cuDNN library op.
    """

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, device=None, dtype=None):
        super().__init__()
        per_axis = lambda value: numeric_tuple(value, length=spatial_dims)  # noqa: E731
        self.spatial_dims, self.groups = spatial_dims, groups
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = per_axis(kernel_size), per_axis(stride), per_axis(dilation)
        self.padding = per_axis(padding) if isinstance(padding, int) else padding  # strings such as "same" pass through
        self.conv_function = getattr(F, f"conv{spatial_dims}d")
        self.weight = _zeros_parameter((out_channels, in_channels // groups) + self.kernel_size, device, dtype)

    def forward(self, x):
        y = self.conv_function(x, self.weight, None, self.stride, self.padding, self.dilation, self.groups)
        self._tally(f"conv{self.spatial_dims}d_flops", y.numel() * (self.in_channels // self.groups) * prod(self.kernel_size))
        return y


class CountedEinsum(_Counted):
    """
    Generic einsum with a MAC counter, kept for API compatibility (library call).  The rel-pos einsums of the
    gated path never come here: they are fused into the attention kernels.
    """

    def forward(self, equation, *operands):
        if self.count_mode:
            self._tally("einsum_flops", torch.einsum(equation, *[torch.ones_like(op) for op in operands]).sum())
        return torch.einsum(equation, *operands)


class CountedLinear(_Counted):
    """Affine layer on the tcgen05 GEMM: weight (out_features, in_features), bias (out_features)."""

    def __init__(self, in_features, out_features, device=None, dtype=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = _zeros_parameter((out_features, in_features), device, dtype)
        self.bias = _zeros_parameter(out_features, device, dtype)

    def reset_self(self):
        self._pad_key = self._pad = None

    def count_linear(self, n_rows, with_bias=True):
        """Counters of this layer applied to n_rows rows (the fused block path calls the GEMM directly)."""
        if with_bias:
            self._tally("bias_flops", n_rows * self.out_features)
        self._tally("linear_flops", n_rows * self.in_features * self.out_features)

    def _padded_parameters(self):
        """16-bit GEMM tiles store 16-byte vectors: an odd feature count (e.g. a 97-class head) runs on zero-padded copies."""
        w, b = self.weight.detach(), self.bias.detach()
        key = (w.data_ptr(), b.data_ptr(), w.dtype, w.device)  # dropped by reset(), like the other cached tables
        if getattr(self, "_pad_key", None) != key:
            f = (self.out_features + 7) // 8 * 8
            wp = torch.zeros((f, self.in_features), dtype=w.dtype, device=w.device)
            bp = torch.zeros((f,), dtype=b.dtype, device=b.device)
            wp[: self.out_features].copy_(w)
            bp[: self.out_features].copy_(b)
            self._pad_key, self._pad = key, (wp, bp)
        return self._pad

    def forward(self, x, act=native.ACT_NONE, out=None, idx=None, count=None, rows=None):
        """`count`: device-side number of valid rows per batch entry; `rows`: their host-side total for the counters."""
        if self.out_features % 8 and self.weight.dtype != torch.float32 and out is None and idx is None:
            wp, bp = self._padded_parameters()
            y = native.linear(x.contiguous(), wp, bp, act=act)[..., : self.out_features]
        else:
            y = native.linear(x.contiguous(), self.weight.detach(), self.bias.detach(), act=act, out=out, idx=idx, count=count)
        self.count_linear(x.numel() // self.in_features if rows is None else rows)
        return y

    def forward_gathered(self, src, a_idx, state=None, act=native.ACT_NONE, out=None, idx=None, count=None, rows=None):
        """forward() on the rows src[b, a_idx[b, j]], gathered by the GEMM itself; `state` is advanced at those rows."""
        y = native.linear_gather(src, a_idx, self.weight.detach(), self.bias.detach(), state=state, act=act, out=out,
                                 idx=idx, count=count)
        self.count_linear(a_idx.numel() if rows is None else rows)
        return y

    def forward_linear(self, x):
        """The product alone (reference counting.py:157-158)."""
        self.count_linear(x.numel() // self.in_features, with_bias=False)
        return native.linear(x.contiguous(), self.weight.detach(), None)

    def forward_bias(self, x):
        """The bias alone: what a zero token maps to, used for window padding (reference counting.py:152-155)."""
        y = native.add(x.contiguous(), self.bias.detach().expand_as(x).contiguous())
        self._tally("bias_flops", y.numel())
        return y


class CountedMatmul(_Counted):
    """Batched matrix product with a MAC counter; strided operands are accepted."""

    def forward(self, a, b):
        y = native.bmm(a, b)
        self._tally("matmul_flops", y.numel() * a.shape[-1])
        return y
