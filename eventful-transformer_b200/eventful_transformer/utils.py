"""
Position encoding, relative-position tables, drop-path and index expanders
(API mirror of the reference's eventful_transformer/utils.py).

Table preparation (bicubic resize of the learned encodings, the Toeplitz gather of the relative
embeddings) happens once per reset() and is cached, exactly as in the reference
(utils.py:53-67,151-156); the per-frame work -- adding the encoding, computing the decomposed
relative-position bias from q -- runs in the CUDA library (fused into the gate / attention kernels
on the block path).
"""

from math import prod

import torch
from torch import nn
from torch.nn import functional as func

from eventful_transformer import _native as native
from eventful_transformer.base import ExtendedModule
from eventful_transformer.counting import CountedAdd, CountedEinsum


class DropPath(ExtendedModule):
    """Stochastic depth; identity in eval mode (utils.py:10-29). Training-time only, off the gated path."""

    def __init__(self, drop_rate):
        super().__init__()
        self.drop_rate = drop_rate

    def forward(self, x):
        if not self.training:
            return x
        keep = torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), device=x.device) > self.drop_rate
        return x.div(1.0 - self.drop_rate) * keep.to(x.dtype)


class PositionEncoding(ExtendedModule):
    """Learned position encoding, bicubically resized to the input grid and cached (utils.py:32-105)."""

    def __init__(self, dim, encoding_size, input_size, has_class_token):
        super().__init__()
        self.encoding_size = tuple(encoding_size)
        self.input_size = tuple(input_size)
        self.has_class_token = has_class_token
        tokens = prod(self.encoding_size) + int(has_class_token)
        self.encoding = nn.Parameter(torch.zeros(1, tokens, dim))
        self.add = CountedAdd()
        self.cached_encoding = None

    def sized_encoding(self, batch=1):
        """The (batch, N, D) encoding for the configured input size; cached until reset()."""
        cached = self.cached_encoding
        if self.training or cached is None or cached.shape[0] != batch:
            cached = self._compute_sized_encoding().detach()
            cached = cached.expand((batch,) + tuple(cached.shape[1:])).contiguous()
            self.cached_encoding = None if self.training else cached
        return cached

    def forward(self, x):
        out = self.add(x.contiguous(), self.sized_encoding(x.shape[0]))
        return out

    def _compute_sized_encoding(self):
        enc = self.encoding
        if self.input_size == self.encoding_size:
            return enc
        cls = None
        if self.has_class_token:  # class token first (models/vivit.py:298)
            cls, enc = enc[:, :1], enc[:, 1:]
        grid = enc.transpose(1, 2).reshape(enc.shape[0], enc.shape[2], *self.encoding_size)
        grid = func.interpolate(grid, self.input_size, mode="bicubic", align_corners=False)
        enc = grid.flatten(start_dim=2).transpose(1, 2)
        if cls is not None:
            enc = torch.concat([cls, enc], dim=1)
        return enc

    def reset_self(self):
        self.cached_encoding = None


class RelativePositionEmbedding(ExtendedModule):
    """
    Decomposed relative position embeddings (utils.py:108-195).  On the block path only `tables()` is
    used: the bias q . R is computed inside the attention kernels from the unscaled q.
    """

    def __init__(self, attention_size, embedding_size, head_dim, pool_size=None):
        super().__init__()
        self.attention_size = tuple(attention_size)
        self.embedding_size = tuple(embedding_size)
        self.pool_size = pool_size
        self.y_embedding = nn.Parameter(torch.zeros(2 * embedding_size[0] - 1, head_dim))
        self.x_embedding = nn.Parameter(torch.zeros(2 * embedding_size[1] - 1, head_dim))
        self.add = CountedAdd()
        self.einsum = CountedEinsum()
        self.y_relative = None
        self.x_relative = None

    def tables(self):
        """(y_relative (ah, kh, dh), x_relative (aw, kw, dh)), contiguous, cached until reset(); the key axes are
        (ah, aw) or, with K/V pooling, the pooled grid (ah / pool_h, aw / pool_w) (utils.py:185-188)."""
        if self.y_relative is None:
            self.y_relative = self._get_relative(self.y_embedding.detach(), 0).contiguous()
            self.x_relative = self._get_relative(self.x_embedding.detach(), 1).contiguous()
        return self.y_relative, self.x_relative

    def count_fused(self, batch_heads, n_query, n_key):
        """Counters of one fused application: two einsums and two adds over the logits."""
        if self.count_mode:
            dh = self.y_embedding.shape[1]
            a = self.attention_size
            if self.pool_size is not None:  # pooled key axes (utils.py:143-147)
                a = (a[0] // self.pool_size[0], a[1] // self.pool_size[1])
            self.einsum.counts["einsum_flops"] += batch_heads * n_query * dh * (a[0] + a[1])
            self.add.counts["add_flops"] += 2 * batch_heads * n_query * n_key

    def forward(self, x, q, inplace=True):
        """Stand-alone form on materialised logits x (B, H, N, N); library ops, off the fused path."""
        a = self.attention_size
        y_rel, x_rel = self.tables()
        if self.pool_size is not None:
            raise NotImplementedError("the stand-alone rel-pos form is un-pooled; pooled keys run in the fused kernels")
        xs = x.view(x.shape[:2] + a + a)
        qs = q.reshape(q.shape[:2] + a + q.shape[-1:])
        term = self.einsum("abhwc,hkc->abhwk", qs, y_rel).unsqueeze(-1)
        xs = xs.add_(term) if inplace else xs + term
        xs.add_(self.einsum("abhwc,wkc->abhwk", qs, x_rel).unsqueeze(-2))
        if self.count_mode:
            self.add.counts["add_flops"] += 2 * xs.numel()
        return xs.view(xs.shape[:2] + (prod(a), prod(a)))

    def _get_relative(self, embedding, dim):
        size = self.embedding_size[dim]
        offsets = torch.arange(size, device=embedding.device)
        rel = embedding[offsets[:, None] - offsets[None, :] + size - 1]
        if self.embedding_size != self.attention_size:
            rel = rel.transpose(0, 2).unsqueeze(0)
            rel = func.interpolate(rel, self.attention_size, mode="bicubic", align_corners=False)
            rel = rel.squeeze(0).transpose(0, 2)
        if self.pool_size is not None:  # average the key axis over the pooling cells (one-off table preparation)
            rel = func.avg_pool1d(rel.transpose(1, 2), self.pool_size[dim]).transpose(1, 2)
        return rel

    def reset_self(self):
        self.y_relative = None
        self.x_relative = None


def expand_col_index(index, target_shape):
    """Broadcast view of a (..., k) index for gather/scatter along the last dim of target_shape."""
    extra = len(target_shape) - index.ndim
    view = index.view(index.shape[:-1] + (1,) * extra + index.shape[-1:])
    return view.expand(tuple(target_shape[:-1]) + (-1,))


def expand_row_index(index, target_shape):
    """Broadcast view of a (..., k) index for gather/scatter along the second-to-last dim."""
    extra = len(target_shape) - index.ndim
    view = index.view(index.shape[:-1] + (1,) * (extra - 1) + (index.shape[-1], 1))
    return view.expand(tuple(target_shape[:-2]) + (-1, target_shape[-1]))
