"""
Transformer blocks: Block, EventfulTokenwiseBlock, EventfulMatmul1Block, EventfulBlock
(API mirror of the reference's blocks.py: same constructor kwargs, sub-module names and
state-dict keys, resolvable by name from backbones.ViTBackbone).

How a gated block runs here (incremental frame), per gate site:
    et_gate_select   [residual add] + LayerNorm + (c - p) + token norm + radix top-k, one launch
    et_gate_gather   c~ = LN(x)[idx], p[idx] = c~
    et_linear        tcgen05 GEMM on the k gathered rows, rows scattered into the TokenBuffer
  (opt-in two-launch form, EVENTFUL_B200_FUSE_GATHER=1: et_gate_select also emits c for every token and
   et_linear_gather gathers c[idx] with TMA tile::gather4, advances p[idx] and scatters -- slower on B200, see FUSE_GATHER)
and for attention
    et_window_attention   windowed blocks (dense, blocks.py:205-240 of the reference)
    et_global_attention   global blocks: softmax statistics + A-gate / v-gate / delta accumulation
A block hands its output to the next one as an un-summed pair (branch, skip) so that the residual
add is fused into the next gate's load; Block.forward(x) is still the ordinary stand-alone call.

Selections travel through a block as (index, count): `count` is None for top-k policies (all k
entries valid) or a device-side int32 tensor (threshold policy, pooled unique indices) that every
consumer kernel takes as a pointer -- no host synchronisation, so those frames are CUDA-graph
capturable too.

Variants: K/V pooling (pool_size; global blocks), matmul_2_cast different from
the model dtype and fp32 models run on the general-precision kernels (csrc/et_generic.cu).
Adaptive token sampling (ats_fraction, blocks.py:150-181 of the reference): et_ats_scores emits the per-head token
scores, the tiny (B, H, N) normalise / sum / top-k / stabilise step runs on the host side exactly as the reference
does it (including its axis handling, which needs batch == heads), and attention then runs on the kept query rows only
(et_global_attention_rows).  Not implemented: drop-path in training mode.
"""

import os
from math import prod, sqrt

import torch
import torch.nn as nn

from eventful_transformer import _native as native
from eventful_transformer.base import ExtendedModule, numeric_tuple
from eventful_transformer.counting import CountedAdd, CountedLinear, CountedMatmul
from eventful_transformer.modules import (
    MatmulBuffer,
    MatmulDeltaAccumulator,
    SimpleSTGTGate,
    TokenBuffer,
    TokenDeltaGate,
    TokenGate,
    _policy_spec,
)
from eventful_transformer.utils import DropPath, RelativePositionEmbedding

LN_EPS = 1e-6

# Gate sites run et_gate_select -> et_gate_gather -> et_linear.  EVENTFUL_B200_FUSE_GATHER=1 switches to the two-launch form
# et_gate_select (+ c for every token) -> et_linear_gather (TMA gather4 A operand + state advance in the GEMM).  It is
# parity-tested but OFF by default: measured on B200 (profiles/r2_gather_gemm_bench.txt) a gather4 moves 512 bytes per TMA
# instruction against 16 KB for a tile load, and the GEMMs of the gate sites run 1.6-2.2x slower (8 streams: 535 vs 622
# frames/s end to end), far more than the deleted gather launch saves.
FUSE_GATHER = os.environ.get("EVENTFUL_B200_FUSE_GATHER", "0") == "1"


class _Gathered:
    """A gate site's selected rows left in place: (source of every token's gate input, index, gate state to advance)."""

    __slots__ = ("src", "index", "state")

    def __init__(self, src, index, state):
        self.src, self.index, self.state = src, index, state

# single-pass attention kernel is used for dense attention over at most this many keys
_SMALL_ATTENTION = 512

def _identity_index(cache, batch, n, device):
    """arange(n) per batch entry, cached per block instance until reset() (a captured CUDA graph may hold it)."""
    key = (batch, n, device)
    idx = cache.get(key)
    if idx is None:
        idx = torch.arange(n, device=device, dtype=torch.int64).repeat(batch, 1).contiguous()
        cache[key] = idx
    return idx


class Block(ExtendedModule):
    """Dense (non-eventful) Transformer block with optional windowed attention and rel-pos bias."""

    def __init__(self, dim, heads, input_size, mlp_ratio, ats_fraction=None, drop_path_rate=0.0,
                 relative_embedding_size=None, matmul_2_cast=None, pool_size=None, window_size=None):
        super().__init__()
        self.heads = heads
        self.input_size = tuple(input_size)
        # token grid handed to the attention kernels: (h, w), or (1, t) for 1-D inputs (ViViT's temporal sub-model)
        self._grid = self.input_size if len(self.input_size) == 2 else (1, prod(self.input_size))
        self._identity = {}
        if ats_fraction is not None:
            assert pool_size is None
            assert window_size is None
            assert not (ats_fraction < 0.0 or ats_fraction > 1.0)
            if relative_embedding_size is not None:
                # the reference's own rel-pos tables are sized to the full token grid and no longer match the
                # pruned token count one block later; no shipped configuration combines the two
                raise NotImplementedError("eventful_b200: adaptive token sampling with relative position embeddings")
        assert not (drop_path_rate < 0.0 or drop_path_rate > 1.0)
        assert matmul_2_cast in [None, "float16", "bfloat16"]
        self.ats_fraction = ats_fraction
        self.last_ats_indices = None
        self._ats_now = None
        self.matmul_2_cast = matmul_2_cast
        self.pool_size = None if pool_size is None else numeric_tuple(pool_size, length=2)
        if self.pool_size is not None and window_size is not None:
            # every shipped config pools the global blocks only (configs/*/vitdet_vid/_spatial.yml: windowed_overrides
            # pool_size null)
            raise NotImplementedError("eventful_b200: K/V pooling inside windowed attention is not implemented")
        if window_size is None:
            self.window_size = None
            attention_size = self.input_size
        else:
            self.window_size = numeric_tuple(window_size, length=2)
            attention_size = self.window_size
            if relative_embedding_size is not None:
                relative_embedding_size = self.window_size
        self.scale = sqrt(dim // heads)
        self.dim = dim

        self.input_layer_norm = nn.LayerNorm(dim, eps=LN_EPS)
        self.qkv = CountedLinear(in_features=dim, out_features=dim * 3)
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else nn.Identity()
        if relative_embedding_size is not None:
            self.relative_position = RelativePositionEmbedding(
                attention_size, tuple(relative_embedding_size), dim // heads, pool_size=self.pool_size)
        else:
            self.relative_position = None
        self.matmul = CountedMatmul()
        self.projection = CountedLinear(in_features=dim, out_features=dim)
        self.add = CountedAdd()
        self.mlp_layer_norm = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp_1 = CountedLinear(in_features=dim, out_features=dim * mlp_ratio)
        self.gelu = nn.GELU()
        self.mlp_2 = CountedLinear(in_features=dim * mlp_ratio, out_features=dim)

    # ------------------------------------------------------------------ public forward
    def forward(self, x):
        branch, skip = self._forward_pair(x.contiguous(), None)
        return native.add(branch, skip)

    def reset_self(self):
        self.last_ats_indices = None
        self._ats_now = None
        self._identity = {}

    # ------------------------------------------------------------------ adaptive token sampling
    def _ats_select(self, qkv, score_dtype):
        """
        Block._adaptive_token_sampling (reference blocks.py:150-176) on the QKV buffer: the stabilised index of the
        tokens whose query rows are kept, (B, n_select) int64.  The raw scores a[..., 0] * |v| come from the
        statistics kernel; what follows works on a (B, H, N) tensor and is kept in the reference's arithmetic
        (same dtype, same reductions): divide by the sum over the non-class tokens, pin the class token, sum over
        axis -3, keep the int(fraction * (N - 1)) + 1 best, sort, stabilise.  Axis -3 of the 3-D score tensor is the
        BATCH axis and row h of the result is then used for batch entry h: defined for batch == heads only (the
        reference raises in .gather() otherwise; ViViT-B runs it with 12 batched views and 12 heads).
        """
        raw = native.ats_scores(qkv, self.heads, score_dtype)
        score = raw / raw[..., 1:].sum(dim=-1, keepdim=True)
        score[..., 0] = float("inf")
        score = score.sum(dim=-3)
        n_select = int(self.ats_fraction * (score.shape[-1] - 1)) + 1
        index = score.topk(n_select, sorted=False)[1].sort(dim=-1)[0]
        if index.shape[0] != qkv.shape[0]:
            raise RuntimeError(f"eventful_b200: adaptive token sampling yields {index.shape[0]} index rows (one per head) "
                               f"for a batch of {qkv.shape[0]}; the reference's gather (blocks.py:179) needs batch == heads")
        index = self._stabilize_ats_indices(index)
        self.last_ats_indices = self._ats_now = index
        return index

    def _stabilize_ats_indices(self, index):
        """Tokens that stay keep last frame's slot; slots of the tokens that left take the new tokens in ascending
        order (reference blocks.py:378-391, a host-side loop there too)."""
        if self.last_ats_indices is None:
            return index
        new, old = index.cpu(), self.last_ats_indices.cpu()
        out = old.clone()
        for r in range(new.shape[0]):
            out[r, ~torch.isin(old[r], new[r])] = new[r, ~torch.isin(new[r], old[r])]
        return out.to(index.device)

    def _ats_rows(self, x):
        """x[b, ats_index[b]]: the kept rows of a (B, N, C) tensor (query rows of the QKV buffer, or the skip tensor:
        Block._gather_ats_skip, reference blocks.py:196-203)."""
        if self.ats_fraction is None or self._ats_now is None:
            return x
        return native.gate_gather(x, self._ats_now)[0]

    def _ats_dense(self, qkv):
        """Dense attention of the kept query rows over all keys (Block / Tokenwise / Matmul1 blocks: ATS precedes the
        matmul_2 cast, reference blocks.py:229-231,499-500)."""
        index = self._ats_select(qkv, qkv.dtype)
        out = native.global_attention_rows(qkv, index, self.heads, native.ATTN_DENSE,
                                           state_dtype=self._state_dtype(qkv.dtype))
        if self.count_mode:
            b, n, _ = qkv.shape
            self.matmul.counts["matmul_flops"] += b * self.heads * (n + index.shape[-1]) * n * (self.dim // self.heads)
        return out

    # ------------------------------------------------------------------ shared helpers
    def _check_input(self, x):
        native.require_device(x)
        if x.dtype not in native._DTYPES:
            raise TypeError(f"eventful_b200: unsupported activation dtype {x.dtype} (float32, bfloat16, float16)")
        if self.qkv.weight.dtype != x.dtype:
            raise TypeError(f"eventful_b200: the input is {x.dtype} but the block's parameters are "
                            f"{self.qkv.weight.dtype}; cast the model or the input")
        if self.training:
            raise NotImplementedError("eventful_b200 is an inference path (no autograd through the kernels, drop-path "
                                      "is identity): call .eval()")
        if torch.is_grad_enabled() and (x.requires_grad or self.qkv.weight.requires_grad):
            raise RuntimeError("eventful_b200 is an inference path: its kernels record no autograd graph, so gradients "
                               "would silently be missing. Run under torch.inference_mode() / torch.no_grad() "
                               "(as scripts/time/vitdet_vid.py:28 of the reference does) or freeze the parameters.")
        if self.window_size is not None and self._state_dtype(x.dtype) != x.dtype:
            raise NotImplementedError("eventful_b200: matmul_2_cast different from the model dtype is implemented for "
                                      "global attention only (the reference's configs set it to null on windowed blocks)")

    def _state_dtype(self, model_dtype):
        """Element type of a, v, the v-gate / A-gate state and the accumulator (_cast_matmul_2, blocks.py:183-189)."""
        return model_dtype if self.matmul_2_cast is None else getattr(torch, self.matmul_2_cast)

    @staticmethod
    def _ln_params(ln):
        return ln.weight.detach(), ln.bias.detach()

    def _layer_norm_all(self, ln, x):
        """Dense LayerNorm over every token (rows gathered with the identity index)."""
        idx = _identity_index(self._identity, x.shape[0], x.shape[1], x.device)
        out, _ = native.gate_gather(x, idx, ln=self._ln_params(ln), eps=ln.eps)
        return out

    def _count_adds(self, x):
        if self.count_mode:  # both residual adds of this block (the second is fused downstream)
            self.add.counts["add_flops"] += 2 * x.numel()

    def _rel_tables(self, dtype):
        if self.relative_position is None:
            return None
        y_rel, x_rel = self.relative_position.tables()
        if y_rel.dtype != dtype:
            y_rel, x_rel = y_rel.to(dtype), x_rel.to(dtype)
        return y_rel, x_rel

    def _window_grid(self):
        wh, ww = self.window_size
        gh, gw = self._grid
        th, tw = gh + (-gh % wh), gw + (-gw % ww)
        return (th // wh) * (tw // ww), (th, tw) != (gh, gw)

    def _dense_attention(self, qkv):
        """Block._forward_attention of the reference, fused (window partition .. recombine)."""
        if self.ats_fraction is not None:
            return self._ats_dense(qkv)
        b, n, _ = qkv.shape
        dh = self.dim // self.heads
        rel = self._rel_tables(qkv.dtype)
        if self.window_size is not None:
            n_win, padded = self._window_grid()
            w2 = prod(self.window_size)
            out = native.window_attention(qkv, self.heads, self._grid, self.window_size,
                                          pad_token=self.qkv.bias.detach(), rel=rel)
            if self.count_mode:
                if padded:
                    self.qkv.counts["bias_flops"] += self.qkv.out_features  # forward_bias on the pad token
                self.matmul.counts["matmul_flops"] += 2 * b * n_win * self.heads * w2 * w2 * dh
                if rel is not None:
                    self.relative_position.count_fused(b * n_win * self.heads, w2, w2)
            return out
        sdt = self._state_dtype(qkv.dtype)
        n_keys = n
        if self.pool_size is not None or sdt != qkv.dtype:
            kvp = None
            if self.pool_size is not None:
                kvp = native.pool_kv(qkv, self._grid, self.pool_size)
                n_keys = kvp.shape[1]
            out = native.global_attention(qkv, self.heads, self._grid, native.ATTN_DENSE, rel=rel, kv_pooled=kvp,
                                          pool=self.pool_size, state_dtype=sdt)
        elif n <= _SMALL_ATTENTION:
            out = native.window_attention(qkv, self.heads, self._grid, None, rel=rel)
        else:
            out = native.global_attention(qkv, self.heads, self._grid, native.ATTN_DENSE, rel=rel)
        if self.count_mode:
            self.matmul.counts["matmul_flops"] += 2 * b * self.heads * n * n_keys * dh
            if rel is not None:
                self.relative_position.count_fused(b * self.heads, n, n_keys)
        return out

    @staticmethod
    def _linear(layer, c, **kw):
        """layer(c) for a materialised c~ or for rows still to be gathered by the GEMM (_Gathered)."""
        if isinstance(c, _Gathered):
            return layer.forward_gathered(c.src, c.index, state=c.state, **kw)
        return layer(c, **kw)

    def _mlp(self, c, out=None, idx=None, count=None, rows=None):
        h = self._linear(self.mlp_1, c, act=native.ACT_GELU, rows=rows, count=count if isinstance(c, _Gathered) else None)
        return self.mlp_2(h, out=out, idx=idx, count=count, rows=rows)

    # Pair protocol: input is xa (+ xb), output is (branch, skip) whose sum is the block output.
    def _forward_pair(self, xa, xb):
        self._check_input(xa)
        x = native.add(xa, xb) if xb is not None else xa
        c = self._layer_norm_all(self.input_layer_norm, x)
        attn = self._dense_attention(self.qkv(c))
        x2 = native.add(self.projection(attn), self._ats_rows(x))
        c2 = self._layer_norm_all(self.mlp_layer_norm, x2)
        branch = self._mlp(c2)
        self._count_adds(x2)
        return branch, x2


class EventfulTokenwiseBlock(Block):
    """Block with token gates / buffers around QKV, projection and MLP (reference blocks.py:399-463)."""

    def __init__(self, gate_before_ln=False, stgt=False, **super_kwargs):
        super().__init__(**super_kwargs)
        self.gate_before_ln = gate_before_ln
        self.stgt = stgt
        token_gate_class = SimpleSTGTGate if stgt else TokenGate
        self.qkv_gate = token_gate_class()
        self.qkv_accumulator = TokenBuffer()
        self.projection_gate = token_gate_class()
        self.projection_accumulator = TokenBuffer()
        self.mlp_gate = token_gate_class()
        self.mlp_accumulator = TokenBuffer()

    # ------------------------------------------------------------------ gate site
    def _gate_site(self, gate, xa, xb, ln):
        """
        One gate site on an incremental frame.  Returns (x, c_tilde, index, count) where x = xa (+ xb) is the
        site input (materialised only when there is a residual to add) and count is None (all entries of index
        valid) or a device-side int32 tensor (threshold policy: index is padded to N, no host synchronisation).
        c_tilde is the gathered tensor, or a _Gathered record when the following linear layer gathers the rows
        itself and advances gate.p in the same kernel (16-bit models, built-in policies, LayerNorm before the gate).
        """
        ln_p = None if ln is None else self._ln_params(ln)
        pre, post = (None, ln_p) if self.gate_before_ln else (ln_p, None)
        if gate.count_mode:
            gate.counts["gate_flops"] += gate.p.numel()
        spec = _policy_spec(gate.policy, xa.shape[-2])
        count = None
        fuse = (FUSE_GATHER and spec is not None and post is None and not self.stgt and xa.dtype != torch.float32
                and xa.shape[-1] % 8 == 0 and gate.p.is_contiguous())
        c_all = torch.empty_like(xa) if (fuse and pre is not None) else None
        if spec is not None:
            if "threshold" in spec:
                assert xa.shape[0] == 1  # policies.py:25
                index, xsum, count = native.gate_select(xa, p=gate.p, xb=xb, want_sum=True, ln=pre, eps=LN_EPS,
                                                        device_count=True, c_out=c_all, **spec)
            else:
                index, xsum = native.gate_select(xa, p=gate.p, xb=xb, want_sum=True, ln=pre, eps=LN_EPS, c_out=c_all,
                                                 **spec)
            x = xsum if xb is not None else xa
        else:  # user-defined policy: materialise the error tensor and call it
            x = native.add(xa, xb) if xb is not None else xa
            c = x if pre is None else self._layer_norm_all(ln, x)
            index = gate.policy(native.sub(c, gate.p), dim=-1)
        index = index.contiguous()
        gate._last_sel = (index, count)  # selection trace (gate.last_index); references, no copy
        if fuse and index.shape[-1] > 0:
            return x, _Gathered(x if c_all is None else c_all, index, gate.p), index, count
        if post is not None:
            c_tilde, _ = native.gate_gather(x, index, p=gate.p, ln=post, eps=LN_EPS, ln_after=True,
                                            full_replace=self.stgt, count=count)
        else:
            c_tilde, _ = native.gate_gather(x, index, p=gate.p, ln=pre, eps=LN_EPS, full_replace=self.stgt, count=count)
        return x, c_tilde, index, count

    def _rows_selected(self, index, count):
        """Host-side number of selected rows, for the op counters only (reads the device count when there is one)."""
        if not self.count_mode:
            return None
        per_entry = index.shape[-1] if count is None else int(count.max().item())
        return per_entry * (index.numel() // max(1, index.shape[-1]))

    def _gate_first(self, gate, x, ln):
        """Frame 0 of a gate site: returns the dense site output and initialises gate.p."""
        gate.first = False
        if ln is None:
            gate.p = x
            return x
        c = self._layer_norm_all(ln, x)
        gate.p = x.clone() if self.gate_before_ln else c
        return c

    @staticmethod
    def _buffer_first(buffer, x):
        buffer.first = False
        buffer.b = x  # x is a fresh tensor owned by this block (the reference clones, modules.py:83)
        return x

    # ------------------------------------------------------------------ attention hooks
    def _attention_first(self, qkv, index):
        return self._dense_attention(qkv)

    def _attention_incremental(self, qkv, index, count=None):
        return self._dense_attention(qkv)

    # ------------------------------------------------------------------ pair protocol
    def forward(self, x):
        branch, skip = self._forward_pair(x.contiguous(), None)
        return native.add(branch, skip)

    def _forward_pair(self, xa, xb):
        self._check_input(xa)
        if self.qkv_gate.first:
            return self._first_pair(xa, xb)
        n = xa.shape[-2]
        # gate-accumulator 1: LN -> gate -> QKV -> buffer
        x, c1, index, count = self._gate_site(self.qkv_gate, xa, xb, self.input_layer_norm)
        qkv = self._linear(self.qkv, c1, out=self.qkv_accumulator.b, idx=index, count=count,
                           rows=self._rows_selected(index, count))
        attn = self._attention_incremental(qkv, index, count)
        x = self._ats_rows(x)  # adaptive token sampling: the skip keeps the sampled tokens only
        # gate-accumulator 2: gate -> projection -> buffer
        _, c2, index2, count2 = self._gate_site(self.projection_gate, attn, None, None)
        proj = self._linear(self.projection, c2, out=self.projection_accumulator.b, idx=index2, count=count2,
                            rows=self._rows_selected(index2, count2))
        # gate-accumulator 3: (+ skip) -> LN -> gate -> MLP -> buffer
        x2, c3, index3, count3 = self._gate_site(self.mlp_gate, proj, x, self.mlp_layer_norm)
        branch = self._mlp(c3, out=self.mlp_accumulator.b, idx=index3, count=count3, rows=self._rows_selected(index3, count3))
        self._count_adds(x2)
        assert self.ats_fraction is not None or n == x2.shape[-2]
        return branch, x2

    def _first_pair(self, xa, xb):
        x = native.add(xa, xb) if xb is not None else xa
        c1 = self._gate_first(self.qkv_gate, x, self.input_layer_norm)
        qkv = self._buffer_first(self.qkv_accumulator, self.qkv(c1))
        attn = self._attention_first(qkv, None)
        x = self._ats_rows(x)
        c2 = self._gate_first(self.projection_gate, attn, None)
        proj = self._buffer_first(self.projection_accumulator, self.projection(c2))
        x2 = native.add(proj, x)
        c3 = self._gate_first(self.mlp_gate, x2, self.mlp_layer_norm)
        branch = self._buffer_first(self.mlp_accumulator, self._mlp(c3))
        self._count_adds(x2)
        return branch, x2


class EventfulMatmul1Block(EventfulTokenwiseBlock):
    """
    Adds eventfulness to the query-key product (reference blocks.py:466-540).  The reference keeps the
    N x Nk product as state and refreshes k rows and k columns; since that state always equals
    (q / scale) k^T of the current QKV buffer, this implementation recomputes the logits tile by
    tile inside the attention kernels instead of storing them.  `matmul_accumulator_1.product` is still
    available: it is computed on demand from the block's current QKV buffer.
    """

    def __init__(self, **super_kwargs):
        super().__init__(**super_kwargs)
        if self.pool_size is not None:  # _pool_index assumes divisibility (reference blocks.py:479-482)
            assert all(s % p == 0 for s, p in zip(self.input_size, self.pool_size))
        assert self.window_size is None  # reference blocks.py:485
        self.matmul_accumulator_1 = MatmulBuffer()
        self.matmul_accumulator_1._producer = self._matmul_1_product

    # -- the state the reference stores, recomputed from the QKV buffer when somebody asks for it
    def _matmul_1_product(self):
        qkv = self.qkv_accumulator.b
        if qkv is None:
            return None
        b, n, _ = qkv.shape
        h, dh = self.heads, self.dim // self.heads
        q = qkv[..., : self.dim].reshape(b, n, h, dh).permute(0, 2, 1, 3)
        if self.pool_size is not None:
            kv = native.pool_kv(qkv, self._grid, self.pool_size)
            keys = kv[..., : self.dim].reshape(b, kv.shape[1], h, dh).permute(0, 2, 3, 1)
        else:
            keys = qkv[..., self.dim: 2 * self.dim].reshape(b, n, h, dh).permute(0, 2, 3, 1)
        return native.bmm(q, keys, alpha=1.0 / self.scale)

    def _pooled(self, qkv, index, count):
        """(pooled [k | v] tensor, pooled index, its device-side count) -- or (None, index, count) without pooling."""
        if self.pool_size is None:
            return None, index, count
        kvp = native.pool_kv(qkv, self._grid, self.pool_size)
        if index is None:
            return kvp, None, None
        index_k, count_k = native.pool_index(index, count, self._grid, self.pool_size)  # blocks.py:525-540
        return kvp, index_k, count_k

    def _count_matmul_1(self, qkv, n_keys, rows_q, cols_k):
        """Counters of MatmulBuffer: full product (rows_q None) or the row + column refresh (modules.py:236-247)."""
        if self.count_mode:
            b, n, _ = qkv.shape
            dh = self.dim // self.heads
            mm = self.matmul_accumulator_1.matmul.counts
            if rows_q is None:
                mm["matmul_flops"] += b * self.heads * n * n_keys * dh
            else:
                mm["matmul_flops"] += b * self.heads * (rows_q * n_keys + n * cols_k) * dh
            if self.relative_position is not None:
                self.relative_position.count_fused(b * self.heads, n, n_keys)

    def _global(self, qkv, mode, **kw):
        return native.global_attention(qkv, self.heads, self._grid, mode, rel=self._rel_tables(qkv.dtype),
                                       pool=self.pool_size, state_dtype=self._state_dtype(qkv.dtype), **kw)

    def _attention_first(self, qkv, index):
        self.matmul_accumulator_1.first = False
        return self._attention_incremental(qkv, None)

    def _attention_incremental(self, qkv, index, count=None):
        b, n, _ = qkv.shape
        if self.ats_fraction is not None:  # no pooling with ATS (asserted by Block)
            if self.count_mode:
                self._count_matmul_1(qkv, n, None if index is None else self._rows_selected(index, count) // b,
                                     None if index is None else self._rows_selected(index, count) // b)
            return self._ats_dense(qkv)
        kvp, index_k, count_k = self._pooled(qkv, index, count)
        n_keys = n if kvp is None else kvp.shape[1]
        if self.count_mode:
            if index is None:
                self._count_matmul_1(qkv, n_keys, None, None)
            else:
                self._count_matmul_1(qkv, n_keys, self._rows_selected(index, count) // b,
                                     self._rows_selected(index_k, count_k) // b)
            self.matmul.counts["matmul_flops"] += b * self.heads * n * n_keys * (self.dim // self.heads)
        return self._global(qkv, native.ATTN_DENSE, kv_pooled=kvp)


class EventfulBlock(EventfulMatmul1Block):
    """
    Also gates the attention-value product (reference blocks.py:543-575): v-gate and A-gate forced by
    the (pooled) QKV gate index, MatmulDeltaAccumulator update -- one fused kernel pair here.

    State (allocated at frame 0 in the matmul_2_cast dtype, exposed with the reference's logical shapes;
    Nk = N, or the pooled key count):
        v_gate.p                      (B, H, Nk, dh)  view of a (B, Nk, D) tensor
        matmul_gate.p                 (B, H, N, Nk)   view of the column-major (B, H, Nk, NP) A-state
        matmul_accumulator_2.product  (B, H, N, dh)   view of a (B, N, D) tensor
    """

    def __init__(self, **super_kwargs):
        super().__init__(**super_kwargs)
        self.v_gate = TokenDeltaGate()
        self.matmul_gate = TokenDeltaGate(structure="col")
        self.matmul_accumulator_2 = MatmulDeltaAccumulator()
        self._a_state = self._v_state = self._acc = self._stats = None

    def reset_self(self):
        super().reset_self()
        self._a_state = self._v_state = self._acc = self._stats = None

    def _attention_first(self, qkv, index):
        b, n, _ = qkv.shape
        d, h = self.dim, self.heads
        dh = d // h
        n_pad = (n + 7) // 8 * 8
        dev, sdt = qkv.device, self._state_dtype(qkv.dtype)
        kvp, _, _ = self._pooled(qkv, None, None)
        n_keys = n if kvp is None else kvp.shape[1]
        if self.ats_fraction is not None:
            # a and v are cast first, then sampled (reference blocks.py:561-562): the scores are formed in the state dtype;
            # from here on the query axis of the gate state / accumulator is the sampled one
            n = self._ats_select(qkv, sdt).shape[-1]
            n_pad = (n + 7) // 8 * 8
        self._a_state = torch.zeros((b, h, n_keys, n_pad), dtype=sdt, device=dev)
        self._v_state = torch.empty((b, n_keys, d), dtype=sdt, device=dev)
        self._acc = torch.empty((b, n, d), dtype=sdt, device=dev)
        self._stats = torch.empty((b, h, n, 2), dtype=torch.float32, device=dev)
        if self.ats_fraction is not None:
            out = native.global_attention_rows(qkv, self._ats_now, h, native.ATTN_FIRST, a_state=self._a_state,
                                               v_state=self._v_state, acc=self._acc, stats=self._stats, state_dtype=sdt)
        else:
            out = self._global(qkv, native.ATTN_FIRST, a_state=self._a_state, v_state=self._v_state, acc=self._acc,
                               stats=self._stats, kv_pooled=kvp)
        self.matmul_accumulator_1.first = False
        self.v_gate.first = self.matmul_gate.first = self.matmul_accumulator_2.first = False
        self.v_gate.p = self._v_state.view(b, n_keys, h, dh).permute(0, 2, 1, 3)
        self.matmul_gate.p = self._a_state[..., :n].transpose(-1, -2)
        self.matmul_accumulator_2.product = self._acc.view(b, n, h, dh).permute(0, 2, 1, 3)
        self._count_matmul_1(qkv, n_keys, None, None)
        if self.count_mode:
            self.matmul_accumulator_2.matmul.counts["matmul_flops"] += b * h * n * n_keys * dh
        return out

    def _attention_incremental(self, qkv, index, count=None):
        b, n, _ = qkv.shape
        h = self.heads
        dh = self.dim // h
        kvp, index_k, count_k = self._pooled(qkv, index, count)
        n_keys = n if kvp is None else kvp.shape[1]
        if self.count_mode:
            rows_q = self._rows_selected(index, count) // b
            cols_k = self._rows_selected(index_k, count_k) // b
            self._count_matmul_1(qkv, n_keys, rows_q, cols_k)
            nq = self._acc.shape[1]  # query rows of the gate state: N, or the sampled rows under ATS
            self.v_gate.counts["gate_flops"] += b * h * n_keys * dh
            self.matmul_gate.counts["gate_flops"] += b * h * nq * n_keys
            self.matmul_accumulator_2.counts["accumulator_flops"] += b * (h * cols_k * dh + 2 * h * nq * dh)
            self.matmul_accumulator_2.matmul.counts["matmul_flops"] += 2 * b * h * nq * cols_k * dh
        if self.ats_fraction is not None:
            self._ats_select(qkv, self._state_dtype(qkv.dtype))
            return native.global_attention_rows(qkv, self._ats_now, h, native.ATTN_DELTA, idx=index_k, count=count_k,
                                                a_state=self._a_state, v_state=self._v_state, acc=self._acc,
                                                stats=self._stats, state_dtype=self._state_dtype(qkv.dtype))
        return self._global(qkv, native.ATTN_DELTA, idx=index_k, count=count_k, a_state=self._a_state,
                            v_state=self._v_state, acc=self._acc, stats=self._stats, kv_pooled=kvp)
