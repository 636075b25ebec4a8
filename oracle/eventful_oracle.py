"""
CPU oracle for the gated sparse-token update path of Eventful Transformers.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this file, and there only
as the checker or as the timed CPU baseline.  The product path (the
``eventful_transformer`` package under ``eventful-transformer_b200/``) never
imports it and has no CPU fallback.

What this is: a *functional* restatement (plain functions over explicit state
dicts, torch CPU tensors, no nn.Module) of the algorithm implemented by the
reference in ``eventful_transformer/{modules,policies,blocks,backbones,utils}.py``.
Every function cites the reference ``file:line`` it follows (paths relative to
the reference checkout).  The arithmetic on the path is PyTorch ATen itself
(``vector_norm``, ``softmax``, ``linear``, ``layer_norm``, ``gelu`` ...), which
is present in this image, so the oracle calls the same ATen ops in the same
order and is bit-identical to the reference on CPU whenever the selected index
sets agree (they do whenever no two candidate norms tie at the k-th value).

Parity pinning: ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference`` in the build container, runs it on seeded inputs and commits
the outputs as fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this file against those fixtures (bitwise for activations in fp32,
set-equality for indices).  The reference ships no tests / golden vectors of
its own.

One deliberate refinement: the reference's ``topk(sorted=False)`` leaves order
and tie-breaking unspecified (policies.py:63).  The oracle pins the rule used by
torch's CUDA radix select -- all elements strictly greater than the k-th value in
ascending index order, then elements equal to it in ascending index order --
because the GPU box's ``torch.topk`` is the reference behaviour the CUDA path
has to reproduce.  ``tests/test_gate_gpu.py`` checks that
rule against ``torch.topk`` on the B200 with crafted ties.
"""

from math import prod, sqrt

import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # blocks.py:23

DENSE = "Block"
TOKENWISE = "EventfulTokenwiseBlock"
MATMUL1 = "EventfulMatmul1Block"
EVENTFUL = "EventfulBlock"


# --------------------------------------------------------------------------
# Policies (policies.py)
# --------------------------------------------------------------------------


def token_norm(e, dim=-1):
    """Per-token L2 error norm, result in e.dtype (policies.py:28,63,93)."""
    return torch.linalg.vector_norm(e, ord=2, dim=dim)


def select_topk(norm, k):
    """
    Indices of the k largest norms along the last dim (policies.py:63), with
    torch-CUDA radix-select order: strictly-greater-than-kth ascending, then
    equal-to-kth ascending.  NaN norms are not modelled (never produced by the
    path on finite inputs).
    """
    n = norm.shape[-1]
    if k > n:
        raise RuntimeError("selected index k out of range")  # torch.topk raises
    flat = norm.reshape(-1, n).float()
    out = torch.empty((flat.shape[0], k), dtype=torch.int64)
    for r in range(flat.shape[0]):
        row = flat[r]
        if k == 0:
            continue
        kth = torch.topk(row, k, sorted=True)[0][-1]
        above = torch.nonzero(row > kth).flatten()
        equal = torch.nonzero(row == kth).flatten()
        out[r] = torch.cat([above, equal[: k - above.numel()]])
    return out.view(tuple(norm.shape[:-1]) + (k,))


def select_threshold(norm, threshold):
    """Ascending indices with norm > threshold; batch must be 1 (policies.py:20-32)."""
    assert all(s == 1 for s in norm.shape[:-1])
    index = torch.nonzero(norm.reshape(-1) > threshold).flatten()
    return index.view((1,) * (norm.ndim - 1) + (-1,))


def make_policy(kind, **kw):
    """Returns f(e, dim) -> int64 index, mirroring the three reference policies."""
    if kind == "topk":  # policies.py:39-68
        return lambda e, dim=-1: select_topk(token_norm(e, dim), kw["k"])
    if kind == "threshold":  # policies.py:6-36
        return lambda e, dim=-1: select_threshold(token_norm(e, dim), kw["threshold"])
    if kind == "fraction":  # policies.py:71-95

        def _fraction(e, dim=-1):
            nrm = token_norm(e, dim)
            return select_topk(nrm, int(kw["fraction"] * nrm.shape[-1]))

        return _fraction
    raise ValueError(kind)


# --------------------------------------------------------------------------
# Index plumbing (eventful_transformer/utils.py:198-211)
# --------------------------------------------------------------------------


def _rows(index, shape):
    """Index view for gather/scatter along dim=-2 (utils.py:206-211)."""
    lead = index.shape[:-1]
    extra = len(shape) - index.ndim
    v = index.view(lead + (1,) * (extra - 1) + (index.shape[-1], 1))
    return v.expand(tuple(shape[:-2]) + (-1, shape[-1]))


def _cols(index, shape):
    """Index view for gather/scatter along dim=-1 (utils.py:198-203)."""
    lead = index.shape[:-1]
    extra = len(shape) - index.ndim
    v = index.view(lead + (1,) * extra + (index.shape[-1],))
    return v.expand(tuple(shape[:-1]) + (-1,))


def _take(x, index, structure):
    if structure == "row":
        return x.gather(-2, _rows(index, x.shape))
    return x.gather(-1, _cols(index, x.shape))


def _put(dst, index, src, structure):
    if structure == "row":
        dst.scatter_(-2, _rows(index, dst.shape), src)
    else:
        dst.scatter_(-1, _cols(index, dst.shape), src)


# --------------------------------------------------------------------------
# Gating primitives (modules.py).  State is a dict; an empty dict == "first".
# --------------------------------------------------------------------------


def token_gate(st, c, policy=None, forced_index=None, structure="row", delta=False):
    """
    TokenGate (modules.py:104-168) / TokenDeltaGate (modules.py:171-201).
    Returns (c_tilde, index) or, with delta=True, (c_tilde, e_tilde, index).
    The reference state p advances only at the selected positions (:151,:200).
    """
    if "p" not in st:  # forward_first, modules.py:135-141 / :182-184
        st["p"] = c  # alias, exactly like the reference
        return (c, None, None) if delta else (c, None)
    e = c - st["p"]  # :149 / :196
    if forced_index is None:  # _apply_policy :154-164
        index = policy(e, dim=(-1 if structure == "row" else -2))
    else:
        index = forced_index
    c_tilde = _take(c, index, structure)  # :150 / :198
    e_tilde = _take(e, index, structure) if delta else None  # :199
    _put(st["p"], index, c_tilde, structure)  # :151 / :200
    return (c_tilde, e_tilde, index) if delta else (c_tilde, index)


def stgt_gate(st, c, policy):
    """SimpleSTGTGate (modules.py:6-49): the whole reference is replaced each step."""
    if "p" not in st:
        st["p"] = c
        return c, None
    index = policy(c - st["p"], dim=-1)  # :42
    c_tilde = _take(c, index, "row")  # :43
    st["p"] = c  # :44
    return c_tilde, index


def token_buffer(st, x, index, structure="row"):
    """TokenBuffer (modules.py:52-101).  Returns the state tensor itself."""
    if "b" not in st:
        st["b"] = x.clone()  # :83
        return st["b"]
    _put(st["b"], index, x, structure)  # :90-96
    return st["b"]


def matmul_buffer(st, q, k, index_q, index_k):
    """
    MatmulBuffer (modules.py:204-252): query-key product with row then column
    refresh.  q is (..., N, dh) already divided by scale, k is (..., dh, N).
    """
    if "product" not in st:
        st["product"] = q @ k  # :228
        return st["product"]
    q_tilde = _take(q, index_q, "row")  # :236
    k_tilde = _take(k, index_k, "col")  # :237
    _put(st["product"], index_q, q_tilde @ k, "row")  # :238-242
    _put(st["product"], index_k, q @ k_tilde, "col")  # :243-247
    return st["product"]


def delta_accumulator(st, a_n, v_n, a_delta, v_delta):
    """MatmulDeltaAccumulator (modules.py:255-299)."""
    if "product" not in st:
        st["product"] = a_n @ v_n  # :281
        return st["product"]
    st["product"] += a_n @ v_delta  # :293
    st["product"] += a_delta @ (v_n - v_delta)  # :294
    return st["product"]


# --------------------------------------------------------------------------
# Position / relative-position tables (eventful_transformer/utils.py)
# --------------------------------------------------------------------------


def sized_position_encoding(encoding, encoding_size, input_size, has_class_token):
    """PositionEncoding._compute_sized_encoding (utils.py:69-100)."""
    encoding_size, input_size = tuple(encoding_size), tuple(input_size)
    if input_size == encoding_size:
        return encoding
    cls = None
    if has_class_token:
        cls, encoding = encoding[:, :1], encoding[:, 1:]
    e = encoding.transpose(1, 2)
    e = e.view(e.shape[:-1] + encoding_size)
    e = F.interpolate(e, input_size, mode="bicubic", align_corners=False)
    e = e.flatten(start_dim=2).transpose(1, 2)
    if has_class_token:
        e = torch.concat([cls, e], dim=1)
    return e


def relative_table(embedding, embedding_size, attention_size, dim, pool=None):
    """RelativePositionEmbedding._get_relative (utils.py:173-189); `pool` averages the key axis (utils.py:185-188)."""
    s = embedding_size[dim]
    r0 = torch.arange(s).unsqueeze(1)
    r1 = torch.arange(s).unsqueeze(0)
    rel = embedding[r0 - r1 + s - 1]
    if tuple(embedding_size) != tuple(attention_size):
        rel = rel.transpose(0, 2).unsqueeze(0)
        rel = F.interpolate(rel, tuple(attention_size), mode="bicubic", align_corners=False)
        rel = rel.squeeze(0).transpose(0, 2)
    if pool is not None:
        rel = F.avg_pool1d(rel.transpose(1, 2), pool[dim]).transpose(1, 2)
    return rel


def add_relative_position(x, q, y_rel, x_rel, attention_size, inplace, key_size=None):
    """
    RelativePositionEmbedding.forward (utils.py:139-171): decomposed rel-pos
    bias computed from the *unscaled* q and added to the logits x (B,H,N,Nk);
    `key_size` is the pooled key grid when K/V pooling is on (utils.py:143-147).
    """
    a = tuple(attention_size)
    p = a if key_size is None else tuple(key_size)
    xs = x.view(x.shape[:2] + a + p)
    qs = q.view(q.shape[:2] + a + q.shape[-1:])
    t = torch.einsum("abhwc,hkc->abhwk", qs, y_rel).unsqueeze(-1)
    if inplace:
        xs += t
    else:
        xs = xs + t
    xs += torch.einsum("abhwc,wkc->abhwk", qs, x_rel).unsqueeze(-2)
    return xs.view(xs.shape[:2] + (prod(a), prod(p)))


# --------------------------------------------------------------------------
# Block-level restatement (blocks.py)
# --------------------------------------------------------------------------


def _pad_concat(x, size, pad_tensor):
    """utils/image.py:31-49 pad_to_size: concat-based right/bottom padding."""
    for dim in range(-1, -len(size) - 1, -1):
        shape = list(x.shape)
        shape[dim] = size[dim] - x.shape[dim]
        if shape[dim] == 0:
            continue
        x = torch.concat([x, pad_tensor.expand(shape)], dim)
    return x


class OracleBackbone:
    """
    ViTBackbone (backbones.py:8-64) + Block / EventfulTokenwiseBlock /
    EventfulMatmul1Block / EventfulBlock (blocks.py:26-575) as plain functions
    over a parameter dict that uses the reference's state-dict key names and a
    per-block state dict.  K/V pooling (blocks.py:303-326,525-540) is restated
    (`pool_size`), and so is adaptive token sampling (`ats_fraction`, blocks.py:150-181,196-203,378-391) with the
    reference's exact axis handling: the per-head scores are summed over the BATCH axis and row h of the resulting index
    serves batch entry h, so it is only defined -- in the reference too -- when batch == heads.
    """

    def __init__(
        self,
        params,
        depth,
        dim,
        heads,
        input_size,
        position_encoding_size,
        mlp_ratio=4,
        block_class=DENSE,
        windowed_class=None,
        window_indices=(),
        window_size=None,
        relative_embedding_size=None,
        has_class_token=False,
        matmul_2_cast=None,
        windowed_matmul_2_cast="same",
        gate_before_ln=False,
        stgt=False,
        pool_size=None,
        windowed_pool_size="same",
        ats_fraction=None,
    ):
        self.w = params
        self.depth, self.dim, self.heads = depth, dim, heads
        self.input_size = tuple(input_size)
        self.position_encoding_size = tuple(position_encoding_size)
        self.has_class_token = has_class_token
        self.scale = sqrt(dim // heads)  # blocks.py:92
        self.gate_before_ln = gate_before_ln
        self.stgt = stgt
        self.blocks = []
        for i in range(depth):  # backbones.py:47-59
            windowed = i in window_indices
            cls = (windowed_class or block_class) if windowed else block_class
            cast = matmul_2_cast
            if windowed and windowed_matmul_2_cast != "same":
                cast = windowed_matmul_2_cast
            ws = tuple(window_size) if (windowed and window_size is not None) else None
            rel = None
            if relative_embedding_size is not None:
                rel = ws if ws is not None else tuple(relative_embedding_size)  # blocks.py:86-91
            pool = pool_size
            if windowed and windowed_pool_size != "same":
                pool = windowed_pool_size
            pool = None if pool is None else tuple(pool)
            if ats_fraction is not None:  # blocks.py:71-74
                assert pool is None and ws is None and 0.0 <= ats_fraction <= 1.0
            self.blocks.append(dict(cls=cls, window=ws, rel=rel, cast=cast, pool=pool, ats=ats_fraction))
        self.policy = None
        self.record_free = False
        self.reset()

    # -- control API (base.py:130-135, utils/misc.py:140-143) ----------------
    def reset(self):
        self.state = [dict() for _ in range(self.depth)]
        self._pos = None
        self._rel = {}
        self.trace = []
        self.free_trace = []

    def set_policy(self, kind, **kw):
        self.policy = make_policy(kind, **kw)

    # -- helpers --------------------------------------------------------------
    def _p(self, i, name):
        return self.w[f"blocks.{i}.{name}"]

    def _st(self, i, name):
        return self.state[i].setdefault(name, {})

    def _linear(self, i, name, x):
        return F.linear(x, self._p(i, name + ".weight"), self._p(i, name + ".bias"))

    def _ln(self, i, name, x):
        return F.layer_norm(
            x, (self.dim,), self._p(i, name + ".weight"), self._p(i, name + ".bias"), LN_EPS
        )

    def _gate(self, i, name, c, forced=None):
        """A policy-driven row gate; records / replays the selected index."""
        st = self._st(i, name)
        if self.stgt:
            out = stgt_gate(st, c, self.policy)
        else:
            f = None
            if forced is not None and "p" in st:
                f = forced.get((i, name))
            if f is not None and self.record_free:
                # checker aid: what the policy WOULD select here on the oracle's own (replayed-history) inputs,
                # plus the norms, so a test can show that a CUDA selection differs only at near-ties
                e = c - st["p"]
                self.free_trace.append(((i, name), self.policy(e, dim=-1), token_norm(e, dim=-1)))
            out = token_gate(st, c, policy=self.policy, forced_index=f)
        if out[1] is not None:
            self.trace.append(((i, name), out[1]))
        return out

    # -- windows / heads (blocks.py:248-301, 329-376) ---------------------------
    def _partition_windows(self, i, x):
        ws = self.blocks[i]["window"]
        if ws is None:
            return x
        h, w = self.input_size
        pad = (-h % ws[0], -w % ws[1])
        x = x.view(x.shape[:1] + self.input_size + x.shape[2:])
        if any(pad):
            s = x.shape
            fill = torch.zeros((1,) * (x.ndim - 1) + s[-1:], dtype=x.dtype)
            fill = fill + self._p(i, "qkv.bias")  # counting.py:146-150 forward_bias
            x = _pad_concat(x, (s[-3] + pad[0], s[-2] + pad[1], s[-1]), fill)
        s = x.shape
        x = x.view(-1, s[-3] // ws[0], ws[0], s[-2] // ws[1], ws[1], s[-1])
        x = x.transpose(-3, -4)
        return x.reshape(-1, prod(ws), s[-1])

    def _recombine_windows(self, i, x):
        ws = self.blocks[i]["window"]
        if ws is None:
            return x
        h, w = self.input_size
        th, tw = h + (-h % ws[0]), w + (-w % ws[1])
        x = x.view(-1, th // ws[0], tw // ws[1], ws[0], ws[1], x.shape[-1])
        x = x.transpose(-3, -4).reshape(-1, th, tw, x.shape[-1])
        if (th, tw) != (h, w):
            x = x[:, :h, :w]
        return x.flatten(start_dim=1, end_dim=2)

    def _heads(self, x):
        x = x.view(x.shape[:-1] + (3, self.heads, x.shape[-1] // (3 * self.heads)))
        return x.permute(2, 0, 3, 1, 4)  # q, k, v each (B,H,N,dh)

    @staticmethod
    def _merge_heads(x):
        x = x.permute(0, 2, 1, 3)
        return x.reshape(x.shape[:-2] + (-1,))

    def _relpos(self, i, x, q, inplace):
        blk = self.blocks[i]
        if blk["rel"] is None:
            return x
        att = blk["window"] if blk["window"] is not None else self.input_size
        pool = blk["pool"]
        if i not in self._rel:  # cached until reset (utils.py:151-156,186-191)
            self._rel[i] = (
                relative_table(self._p(i, "relative_position.y_embedding"), blk["rel"], att, 0, pool),
                relative_table(self._p(i, "relative_position.x_embedding"), blk["rel"], att, 1, pool),
            )
        keys = None if pool is None else (att[0] // pool[0], att[1] // pool[1])
        return add_relative_position(x, q, self._rel[i][0], self._rel[i][1], att, inplace, keys)

    def _pool_tokens(self, i, x):
        """Block._pool_tokens (blocks.py:303-326): average-pools k / v over the token grid."""
        blk = self.blocks[i]
        if blk["pool"] is None:
            return x
        w = blk["window"] if blk["window"] is not None else self.input_size
        s = x.shape
        x = x.reshape((-1,) + tuple(w) + s[-1:]).permute(0, 3, 1, 2)
        x = F.avg_pool2d(x, blk["pool"]).permute(0, 2, 3, 1)
        return x.reshape(s[:-2] + (-1,) + s[-1:])

    def _pool_index(self, i, index):
        """EventfulMatmul1Block._pool_index (blocks.py:525-540): token index -> sorted unique pooled-cell index."""
        pool = self.blocks[i]["pool"]
        if pool is None or index is None:
            return index
        width = self.input_size[1]
        iy = index.div(width, rounding_mode="floor").div(pool[0], rounding_mode="floor")
        ix = index.remainder(width).div(pool[1], rounding_mode="floor")
        return (iy * (width // pool[1]) + ix).unique(dim=-1)

    @staticmethod
    def _cast(i_cast, a, v):
        old = a.dtype
        if i_cast is not None:  # blocks.py:183-189
            dt = getattr(torch, i_cast)
            a, v = a.to(dt), v.to(dt)
        return a, v, old

    # -- adaptive token sampling ------------------------------------------------------
    def _ats(self, i, a, v, forced=None):
        """
        Block._adaptive_token_sampling (blocks.py:150-181): the top-k variant of ATS.  a: (B, H, N, N) attention
        probabilities, v: (B, H, N, dh).  Returns (rows of `a` kept, index (B', n_select)) or (a, None).
        score[b, h, t] = a[b, h, t, 0] * |v[b, h, t]|, divided by its sum over t >= 1 (:155-157); the class token always
        wins (:160); the scores are then summed over axis -3 (:163) -- for the (B, H, N) score tensor of a 3-D block
        input that is the batch axis -- and the n_select best tokens per remaining row are kept, sorted and stabilised
        against the previous frame's set (:168-176).  The gather at :179 broadcasts index row r to batch entry r, which
        requires the index to have B rows, i.e. B == H (torch raises otherwise, and so does this restatement).
        `forced[(i, "ats")]`: replay a given stabilised index instead of selecting (parity at identical index sets).
        """
        frac = self.blocks[i]["ats"]
        if frac is None:
            return a, None
        st = self._st(i, "ats")
        replay = forced is not None and (i, "ats") in forced
        if not replay or self.record_free:
            raw = a[..., 0] * torch.linalg.vector_norm(v, dim=-1)
            score = raw / raw[..., 1:].sum(dim=-1, keepdim=True)
            score[..., 0] = float("inf")
            score = score.sum(dim=-3)
            n_select = int(frac * (score.shape[-1] - 1)) + 1  # :165
            free = score.topk(n_select, sorted=False)[1].sort(dim=-1)[0]  # :168, :379
        if replay:
            index = forced[(i, "ats")].clone()
            if self.record_free:  # checker aid: the set the oracle would sample here, and the scores it is based on
                self.ats_free.append((i, free, score))
        else:
            index = self._stabilize(st.get("last"), free)
        st["last"] = index
        self.trace.append(((i, "ats"), index))
        if index.shape[:-1] != a.shape[:1]:
            raise RuntimeError(f"ATS index has {tuple(index.shape[:-1])} rows for a batch of {a.shape[0]} "
                               "(the reference's gather, blocks.py:179, needs batch == heads)")
        rows = index.view(index.shape[0], 1, index.shape[-1], 1).expand(-1, a.shape[1], -1, a.shape[-1])
        return a.gather(-2, rows), index

    @staticmethod
    def _stabilize(last, index):
        """Block._stabilize_ats_indices (blocks.py:378-391): tokens that stay keep their slot of the previous frame's
        index; the slots of tokens that left are filled, in order, with the tokens that entered."""
        if last is None:
            return index
        out = last.clone()
        for r in range(index.shape[0]):
            left = ~torch.isin(last[r], index[r])
            entered = ~torch.isin(index[r], last[r])
            out[r, left] = index[r, entered]
        return out

    @staticmethod
    def _ats_skip(skip, index):
        """Block._gather_ats_skip (blocks.py:196-203)."""
        if index is None:
            return skip
        return skip.gather(-2, index.unsqueeze(-1).expand(-1, -1, skip.shape[-1]))

    # -- attention variants --------------------------------------------------------
    def _attention_dense(self, i, x, forced=None):
        """Block._forward_attention (blocks.py:205-240)."""
        x = self._partition_windows(i, x)
        q, k, v = self._heads(x)
        k, v = self._pool_tokens(i, k), self._pool_tokens(i, v)  # :216-217
        a = (q / self.scale) @ k.transpose(-2, -1)  # :223
        a = self._relpos(i, a, q, inplace=True)  # :225
        a = a.softmax(dim=-1)  # :226
        a, ats = self._ats(i, a, v, forced)  # :229
        a, v, old = self._cast(self.blocks[i]["cast"], a, v)  # :231
        x = a @ v  # :232
        x = self._merge_heads(x)
        x = self._recombine_windows(i, x)
        return (x.to(old) if self.blocks[i]["cast"] is not None else x), ats

    def _matmul_1(self, i, x, index):
        """EventfulMatmul1Block._forward_matmul_1 (blocks.py:506-523)."""
        q, k, v = self._heads(x)
        k, v = self._pool_tokens(i, k), self._pool_tokens(i, v)  # :509-510
        index_k = self._pool_index(i, index)  # :511
        a = matmul_buffer(
            self._st(i, "matmul_accumulator_1"), q / self.scale, k.transpose(-2, -1), index, index_k
        )
        a = self._relpos(i, a, q, inplace=False)  # :521
        return a.softmax(dim=-1), v, index_k

    def _attention_matmul1(self, i, x, index, forced=None):
        """EventfulMatmul1Block._forward_attention (blocks.py:497-504)."""
        a, v, _ = self._matmul_1(i, x, index)
        a, ats = self._ats(i, a, v, forced)  # :499
        a, v, old = self._cast(self.blocks[i]["cast"], a, v)
        x = self._merge_heads(a @ v)
        return (x.to(old) if self.blocks[i]["cast"] is not None else x), ats

    def _attention_eventful(self, i, x, index, forced=None):
        """EventfulBlock._forward_attention (blocks.py:558-575)."""
        a, v, index_k = self._matmul_1(i, x, index)
        cast = self.blocks[i]["cast"]
        a, v, old = self._cast(cast, a, v)  # :561
        a, ats = self._ats(i, a, v, forced)  # :562 (after the cast: the scores are computed in the cast dtype)
        if not cast:
            v = v.clone()  # :563-566
        v_n, v_d, index_v = token_gate(self._st(i, "v_gate"), v, forced_index=index_k, delta=True)
        a_n, a_d, _ = token_gate(
            self._st(i, "matmul_gate"), a, forced_index=index_v, structure="col", delta=True
        )
        x = delta_accumulator(self._st(i, "matmul_accumulator_2"), a_n, v_n, a_d, v_d)  # :569
        x = self._merge_heads(x)  # :573
        return (x.to(old) if cast is not None else x), ats

    # -- blocks ----------------------------------------------------------------------
    def _block_dense(self, i, x, forced=None):
        """Block.forward (blocks.py:117-137)."""
        skip = x
        x = self._linear(i, "qkv", self._ln(i, "input_layer_norm", x))
        x, ats = self._attention_dense(i, x, forced)
        skip = self._ats_skip(skip, ats)  # :126
        x = self._linear(i, "projection", x) + skip
        skip = x
        x = self._linear(i, "mlp_1", self._ln(i, "mlp_layer_norm", x))
        x = self._linear(i, "mlp_2", F.gelu(x))
        return x + skip

    def _block_gated(self, i, x, forced):
        """
        EventfulTokenwiseBlock.forward (blocks.py:422-463), and the two
        subclasses' forward (blocks.py:489-495) which only swap the attention.
        """
        cls = self.blocks[i]["cls"]
        skip = x
        if self.gate_before_ln:  # :456-458
            x, index = self._gate(i, "qkv_gate", x, forced)
            x = self._ln(i, "input_layer_norm", x)
        else:  # :459-461
            x = self._ln(i, "input_layer_norm", x)
            x, index = self._gate(i, "qkv_gate", x, forced)
        x = self._linear(i, "qkv", x)  # :462
        x = token_buffer(self._st(i, "qkv_accumulator"), x, index)  # :424
        if cls == TOKENWISE:
            x, ats = self._attention_dense(i, x, forced)
        elif cls == MATMUL1:
            x, ats = self._attention_matmul1(i, x, index, forced)
        else:
            x, ats = self._attention_eventful(i, x, index, forced)
        skip = self._ats_skip(skip, ats)  # :426, :493
        x, index = self._gate(i, "projection_gate", x, forced)  # :432
        x = self._linear(i, "projection", x)
        x = token_buffer(self._st(i, "projection_accumulator"), x, index)
        x = x + skip  # :436
        skip = x
        if self.gate_before_ln:  # :440-445
            x, index = self._gate(i, "mlp_gate", x, forced)
            x = self._ln(i, "mlp_layer_norm", x)
        else:
            x = self._ln(i, "mlp_layer_norm", x)
            x, index = self._gate(i, "mlp_gate", x, forced)
        x = self._linear(i, "mlp_1", x)  # blocks.py:242-246
        x = self._linear(i, "mlp_2", F.gelu(x))
        x = token_buffer(self._st(i, "mlp_accumulator"), x, index)  # :447
        return x + skip  # :448

    # -- backbone ----------------------------------------------------------------------
    def forward(self, x, forced=None):
        """
        ViTBackbone.forward (backbones.py:61-64).  `forced` optionally maps
        (block, gate_name) -> int64 index to replay a selection trace (used to
        compare activations given identical index sets).
        """
        self.trace = []
        self.free_trace = []
        self.ats_free = []
        if self._pos is None:  # utils.py:53-67
            self._pos = sized_position_encoding(
                self.w["position_encoding.encoding"],
                self.position_encoding_size,
                self.input_size,
                self.has_class_token,
            )
        x = x + self._pos
        for i in range(self.depth):
            if self.blocks[i]["cls"] == DENSE:
                x = self._block_dense(i, x, forced)
            else:
                x = self._block_gated(i, x, forced)
        return x


# --------------------------------------------------------------------------
# Operation counters (base.py:7-78; counting.py; modules.py:41,148,195,290-292)
# --------------------------------------------------------------------------


def incremental_counts(n, k, dim, heads, mlp_ratio, cls, window=None, grid=None, rel=False, batch=1):
    """
    Closed-form MAC counters of ONE incremental frame of one block, matching
    the reference's Counted* modules (verified against the real counters in
    tests/golden/make_golden.py).  Returns a dict keyed like base.Counts.
    """
    dh = dim // heads
    c = dict(gate_flops=0, accumulator_flops=0, linear_flops=0, bias_flops=0, matmul_flops=0,
             add_flops=0, einsum_flops=0)
    # three token gates + three linears on k rows + MLP
    c["gate_flops"] += 3 * n * dim
    outs = [3 * dim, dim, mlp_ratio * dim, dim]
    ins = [dim, dim, dim, mlp_ratio * dim]
    for fi, fo in zip(ins, outs):
        c["linear_flops"] += k * fi * fo
        c["bias_flops"] += k * fo
    c["add_flops"] += 2 * n * dim  # two residual adds
    if window is not None:
        th = grid[0] + (-grid[0] % window[0])
        tw = grid[1] + (-grid[1] % window[1])
        nw = (th // window[0]) * (tw // window[1])
        w2 = window[0] * window[1]
        c["matmul_flops"] += 2 * nw * heads * w2 * w2 * dh
        if (th, tw) != tuple(grid):
            c["bias_flops"] += 3 * dim  # forward_bias on the 1x..x3D pad token
        if rel:
            c["einsum_flops"] += nw * heads * w2 * dh * (window[0] + window[1])
            c["add_flops"] += 2 * nw * heads * w2 * w2
    else:
        if cls == EVENTFUL:
            c["matmul_flops"] += 2 * heads * n * k * dh  # matmul 1: rows then cols
            c["gate_flops"] += heads * n * dh + heads * n * n  # v gate + A gate
            c["matmul_flops"] += 2 * heads * n * k * dh  # two delta products
            c["accumulator_flops"] += heads * k * dh + 2 * heads * n * dh
        elif cls == MATMUL1:
            c["matmul_flops"] += 2 * heads * n * k * dh + heads * n * n * dh
        else:
            c["matmul_flops"] += 2 * heads * n * n * dh
        if rel:
            c["einsum_flops"] += heads * n * dh * (grid[0] + grid[1])
            c["add_flops"] += 2 * heads * n * n
    return {key: v * batch for key, v in c.items()}
