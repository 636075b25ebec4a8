"""
Gating primitives: TokenGate, TokenDeltaGate, TokenBuffer, MatmulBuffer, MatmulDeltaAccumulator,
SimpleSTGTGate (API mirror of the reference's modules.py).

State lives in plain attributes (`first`, `p`, `b`, `product`), not registered buffers, and the
aliasing contract of the reference is kept: buffers / accumulators return their state tensor, a
gate's `p` is the first input it saw (modules.py:70-71,125-126,216-217,267-268).

Each module works stand-alone through the generic kernels below.  The Eventful blocks do not call
these forward() methods on their steady-state path; they drive the same state tensors through
fused kernels (blocks.py of this package) and keep `first` / `p` / `b` consistent.
"""

from eventful_transformer import _native as native
from eventful_transformer.base import ExtendedModule
from eventful_transformer.counting import CountedMatmul


def _policy_spec(policy, n_tokens):
    """Fused-selection parameters of a policy module, or None if it must be called as a black box."""
    spec = getattr(policy, "fused_spec", None)
    return None if spec is None else spec(n_tokens)


def _broadcast_index(index, target):
    """(B, k) index -> (leading dims of target..., k) contiguous (expand_row_index semantics)."""
    lead = tuple(target.shape[:-2])
    if tuple(index.shape[:-1]) == lead:
        return index.contiguous()
    extra = len(lead) - (index.ndim - 1)
    view = index.view(index.shape[:-1] + (1,) * extra + index.shape[-1:])
    return view.expand(lead + (index.shape[-1],)).contiguous()


class _FrameState(ExtendedModule):
    """
    Shared skeleton of every gating primitive: `first` tells frame 0 from the incremental frames and
    `_state` names the attributes that hold the per-stream tensors.  reset_self() (called by
    ExtendedModule.reset, base.py) drops them, so the next call is a first frame again.
    """

    _state = ()

    def __init__(self):
        super().__init__()
        self._drop_state()

    def _drop_state(self):
        self.first = True
        self._last_sel = None
        for name in self._state:
            setattr(self, name, None)

    # Selection trace of the fused block path (tests / visualisation): the index tensor the gate last used.  With a
    # device-side count (threshold policy) the padded index is cut to its valid length here, on demand -- the only
    # place that reads the count back to the host.
    @property
    def last_index(self):
        if self._last_sel is None:
            return None
        index, count = self._last_sel
        return index if count is None else index[..., : int(count.max().item())]

    def reset_self(self):
        self._drop_state()

    def _tally(self, key, amount):
        if self.count_mode:
            self.counts[key] += amount


class _GateBase(_FrameState):
    """
    Row / column gate against a reference tensor p (modules.py:104-201).  `_with_delta` selects the
    TokenDeltaGate flavour, whose results carry the gathered error e~ as well.
    """

    _state = ("p",)
    _with_delta = False

    def __init__(self, structure="row"):
        assert structure in ("row", "col")
        super().__init__()
        self.structure = structure
        self.policy = None

    # -- selection and state advance through the gate kernels
    def _select(self, c, p):
        along_rows = self.structure == "row"
        spec = _policy_spec(self.policy, c.shape[-2]) if along_rows else None
        if spec is None:  # user-defined policy (or column structure): materialise the error tensor
            return self.policy(native.sub(c, p), dim=-1 if along_rows else -2)
        if "threshold" in spec:
            assert all(size == 1 for size in c.shape[:-2])  # policies.py:25
        return native.gate_select(c, p=p, **spec)[0]

    def _update(self, c, index):
        if self.structure == "row":
            return native.gate_gather(c, _broadcast_index(index, c), p=self.p, want_delta=self._with_delta)
        return native.gate_gather_cols(c, index.contiguous(), p=self.p, want_delta=self._with_delta)

    def _pack(self, c_tilde, e_tilde, index):
        return (c_tilde, e_tilde, index) if self._with_delta else (c_tilde, index)

    # -- public protocol of the reference
    def forward(self, c, forced_index=None):
        """Note: p aliases the first input ever seen and is advanced in place afterwards."""
        return self.forward_first(c) if self.first else self.forward_incremental(c, forced_index=forced_index)

    def forward_first(self, c):
        # p aliases the first input (modules.py:140) when that input is dense; a strided view (e.g. the v-gate's
        # clone of a head-partitioned tensor, blocks.py:566) is materialised once, because the gate kernels address
        # the state as a dense (..., N, D) array
        self.first, self.p = False, (c if c.is_contiguous() else c.contiguous())
        return self._pack(c, None, None)

    def forward_incremental(self, c, forced_index=None):
        self._tally("gate_flops", self.p.numel())
        c = c.contiguous()
        index = forced_index if forced_index is not None else self._select(c, self.p)
        c_tilde, e_tilde = self._update(c, index)
        return self._pack(c_tilde, e_tilde, index)


class TokenGate(_GateBase):
    """Selects the tokens whose value drifted from the reference state p; returns (c~, index)."""


class TokenDeltaGate(_GateBase):
    """TokenGate that also returns the gathered error: (c~, e~, index) (modules.py:171-201)."""

    _with_delta = True


class SimpleSTGTGate(_FrameState):
    """Baseline gate of "Spatio-Temporal Gated Transformers": p is replaced wholesale each frame (modules.py:6-49)."""

    _state = ("p",)

    def __init__(self, structure="row"):
        assert structure == "row"
        super().__init__()
        self.policy = None

    def forward(self, c):
        if self.first:
            self.first, self.p = False, (c if c.is_contiguous() else c.contiguous())
            return c, None
        self._tally("gate_flops", c.numel())
        c = c.contiguous()
        spec = _policy_spec(self.policy, c.shape[-2])
        if spec is None:
            index = self.policy(native.sub(c, self.p), dim=-1)
        else:
            index = native.gate_select(c, p=self.p, **spec)[0]
        self.p = c
        return native.gate_gather(c, _broadcast_index(index, c))[0], index


class TokenBuffer(_FrameState):
    """Persistent tensor b into which updated rows / columns are scattered (modules.py:52-101); returns b itself."""

    _state = ("b",)

    def __init__(self, structure="row"):
        assert structure in ("row", "col")
        super().__init__()
        self.structure = structure

    def forward(self, x, index):
        return self.forward_first(x) if self.first else self.forward_incremental(x, index)

    def forward_first(self, x):
        self.first, self.b = False, x.clone()
        return self.b

    def forward_incremental(self, x, index):
        native.buffer_scatter(self.b, x.contiguous(), index.contiguous(), structure=self.structure)
        return self.b


class _ProductState(_FrameState):
    _state = ("_product",)

    def __init__(self):
        super().__init__()
        self.matmul = CountedMatmul()
        self._producer = None

    # `product` is the stored state tensor; the fused Eventful blocks do not store the query-key product and install
    # a producer that recomputes it from their QKV buffer on demand instead.
    @property
    def product(self):
        if self._product is None and self._producer is not None and not self.first:
            return self._producer()
        return self._product

    @product.setter
    def product(self, value):
        self._product = value

    def _first_product(self, a, b):
        self.first, self.product = False, self.matmul(a, b)
        return self.product


class MatmulBuffer(_ProductState):
    """
    Query-key product buffer (modules.py:204-252): rows index_q, then columns index_k are refreshed; the
    result is the state tensor itself.  The Eventful blocks of this package do not keep this N x N state:
    the product always equals (q / scale) k^T of the current QKV buffer, so the attention kernels recompute
    it on tensor cores.
    """

    def forward(self, q, k, index_q, index_k):
        if self.first:
            return self._first_product(q, k)
        rows = native.gate_gather(q.contiguous(), _broadcast_index(index_q, q))[0]
        cols = native.gate_gather_cols(k.contiguous(), index_k.contiguous())[0]
        native.buffer_scatter(self.product, self.matmul(rows, k), index_q.contiguous(), structure="row")
        native.buffer_scatter(self.product, self.matmul(q, cols), index_k.contiguous(), structure="col")
        return self.product


class MatmulDeltaAccumulator(_ProductState):
    """Attention-value product updated by the two delta terms of modules.py:285-295; returns the state tensor."""

    def forward(self, a_n_tilde, v_n_tilde, a_delta_tilde, v_delta_tilde):
        if self.first:
            return self._first_product(a_n_tilde, v_n_tilde)
        if self.count_mode:
            self.counts["accumulator_flops"] += v_n_tilde.numel() + 2 * self.product.numel()
            self.matmul.counts["matmul_flops"] += 2 * self.product.numel() * a_n_tilde.shape[-1]
        residual = native.sub(v_n_tilde.contiguous(), v_delta_tilde.contiguous())
        native.bmm(a_n_tilde, v_delta_tilde, out=self.product, accumulate=True)
        native.bmm(a_delta_tilde, residual, out=self.product, accumulate=True)
        return self.product
