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


class _GateBase(ExtendedModule):
    def __init__(self, structure="row"):
        super().__init__()
        assert structure in ["row", "col"]
        self.structure = structure
        self.first = True
        self.policy = None
        self.p = None

    def reset_self(self):
        self.first = True
        self.p = None

    def _select(self, c, p):
        """Index chosen by self.policy for gate input c against state p."""
        if self.structure == "row":
            spec = _policy_spec(self.policy, c.shape[-2])
            if spec is not None:
                if "threshold" in spec:
                    assert all(size == 1 for size in c.shape[:-2])  # policies.py:25
                index, _ = native.gate_select(c, p=p, **spec)
                return index
            return self.policy(native.sub(c, p), dim=-1)
        return self.policy(native.sub(c, p), dim=-2)

    def _update(self, c, index, want_delta):
        """c~ (and e~) at index, and p[index] = c~."""
        if self.structure == "row":
            return native.gate_gather(c, _broadcast_index(index, c), p=self.p, want_delta=want_delta)
        return native.gate_gather_cols(c, index.contiguous(), p=self.p, want_delta=want_delta)


class TokenGate(_GateBase):
    """Token gate: selects tokens whose value drifted from the reference state p (modules.py:104-168)."""

    def forward(self, c, forced_index=None):
        """Warning - self.p keeps a direct reference to the first input and is updated in place."""
        if self.first:
            return self.forward_first(c)
        return self.forward_incremental(c, forced_index=forced_index)

    def forward_first(self, c):
        self.first = False
        self.p = c
        return c, None

    def forward_incremental(self, c, forced_index=None):
        if self.count_mode:
            self.counts["gate_flops"] += self.p.numel()
        c = c.contiguous()
        index = self._select(c, self.p) if forced_index is None else forced_index
        c_tilde, _ = self._update(c, index, want_delta=False)
        return c_tilde, index


class TokenDeltaGate(_GateBase):
    """Token gate that also returns the gathered error e~ (modules.py:171-201)."""

    def forward(self, c, forced_index=None):
        if self.first:
            return self.forward_first(c)
        return self.forward_incremental(c, forced_index=forced_index)

    def forward_first(self, c):
        self.first = False
        self.p = c
        return c, None, None

    def forward_incremental(self, c, forced_index=None):
        if self.count_mode:
            self.counts["gate_flops"] += self.p.numel()
        c = c.contiguous()
        index = self._select(c, self.p) if forced_index is None else forced_index
        c_tilde, e_tilde = self._update(c, index, want_delta=True)
        return c_tilde, e_tilde, index


class SimpleSTGTGate(ExtendedModule):
    """Baseline gate of "Spatio-Temporal Gated Transformers": p is replaced wholesale (modules.py:6-49)."""

    def __init__(self, structure="row"):
        super().__init__()
        assert structure == "row"
        self.first = True
        self.policy = None
        self.p = None

    def forward(self, c):
        if self.first:
            self.first = False
            self.p = c
            return c, None
        if self.count_mode:
            self.counts["gate_flops"] += c.numel()
        c = c.contiguous()
        spec = _policy_spec(self.policy, c.shape[-2])
        if spec is not None:
            index, _ = native.gate_select(c, p=self.p, **spec)
        else:
            index = self.policy(native.sub(c, self.p), dim=-1)
        c_tilde, _ = native.gate_gather(c, _broadcast_index(index, c))
        self.p = c
        return c_tilde, index

    def reset_self(self):
        self.first = True
        self.p = None


class TokenBuffer(ExtendedModule):
    """Token buffer: scatters updated tokens into a persistent tensor (modules.py:52-101)."""

    def __init__(self, structure="row"):
        super().__init__()
        assert structure in ["row", "col"]
        self.structure = structure
        self.first = True
        self.b = None

    def forward(self, x, index):
        """Warning - the output is a direct reference to self.b."""
        if self.first:
            return self.forward_first(x)
        return self.forward_incremental(x, index)

    def forward_first(self, x):
        self.first = False
        self.b = x.clone()
        return self.b

    def forward_incremental(self, x, index):
        native.buffer_scatter(self.b, x.contiguous(), index.contiguous(), structure=self.structure)
        return self.b

    def reset_self(self):
        self.first = True
        self.b = None


class MatmulBuffer(ExtendedModule):
    """
    Query-key product buffer (modules.py:204-252): rows index_q then columns index_k are refreshed.
    The Eventful blocks of this package do not keep this N x N state -- the product always equals
    (q / scale) k^T of the current QKV buffer, so the attention kernels recompute it on tensor cores.
    """

    def __init__(self):
        super().__init__()
        self.first = True
        self.product = None
        self.matmul = CountedMatmul()

    def forward(self, q, k, index_q, index_k):
        """Warning - the output is a direct reference to self.product."""
        if self.first:
            self.first = False
            self.product = self.matmul(q, k)
            return self.product
        q_tilde, _ = native.gate_gather(q.contiguous(), _broadcast_index(index_q, q))
        k_tilde, _ = native.gate_gather_cols(k.contiguous(), index_k.contiguous())
        native.buffer_scatter(self.product, self.matmul(q_tilde, k), index_q.contiguous(), structure="row")
        native.buffer_scatter(self.product, self.matmul(q, k_tilde), index_k.contiguous(), structure="col")
        return self.product

    def reset_self(self):
        self.first = True
        self.product = None


class MatmulDeltaAccumulator(ExtendedModule):
    """Attention-value product accumulator (modules.py:255-299)."""

    def __init__(self):
        super().__init__()
        self.first = True
        self.product = None
        self.matmul = CountedMatmul()

    def forward(self, a_n_tilde, v_n_tilde, a_delta_tilde, v_delta_tilde):
        """Warning - the output is a direct reference to self.product."""
        if self.first:
            self.first = False
            self.product = self.matmul(a_n_tilde, v_n_tilde)
            return self.product
        if self.count_mode:
            self.counts["accumulator_flops"] += v_n_tilde.numel() + 2 * self.product.numel()
            self.matmul.counts["matmul_flops"] += 2 * self.product.numel() * a_n_tilde.shape[-1]
        native.bmm(a_n_tilde, v_delta_tilde, out=self.product, accumulate=True)
        native.bmm(a_delta_tilde, native.sub(v_n_tilde.contiguous(), v_delta_tilde.contiguous()), out=self.product,
                   accumulate=True)
        return self.product

    def reset_self(self):
        self.first = True
        self.product = None
