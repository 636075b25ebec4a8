"""
Token selection policies (API mirror of the reference's policies.py).

Called stand-alone, a policy receives the error tensor e = c - p and returns int64 indices; the
norm + selection run as one fused kernel (et_gate_select with p = NULL).  Inside TokenGate and the
Eventful blocks the same kernel also performs the subtraction (and the preceding LayerNorm /
residual add), so the error tensor is never materialised; the gate reads `policy.fused_spec()`.
"""

from eventful_transformer import _native as native
from eventful_transformer.base import ExtendedModule


class _NormPolicy(ExtendedModule):
    order = 2

    def _rows(self, x, dim):
        """Returns x arranged so that the norm is reduced over the last dim."""
        if self.order != 2:
            raise NotImplementedError("eventful_b200 implements the L2 token norm (order=2) only")
        if dim in (-1, x.ndim - 1):
            return x.contiguous()
        if dim in (-2, x.ndim - 2):  # column-structure gates: reduce over rows (modules.py:157)
            return x.transpose(-1, -2).contiguous()
        raise ValueError(f"unsupported reduction dim {dim}")


class TokenNormThreshold(_NormPolicy):
    """Selects tokens whose error norm exceeds a threshold (policies.py:6-36). Batch size must be 1."""

    def __init__(self, threshold=0.0, order=2):
        super().__init__()
        self.threshold = threshold
        self.order = order

    def fused_spec(self, n_tokens):
        return None if self.order != 2 else dict(threshold=float(self.threshold))

    def forward(self, x, dim=-1):
        assert all(size == 1 for size in x.shape[:-2])  # policies.py:25
        index, _ = native.gate_select(self._rows(x, dim), threshold=float(self.threshold))
        return index.view((1,) * (x.ndim - 2) + (-1,))


class TokenNormTopK(_NormPolicy):
    """Selects the k tokens with the largest error norm (policies.py:39-68)."""

    def __init__(self, k, order=2, save_status=False):
        super().__init__()
        self.k = k
        self.order = order
        self.save_status = save_status
        self.last_input = None
        self.last_output = None

    def fused_spec(self, n_tokens):
        if self.order != 2 or self.save_status:
            return None  # save_status needs the materialised error tensor
        return dict(k=int(self.k))

    def forward(self, x, dim=-1):
        output, _ = native.gate_select(self._rows(x, dim), k=int(self.k))
        if self.save_status:
            self.last_input = x.clone()
            self.last_output = output.clone()
        return output


class TokenNormTopFraction(_NormPolicy):
    """Selects the top int(fraction * N) tokens by error norm (policies.py:71-95)."""

    def __init__(self, fraction, order=2):
        super().__init__()
        assert not (fraction < 0.0 or fraction > 1.0)
        self.fraction = fraction
        self.order = order

    def fused_spec(self, n_tokens):
        return None if self.order != 2 else dict(k=int(self.fraction * n_tokens))

    def forward(self, x, dim=-1):
        rows = self._rows(x, dim)
        output, _ = native.gate_select(rows, k=int(self.fraction * rows.shape[-2]))
        return output
