"""Shared helpers for the -m gpu parity tests (CUDA path vs. oracle/)."""
import torch

import eventful_oracle as orc
from cases import GATES
from eventful_transformer import backbones, modules, policies
import et_synthetic as syn

DEV = "cuda"


def build_gpu_backbone(case, params, dtype=torch.bfloat16, cast=None):
    """`cast`: matmul_2_cast of the global blocks (the windowed ones get null, as in the reference's configs)."""
    kw = syn.backbone_kwargs(case["cfg"], case["input_size"], block_class=case["block_class"],
                             windowed_class=case.get("windowed_class", "EventfulTokenwiseBlock"),
                             matmul_2_cast=cast, has_class_token=case.get("has_class_token", False),
                             pool_size=case.get("pool_size"), ats_fraction=case.get("ats_fraction"))
    if kw.get("windowed_class") is None:
        kw.pop("windowed_class", None)
    for flag in ("gate_before_ln", "stgt"):
        if case.get(flag):
            kw["block_config"][flag] = True
    model = backbones.ViTBackbone(**kw)
    model.load_state_dict(params, strict=True)
    model = model.to(DEV).to(dtype).eval()
    if case["policy"] is not None:
        kind, pk = case["policy"]
        cls = dict(topk=policies.TokenNormTopK, threshold=policies.TokenNormThreshold,
                   fraction=policies.TokenNormTopFraction)[kind]
        for gate_cls in (modules.SimpleSTGTGate, modules.TokenDeltaGate, modules.TokenGate):
            for gate in model.modules_of_type(gate_cls):
                gate.policy = cls(**pk)
    return model


def gpu_trace(model):
    """(block, gate) -> last selected index (CPU int64) of the policy-driven gates."""
    out = {}
    for i, block in enumerate(model.blocks):
        for gate in GATES:
            g = getattr(block, gate, None)
            if g is not None and getattr(g, "last_index", None) is not None:
                out[(i, gate)] = g.last_index.detach().cpu()
    return out


def ats_trace(model):
    """(block, "ats") -> the stabilised adaptive-token-sampling index of the last frame (CPU int64)."""
    return {(i, "ats"): block.last_ats_indices.detach().cpu() for i, block in enumerate(model.blocks)
            if getattr(block, "last_ats_indices", None) is not None}


def rounded(params, dtype):
    """Parameters rounded to `dtype` but held in fp32 (the 'exact arithmetic on the same weights' reference)."""
    return {k: v.to(dtype).float() for k, v in params.items()}


def rel_err(got, want):
    return float((got.float() - want.float()).abs().max() / want.float().abs().max().clamp_min(1e-6))


def one_block_oracle(params, dim, heads, input_size, cls, window=None, rel=None, has_class_token=False):
    return orc.OracleBackbone(params, depth=1, dim=dim, heads=heads, input_size=input_size,
                              position_encoding_size=input_size, block_class=cls, windowed_class=cls,
                              window_indices=(0,) if window else (), window_size=window,
                              relative_embedding_size=rel, has_class_token=has_class_token)


# ------------------------------------------------------------------------------------------
# measured-parity log: every GPU parity test records what it measured (not just pass / fail);
# gpurun brings gpurun_out/ back, and the numbers are quoted in DESIGN.md / profiles/.
# ------------------------------------------------------------------------------------------
import json
import os

_PARITY_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_measured.jsonl")


def record(test, **values):
    try:
        os.makedirs(os.path.dirname(_PARITY_LOG), exist_ok=True)
        with open(_PARITY_LOG, "a") as f:
            f.write(json.dumps(dict(test=test, **values)) + "\n")
    except OSError:
        pass


def elem_err(got, want):
    """Per-element error statistics: (max |got - want|, max |got - want| / (atol_unit + |want|), rms(want))."""
    got, want = got.float(), want.float()
    diff = (got - want).abs()
    rms = float(want.pow(2).mean().sqrt())
    return float(diff.max()), float((diff / (rms + want.abs())).max()), rms


def allclose_report(got, want, rtol, atol_rms, what):
    """|got - want| <= atol_rms * rms(want) + rtol * |want| per element; returns the measured worst ratio."""
    got, want = got.float(), want.float()
    rms = float(want.pow(2).mean().sqrt())
    bound = atol_rms * rms + rtol * want.abs()
    ratio = float(((got - want).abs() / bound).max())
    assert ratio <= 1.0, f"{what}: worst |err| / (atol + rtol |want|) = {ratio:.3f} (rtol={rtol}, atol={atol_rms} x rms={rms:.4f})"
    return ratio


def selection_agreement(gpu_index, free_index, norm, threshold=None):
    """
    Compares a CUDA selection with the oracle's own selection on (near-)identical inputs.
    Returns (overlap fraction |A & B| / max(|A|, |B|), worst relative distance of a token in the symmetric difference
    from the decision boundary: the k-th largest norm for top-k policies, the threshold for TokenNormThreshold).
    """
    lead = norm.shape[:-1]
    rows = 1
    for s_ in lead:
        rows *= s_
    gi = gpu_index.reshape(rows, -1)
    fi = free_index.reshape(rows, -1)
    nr = norm.reshape(rows, -1).float()
    overlap, worst = 1.0, 0.0
    for r in range(rows):
        a, b = set(gi[r].tolist()), set(fi[r].tolist())
        assert len(a) == gi.shape[-1], "CUDA selection holds duplicate indices"
        if threshold is None:
            assert len(a) == len(b)
        if not a and not b:
            continue
        edge = float(threshold) if threshold is not None else float(nr[r].topk(len(b))[0][-1])
        overlap = min(overlap, len(a & b) / max(len(a), len(b)))
        for t in a ^ b:
            worst = max(worst, abs(float(nr[r, t]) - edge) / max(edge, 1e-30))
    return overlap, worst
