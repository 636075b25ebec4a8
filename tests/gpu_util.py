"""Shared helpers for the -m gpu parity tests (CUDA path vs. oracle/)."""
import torch

import eventful_oracle as orc
from cases import GATES
from eventful_transformer import backbones, modules, policies
import et_synthetic as syn

DEV = "cuda"


def build_gpu_backbone(case, params, dtype=torch.bfloat16):
    kw = syn.backbone_kwargs(case["cfg"], case["input_size"], block_class=case["block_class"],
                             windowed_class=case.get("windowed_class", "EventfulTokenwiseBlock"),
                             matmul_2_cast=None, has_class_token=case.get("has_class_token", False))
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
