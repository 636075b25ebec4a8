"""
Seeded synthetic weights and token streams for the gated-update path.

The reference initialises every parameter to zero (counting.py:143-144,
eventful_transformer/utils.py:49,125-130), so every test / benchmark has to
initialise explicitly.  This module is shared by tests, bench.py and the golden
fixture generator so that all of them see bit-identical inputs (CPU generator,
float32 draws, cast afterwards).
"""

from math import prod

import torch

VITDET_B = dict(  # configs/models/vitdet_b_coco.yml
    depth=12,
    dim=768,
    heads=12,
    mlp_ratio=4,
    position_encoding_size=(14, 14),
    window_indices=(0, 1, 3, 4, 6, 7, 9, 10),
    window_size=(14, 14),
    relative_embedding_size=(64, 64),
)


def backbone_kwargs(cfg, input_size, block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                    matmul_2_cast=None, has_class_token=False, pool_size=None, ats_fraction=None):
    """kwargs for ViTBackbone(...) in the shape configs/models/*.yml feeds it (backbones.py:13-24)."""
    block_config = dict(dim=cfg["dim"], heads=cfg["heads"], mlp_ratio=cfg["mlp_ratio"])
    if cfg.get("relative_embedding_size") is not None:
        block_config["relative_embedding_size"] = list(cfg["relative_embedding_size"])
    if cfg.get("window_size") is not None:
        block_config["window_size"] = list(cfg["window_size"])
    if matmul_2_cast is not None:
        block_config["matmul_2_cast"] = matmul_2_cast
    if pool_size is not None:
        block_config["pool_size"] = list(pool_size)
    if ats_fraction is not None:
        block_config["ats_fraction"] = ats_fraction
    kw = dict(
        block_config=block_config,
        depth=cfg["depth"],
        position_encoding_size=list(cfg["position_encoding_size"]),
        input_size=tuple(input_size),
        block_class=block_class,
        has_class_token=has_class_token,
        window_indices=tuple(cfg.get("window_indices", ())),
    )
    if cfg.get("window_indices"):
        kw["windowed_class"] = windowed_class
        # the reference's configs switch both variants off on windowed blocks (configs/*/vitdet_vid/_spatial.yml, _half.yml)
        overrides = {}
        if matmul_2_cast is not None:
            overrides["matmul_2_cast"] = None
        if pool_size is not None:
            overrides["pool_size"] = None
        if overrides:
            kw["windowed_overrides"] = overrides
    return kw


def param_shapes(cfg, has_class_token=False):
    """Ordered {state-dict key: shape} for a ViTBackbone (SURVEY.md 8(b) key table)."""
    d, h = cfg["dim"], cfg["heads"]
    m = cfg["mlp_ratio"] * d
    shapes = {
        "position_encoding.encoding": (1, prod(cfg["position_encoding_size"]) + int(has_class_token), d)
    }
    for i in range(cfg["depth"]):
        pre = f"blocks.{i}."
        shapes[pre + "input_layer_norm.weight"] = (d,)
        shapes[pre + "input_layer_norm.bias"] = (d,)
        shapes[pre + "qkv.weight"] = (3 * d, d)
        shapes[pre + "qkv.bias"] = (3 * d,)
        rel = cfg.get("relative_embedding_size")
        if rel is not None:
            if i in cfg.get("window_indices", ()) and cfg.get("window_size") is not None:
                rel = cfg["window_size"]
            shapes[pre + "relative_position.y_embedding"] = (2 * rel[0] - 1, d // h)
            shapes[pre + "relative_position.x_embedding"] = (2 * rel[1] - 1, d // h)
        shapes[pre + "projection.weight"] = (d, d)
        shapes[pre + "projection.bias"] = (d,)
        shapes[pre + "mlp_layer_norm.weight"] = (d,)
        shapes[pre + "mlp_layer_norm.bias"] = (d,)
        shapes[pre + "mlp_1.weight"] = (m, d)
        shapes[pre + "mlp_1.bias"] = (m,)
        shapes[pre + "mlp_2.weight"] = (d, m)
        shapes[pre + "mlp_2.bias"] = (d,)
    return shapes


def seeded_params(cfg, seed=0, std=0.02, dtype=torch.float32, has_class_token=False):
    """
    Deterministic random parameters: matrices / tables ~ N(0, std), LayerNorm
    weight ~ 1 + N(0, 0.1), all biases ~ N(0, std).  Drawn in float32 on the CPU
    generator in key order, then cast.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key, shape in param_shapes(cfg, has_class_token).items():
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if key.endswith("layer_norm.weight"):
            t = 1.0 + 0.1 * t
        else:
            t = std * t
        out[key] = t.to(dtype)
    return out


def token_stream(batch, tokens, dim, frames, seed=0, mode="drift", dtype=torch.float32, step=0.1):
    """
    Synthetic backbone-level video: list of `frames` tensors (batch, tokens, dim).
      drift : x_t = x_0 + step * t * eps_t                      
      patch : static background, a moving block of tokens changes   (many exact-zero deltas -> ties)
      dyadic: deltas are multiples of 2^-4 so sums of squares are exact in fp32
    """
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn((batch, tokens, dim), generator=g, dtype=torch.float32)
    if mode == "dyadic":
        x0 = torch.round(x0 * 4.0) / 4.0
    out = [x0.to(dtype)]
    for t in range(1, frames):
        if mode == "drift":
            xt = x0 + step * t * torch.randn(x0.shape, generator=g, dtype=torch.float32)
        elif mode == "patch":
            xt = x0.clone()
            width = max(1, tokens // 3)
            start = (t * max(1, tokens // 7)) % tokens
            sel = (torch.arange(width) + start) % tokens
            xt[:, sel] = xt[:, sel] + torch.randn((batch, width, dim), generator=g, dtype=torch.float32)
        elif mode == "dyadic":
            ints = torch.randint(-3, 4, x0.shape, generator=g).float()
            keep = (torch.rand((batch, tokens, 1), generator=g) < 0.6).float()
            xt = x0 + ints * keep / 16.0
        else:
            raise ValueError(mode)
        out.append(xt.to(dtype))
    return out
