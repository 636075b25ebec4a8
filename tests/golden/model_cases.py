"""Tiny model-level cases shared by make_golden_models.py (reference side) and tests/test_models_gpu.py (this package)."""
import torch

VIVIT_TINY = dict(
    seed=31, k=3,
    video_shape=(1, 12, 3, 32, 48),  # 12 frames of 32 x 48: two temporal views of 8 frames, two 32 x 32 crops, no resizing
    model=dict(
        classes=10, input_shape=(8, 3, 32, 32), normalize_mean=0.45, normalize_std=0.225,
        spatial_config=dict(depth=2, position_encoding_size=[2, 2], block_class="EventfulBlock",
                            block_config=dict(dim=32, heads=2, mlp_ratio=2)),
        spatial_views=2,
        temporal_config=dict(depth=1, position_encoding_size=[4], block_class="Block",
                             block_config=dict(dim=32, heads=2, mlp_ratio=2)),
        temporal_stride=1, temporal_views=2, tubelet_shape=(2, 16, 16)),
)

# The EPIC-Kitchens "temporal + ATS" shape (configs/evaluate/vivit_epic_kitchens/temporal_ats_200.yml): adaptive token sampling
# in the spatial blocks.  2 spatial x 2 temporal views are batched, and the spatial blocks have 4 heads: the reference's ATS
# code needs batch == heads (ViViT-B: 12 views, 12 heads).  48 x 48 crops -> 3 x 3 patches + class token = 10 tokens,
# sampled down to 7, then 5.
VIVIT_TINY_ATS = dict(
    seed=41, k=4,
    video_shape=(1, 12, 3, 48, 72),
    model=dict(
        classes=10, input_shape=(8, 3, 48, 48), normalize_mean=0.45, normalize_std=0.225,
        spatial_config=dict(depth=2, position_encoding_size=[3, 3], block_class="EventfulBlock",
                            block_config=dict(dim=32, heads=4, mlp_ratio=2, ats_fraction=0.7)),
        spatial_views=2,
        temporal_config=dict(depth=1, position_encoding_size=[4], block_class="Block",
                             block_config=dict(dim=32, heads=2, mlp_ratio=2)),
        temporal_stride=1, temporal_views=2, tubelet_shape=(2, 16, 16)),
)

VITDET_STEM_TINY = dict(
    seed=37, k=6, frames=3, image_shape=(3, 50, 60),  # padded to 64 x 64 by the preprocessing
    input_shape=(3, 64, 64), normalize_mean=[123.675, 116.28, 103.53], normalize_std=[58.395, 57.12, 57.375], patch_size=16,
    backbone_config=dict(depth=2, position_encoding_size=[2, 2], block_class="EventfulBlock",
                         windowed_class="EventfulTokenwiseBlock", window_indices=[0],
                         block_config=dict(dim=32, heads=2, mlp_ratio=2, window_size=[2, 2], relative_embedding_size=[3, 3])),
)


def seeded_state(template, seed):
    """A state dict with the template's keys / shapes: matrices ~ N(0, 0.2), LayerNorm weights ~ 1, everything else small."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key, value in template.items():
        if key.endswith("layer_norm.weight"):
            out[key] = 1.0 + 0.1 * torch.randn(value.shape, generator=g)
        elif value.ndim >= 2:
            out[key] = 0.2 * torch.randn(value.shape, generator=g)
        else:
            out[key] = 0.1 * torch.randn(value.shape, generator=g)
    return out


def vivit_video(cfg):
    """uint8 video with temporal structure: a random base image plus slowly growing noise."""
    g = torch.Generator().manual_seed(cfg["seed"] + 1)
    b, t, c, h, w = cfg["video_shape"]
    base = torch.rand((b, 1, c, h, w), generator=g)
    drift = torch.rand((b, t, c, h, w), generator=g) * torch.linspace(0.0, 0.4, t).view(1, t, 1, 1, 1)
    return ((base * 0.6 + drift).clamp(0, 1) * 255).to(torch.uint8)


def vitdet_frames(cfg):
    g = torch.Generator().manual_seed(cfg["seed"] + 1)
    c, h, w = cfg["image_shape"]
    base = torch.rand((1, c, h, w), generator=g)
    drift = torch.rand((cfg["frames"], c, h, w), generator=g) * torch.linspace(0.0, 0.3, cfg["frames"]).view(-1, 1, 1, 1)
    return ((base * 0.7 + drift).clamp(0, 1) * 255).to(torch.uint8)
