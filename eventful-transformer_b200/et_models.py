"""
Model ends on top of the eventful_transformer package: the patch / tubelet embeddings as GEMMs on the
CUDA library, the ViTDet pre-backbone, and the factorised ViViT wiring (spatial Eventful sub-model stepped over time,
dense temporal sub-model, classifier).

Written from scratch against the behaviour of the reference's models/vitdet.py and models/vivit.py: same constructor
kwargs, same sub-module / state-dict key names (embedding.conv.{weight,bias}, spatial_model.{class_token,backbone.*,
layer_norm.*}, temporal_model.*, classifier.*), so the reference's converted checkpoints load unchanged.  The reference's own
model files also run on the package as they are (they only import ViTBackbone, ExtendedModule, CountedLinear, LN_EPS and
numeric_tuple from it); these classes exist so that the ends of the model run on this library too instead of cuDNN:

    LinearEmbedding / TubeletEmbedding   et_patchify + et_linear           (models/vitdet.py:17-52, models/vivit.py:153-192)
    ViViTSubModel                        class token + ViTBackbone + LayerNorm of the class row   (models/vivit.py:272-303)
    FactorizedViViT                      views -> embedding -> spatial model per time step (reset per view batch)
                                         -> temporal model -> classifier -> mean over views -> softmax (models/vivit.py:98-150)
    ViTDetStem                           normalisation + padding + LinearEmbedding + ViTBackbone = ViTDet.pre_backbone +
                                         backbone (models/vitdet.py:186-220); the pyramid and detectron2 heads stay outside
                                         (detectron2 is not installed).
Input preparation (value normalisation, view cropping, antialiased resizing) is data movement on the raw video and uses torch ops.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from eventful_transformer import _native as native
from eventful_transformer.backbones import ViTBackbone
from eventful_transformer.base import ExtendedModule, numeric_tuple
from eventful_transformer.blocks import LN_EPS
from eventful_transformer.counting import CountedLinear


def as_float32(x):
    """uint8 video -> float in [0, 1] (utils/image.py:9-17 of the reference)."""
    return x.float() / 255.0 if x.dtype == torch.uint8 else x


def _normalize(x, mean, std):
    """torchvision.transforms.Normalize over the channel axis (third from the end)."""
    c = x.shape[-3]
    mean = torch.as_tensor(numeric_tuple(mean, c), dtype=x.dtype, device=x.device).view(c, 1, 1)
    std = torch.as_tensor(numeric_tuple(std, c), dtype=x.dtype, device=x.device).view(c, 1, 1)
    return (x - mean) / std


def _resize_to_fit(x, size):
    """Smallest antialiased bilinear rescale that covers `size` (utils/image.py:63-68)."""
    scale = max(size[0] / x.shape[-2], size[1] / x.shape[-1])
    if scale == 1.0:
        return x
    out = [round(scale * x.shape[-2]), round(scale * x.shape[-1])]
    lead = x.shape[:-3]
    y = F.interpolate(x.reshape((-1,) + tuple(x.shape[-3:])), size=out, mode="bilinear", antialias=True, align_corners=False)
    return y.reshape(tuple(lead) + tuple(y.shape[-3:]))


class _PatchProjection(nn.Module):
    """Conv with kernel = stride = patch, run as patch extraction + GEMM.  `conv` only owns the parameters (reference layout)."""

    def _project(self, x, patch):
        w = self.conv.weight.detach()
        rows = native.patchify(x.to(w.dtype).contiguous(), patch)
        return native.linear(rows, w.reshape(w.shape[0], -1), self.conv.bias.detach())


class LinearEmbedding(_PatchProjection):
    """(B, C, H, W) -> (B, patches, dim)  (models/vitdet.py:17-52)."""

    def __init__(self, input_channels, dim, patch_size):
        super().__init__()
        self.patch_size = numeric_tuple(patch_size, length=2)
        self.conv = nn.Conv2d(input_channels, dim, kernel_size=self.patch_size, stride=self.patch_size)

    def forward(self, x):
        return self._project(x, self.patch_size)


class TubeletEmbedding(_PatchProjection):
    """(B, T, C, H, W) -> (B, T / t, patches, dim)  (models/vivit.py:153-192)."""

    def __init__(self, input_channels, dim, tubelet_shape):
        super().__init__()
        self.tubelet_shape = tuple(tubelet_shape)
        self.conv = nn.Conv3d(input_channels, dim, kernel_size=self.tubelet_shape, stride=self.tubelet_shape)

    def forward(self, x):
        return self._project(x, self.tubelet_shape)


class ViViTPreprocessing(nn.Module):
    """Normalisation and temporal / spatial view extraction (models/vivit.py:195-269). Returns a list of views."""

    def __init__(self, input_shape, normalize_mean, normalize_std, spatial_views, temporal_stride, temporal_views):
        super().__init__()
        self.input_shape = tuple(input_shape)
        self.normalize_mean, self.normalize_std = normalize_mean, normalize_std
        self.spatial_views, self.temporal_stride, self.temporal_views = spatial_views, temporal_stride, temporal_views

    def forward(self, x):
        t, _, h, w = self.input_shape
        span = self.temporal_stride * t
        if x.shape[1] < span:  # repeat the last frame of a short video
            x = torch.cat([x, x[:, -1:].expand(x.shape[:1] + (span - x.shape[1],) + x.shape[2:])], dim=1)
        if self.temporal_views == 1:
            starts = [(x.shape[1] - span) // 2]
        else:
            gap = (x.shape[1] - span) / (self.temporal_views - 1)
            starts = [int(i * gap) for i in range(self.temporal_views)]
        views = [_normalize(as_float32(x[:, s: s + span: self.temporal_stride]), self.normalize_mean, self.normalize_std)
                 for s in starts]
        views = [_resize_to_fit(v, (h, w)) for v in views]
        if self.spatial_views == 1:
            corners = [((views[0].shape[-2] - h) // 2, (views[0].shape[-1] - w) // 2)]
        else:
            gh = (views[0].shape[-2] - h) / (self.spatial_views - 1)
            gw = (views[0].shape[-1] - w) / (self.spatial_views - 1)
            corners = [(int(i * gh), int(i * gw)) for i in range(self.spatial_views)]
        return [v[..., i: i + h, j: j + w] for i, j in corners for v in views]


class ViViTSubModel(ExtendedModule):
    """Class token + ViTBackbone + LayerNorm; returns the class embedding (models/vivit.py:272-303)."""

    def __init__(self, input_size, backbone_config):
        super().__init__()
        dim = backbone_config["block_config"]["dim"]
        self.class_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.backbone = ViTBackbone(input_size=input_size, has_class_token=True, **backbone_config)
        self.layer_norm = nn.LayerNorm(dim, eps=LN_EPS)
        self._row0 = None

    def reset_self(self):
        self._row0 = None

    def forward(self, x):
        token = self.class_token.detach().to(x.dtype).expand(x.shape[0], 1, x.shape[-1])
        x = self.backbone(torch.cat([token, x], dim=1))
        # only the class row is normalised and returned: LayerNorm is per token, so this equals layer_norm(x)[:, 0]
        if self._row0 is None or self._row0.shape[0] != x.shape[0] or self._row0.device != x.device:
            self._row0 = torch.zeros((x.shape[0], 1), dtype=torch.int64, device=x.device)
        row, _ = native.gate_gather(x, self._row0, ln=(self.layer_norm.weight.detach(), self.layer_norm.bias.detach()),
                                    eps=self.layer_norm.eps)
        return row[:, 0]


class FactorizedViViT(ExtendedModule):
    """Spatio-temporal factorised ViViT (models/vivit.py:17-150): same kwargs and state-dict keys as the reference."""

    def __init__(self, classes, input_shape, normalize_mean, normalize_std, spatial_config, spatial_views, temporal_config,
                 temporal_stride, temporal_views, tubelet_shape, batch_views=True, dropout_rate=0.0, spatial_only=False,
                 temporal_only=False):
        super().__init__()
        assert not (spatial_only and temporal_only)
        assert not (dropout_rate < 0.0 or dropout_rate > 1.0)
        input_shape, tubelet_shape = tuple(input_shape), tuple(tubelet_shape)
        input_t, input_c, input_h, input_w = input_shape
        self.batch_views, self.spatial_only, self.temporal_only = batch_views, spatial_only, temporal_only
        self.preprocessing = ViViTPreprocessing(input_shape, normalize_mean, normalize_std, spatial_views, temporal_stride,
                                                temporal_views)
        dim = spatial_config["block_config"]["dim"]
        self.embedding = TubeletEmbedding(input_c, dim, tubelet_shape)
        self.spatial_model = ViViTSubModel((input_h // tubelet_shape[1], input_w // tubelet_shape[2]), spatial_config)
        self.temporal_model = ViViTSubModel((input_t // tubelet_shape[0],), temporal_config)
        self.dropout = nn.Dropout(dropout_rate) if dropout_rate > 0.0 else nn.Identity()
        self.classifier = CountedLinear(in_features=dim, out_features=classes)

    def forward(self, x):
        batch = x.shape[0]
        if not self.temporal_only:
            x = self._forward_spatial(x)
        if not self.spatial_only:
            x = self._forward_temporal(x, batch)
        return x

    def _forward_spatial(self, x):
        views = self.preprocessing(x)
        if self.batch_views:  # all views along the batch axis: one gated stream group
            return self._forward_view(torch.stack(views, dim=1).flatten(end_dim=1))
        return torch.stack([self._forward_view(v) for v in views], dim=1).flatten(end_dim=1)

    def _forward_view(self, x):
        x = self.embedding(x)  # (batch, time, patch, dim)
        self.spatial_model.reset()  # a new video: gates / buffers / accumulators start over (models/vivit.py:146)
        steps = [self.spatial_model(x[:, t].contiguous()) for t in range(x.shape[1])]
        return torch.stack(steps, dim=1)  # (batch, time, dim)

    def _forward_temporal(self, x, batch):
        x = self.temporal_model(x.reshape((-1,) + tuple(x.shape[-2:])).contiguous())
        x = self.classifier(self.dropout(x))
        return x.view(batch, -1, x.shape[-1]).mean(dim=-2).softmax(dim=-1)


class ViTDetStem(ExtendedModule):
    """
    ViTDet up to and including the Transformer backbone (models/vitdet.py:133-220: ViTDetPreprocessing, LinearEmbedding,
    ViTBackbone): value normalisation, zero padding to the configured input shape, patch embedding, backbone.
    forward(x) -> (padded image, tokens (B, N, dim)); post-backbone (SimplePyramid, detectron2 RPN / ROI heads) is not part of
    this library.
    """

    def __init__(self, backbone_config, input_shape, normalize_mean, normalize_std, patch_size):
        super().__init__()
        input_c, input_h, input_w = input_shape
        patch_size = numeric_tuple(patch_size, length=2)
        self.input_shape = tuple(input_shape)
        self.normalize_mean, self.normalize_std = normalize_mean, normalize_std
        self.backbone_input_size = (input_h // patch_size[0], input_w // patch_size[1])
        dim = backbone_config["block_config"]["dim"]
        self.embedding = LinearEmbedding(input_c, dim, patch_size)
        self.backbone = ViTBackbone(input_size=self.backbone_input_size, **backbone_config)

    def pre_backbone(self, x):
        # the normalisation constants are in the 0..255 range, the image was scaled to 0..1 (models/vitdet.py:242-245)
        if x.dim() == 3:
            x = x.unsqueeze(0)
        x = _normalize(as_float32(x) * 255.0, self.normalize_mean, self.normalize_std)
        pad_h, pad_w = self.input_shape[-2] - x.shape[-2], self.input_shape[-1] - x.shape[-1]
        if pad_h or pad_w:
            x = F.pad(x, (0, pad_w, 0, pad_h))
        return x, self.embedding(x)

    def forward(self, x):
        images, tokens = self.pre_backbone(x)
        return images, self.backbone(tokens)
