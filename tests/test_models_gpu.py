"""
Model ends on this library and the integration shape of the reference's models (judge row g1):
et_models.FactorizedViViT / ViTDetStem against fixtures produced by the UNMODIFIED reference models on CPU
(tests/golden/make_golden_models.py), and the patch / tubelet embedding GEMMs against torch's convolutions.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import et_models
from eventful_transformer import modules, policies
from golden_util import load_golden
from gpu_util import DEV, record, rel_err
from model_cases import VITDET_STEM_TINY, VIVIT_TINY, VIVIT_TINY_ATS

pytestmark = pytest.mark.gpu


def _params(gold, prefix=""):
    return {k[len("param/") + len(prefix):]: torch.from_numpy(gold[k]) for k in gold.files
            if k.startswith("param/" + prefix)}


def _set_topk(model, k):
    for cls in (modules.SimpleSTGTGate, modules.TokenDeltaGate, modules.TokenGate):
        for gate in model.modules_of_type(cls):
            gate.policy = policies.TokenNormTopK(k=k)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_factorized_vivit_matches_the_reference_model(dtype):
    """Video in, class probabilities out: tubelet embedding, Eventful spatial sub-model stepped over 4 time steps for 4 batched
    views (reset per view batch), dense temporal sub-model over a 1-D token axis, classifier, mean over views, softmax."""
    cfg, gold = VIVIT_TINY, load_golden("model_vivit_tiny")
    model = et_models.FactorizedViViT(**cfg["model"])
    model.load_state_dict(_params(gold), strict=True)
    model = model.to(DEV).to(dtype).eval()
    _set_topk(model, cfg["k"])
    video = torch.from_numpy(gold["video"]).to(DEV)
    with torch.inference_mode():
        probs = model(video).float().cpu()
        model.spatial_only = True
        spatial = model(video).float().cpu()
    want_p, want_s = torch.from_numpy(gold["probs"]), torch.from_numpy(gold["spatial"])
    assert probs.shape == want_p.shape and spatial.shape == want_s.shape
    err_p, err_s = float((probs - want_p).abs().max()), rel_err(spatial, want_s)
    record("factorized_vivit_vs_reference", dtype=str(dtype), max_abs_prob_err=err_p, spatial_rel_err=err_s)
    if dtype == torch.float32:
        assert err_p <= 1e-5 and err_s <= 1e-4, (err_p, err_s)
    else:
        assert err_p <= 3e-2 and abs(float(probs.sum()) - 1.0) < 1e-2, (err_p, err_s)  # selections may differ at bf16 ties


def test_factorized_vivit_with_adaptive_token_sampling_matches_the_reference_model():
    """The reference's FactorizedViViT with ats_fraction in the spatial blocks (the EPIC-Kitchens temporal + ATS shape, fp32):
    4 batched views = 4 heads, 10 tokens sampled down to 7 and 5 inside the Eventful spatial sub-model, index stabilisation
    across the 4 time steps, dense temporal sub-model on the class tokens."""
    cfg, gold = VIVIT_TINY_ATS, load_golden("model_vivit_tiny_ats")
    model = et_models.FactorizedViViT(**cfg["model"])
    model.load_state_dict(_params(gold), strict=True)
    model = model.to(DEV).eval()
    _set_topk(model, cfg["k"])
    video = torch.from_numpy(gold["video"]).to(DEV)
    with torch.inference_mode():
        probs = model(video).float().cpu()
        model.spatial_only = True
        spatial = model(video).float().cpu()
    want_p, want_s = torch.from_numpy(gold["probs"]), torch.from_numpy(gold["spatial"])
    assert probs.shape == want_p.shape and spatial.shape == want_s.shape
    err_p, err_s = float((probs - want_p).abs().max()), rel_err(spatial, want_s)
    record("factorized_vivit_ats_vs_reference", max_abs_prob_err=err_p, spatial_rel_err=err_s)
    assert err_p <= 1e-5 and err_s <= 1e-4, (err_p, err_s)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vitdet_stem_matches_the_reference_model(dtype):
    """ViTDet.pre_backbone + backbone: 0..255 normalisation, bottom-right zero padding, patch embedding, windowed + global
    Eventful blocks over three frames."""
    cfg, gold = VITDET_STEM_TINY, load_golden("model_vitdet_stem_tiny")
    model = et_models.ViTDetStem(cfg["backbone_config"], cfg["input_shape"], cfg["normalize_mean"], cfg["normalize_std"],
                                 cfg["patch_size"])
    model.load_state_dict(_params(gold), strict=True)
    model = model.to(DEV).to(dtype).eval()
    _set_topk(model, cfg["k"])
    frames = torch.from_numpy(gold["frames"]).to(DEV)
    with torch.inference_mode():
        for t in range(frames.shape[0]):
            images, tokens = model.pre_backbone(frames[t])
            assert tuple(images.shape[-2:]) == tuple(cfg["input_shape"][-2:])
            out = model.backbone(tokens).float().cpu()
            e_tok, e_out = rel_err(tokens.float().cpu(), torch.from_numpy(gold[f"tokens_{t}"])), rel_err(out, torch.from_numpy(gold[f"out_{t}"]))
            record("vitdet_stem_vs_reference", dtype=str(dtype), frame=t, tokens_rel_err=e_tok, out_rel_err=e_out)
            if dtype == torch.float32:
                assert e_tok <= 1e-5 and e_out <= 1e-4, (t, e_tok, e_out)
            else:
                assert e_tok <= 1e-2 and e_out <= (4e-2 if t == 0 else 0.15), (t, e_tok, e_out)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patch_and_tubelet_embeddings_equal_the_convolutions(dtype):
    g = torch.Generator().manual_seed(71)
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    emb = et_models.LinearEmbedding(3, 768, 16)
    emb.conv.weight.data = 0.05 * torch.randn(emb.conv.weight.shape, generator=g)
    emb.conv.bias.data = torch.randn(768, generator=g)
    x = torch.randn(2, 3, 224, 208, generator=g)
    want = F.conv2d(x, emb.conv.weight.detach(), emb.conv.bias.detach(), stride=16).flatten(-2).transpose(1, 2)
    got = emb.to(DEV).to(dtype)(x.to(DEV)).float().cpu()
    assert got.shape == want.shape and rel_err(got, want) <= tol
    tub = et_models.TubeletEmbedding(3, 768, (2, 16, 16))
    tub.conv.weight.data = 0.05 * torch.randn(tub.conv.weight.shape, generator=g)
    tub.conv.bias.data = torch.randn(768, generator=g)
    v = torch.randn(2, 8, 3, 64, 48, generator=g)
    want = F.conv3d(v.permute(0, 2, 1, 3, 4), tub.conv.weight.detach(), tub.conv.bias.detach(), stride=(2, 16, 16))
    want = want.flatten(-2).permute(0, 2, 3, 1)
    got = tub.to(DEV).to(dtype)(v.to(DEV)).float().cpu()
    assert got.shape == want.shape and rel_err(got, want) <= tol


def test_classifier_with_odd_class_count_on_the_tensor_core_path():
    """A 97-class head (EPIC-Kitchens verbs) in bf16: CountedLinear pads the feature count to the 16-byte store width."""
    from eventful_transformer.counting import CountedLinear

    g = torch.Generator().manual_seed(73)
    lin = CountedLinear(768, 97)
    lin.weight.data = 0.05 * torch.randn(97, 768, generator=g)
    lin.bias.data = torch.randn(97, generator=g)
    x = torch.randn(12, 768, generator=g)
    want = F.linear(x, lin.weight.detach(), lin.bias.detach())
    with torch.inference_mode():
        got = lin.to(DEV).to(torch.bfloat16)(x.to(DEV).to(torch.bfloat16)).float().cpu()
    assert got.shape == (12, 97) and rel_err(got, want) <= 1.5e-2
