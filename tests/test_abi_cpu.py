"""CPU-side checks of the boundary: the C-ABI library loads and exports what include/eventful_b200.h
declares, the Python surface mirrors the reference's, and nothing computes without a GPU."""
import os
import re
import sys
import types

import pytest
import torch

import et_synthetic as syn
from cases import CASES
from eventful_transformer import _native as native
from eventful_transformer import backbones, base, blocks, counting, modules, policies, utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "eventful_b200.h")).read()
    declared = set(re.findall(r"\b(et_[a-z0-9_]+)\s*\(", header))
    declared -= {"et_dtype", "et_status"}
    assert declared, "no declarations parsed"
    lib = native.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(native.exported_symbols()), declared ^ set(native.exported_symbols())
    assert lib.et_version() >= 100


def test_no_cpu_fallback():
    model = backbones.ViTBackbone(**syn.backbone_kwargs(CASES["tiny_vitdet"]["cfg"], (7, 7)))
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 49, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        policies.TokenNormTopK(k=2)(torch.zeros(1, 4, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        modules.TokenBuffer()(torch.zeros(1, 4, 8), None).sum() if False else native.add(torch.zeros(8), torch.zeros(8))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libeventful_b200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        native.lib()


@pytest.mark.parametrize("name", ["tiny_vitdet", "tiny_vivit", "small_vitdet_b"])
def test_state_dict_keys_and_shapes(name):
    case = CASES[name]
    kw = syn.backbone_kwargs(case["cfg"], case["input_size"], block_class=case["block_class"],
                             has_class_token=case.get("has_class_token", False))
    model = backbones.ViTBackbone(**kw)
    want = syn.param_shapes(case["cfg"], case.get("has_class_token", False))
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    assert not list(model.named_buffers())  # gate / buffer state is attributes, not buffers
    for key, v in model.state_dict().items():  # zero init like the reference (nn.LayerNorm weight is 1)
        if not key.endswith("layer_norm.weight"):
            assert float(v.abs().sum()) == 0.0, key


def test_vitdet_b_surface():
    model = backbones.ViTBackbone(**syn.backbone_kwargs(syn.VITDET_B, (64, 64), matmul_2_cast="bfloat16"))
    assert len(model.state_dict()) == 169
    assert sum(v.numel() for v in model.state_dict().values()) == 85297664
    kinds = [type(b).__name__ for b in model.blocks]
    assert kinds == ["EventfulTokenwiseBlock", "EventfulTokenwiseBlock", "EventfulBlock"] * 4
    assert model.blocks[0].matmul_2_cast is None and model.blocks[2].matmul_2_cast == "bfloat16"
    assert model.blocks[0].relative_position.y_embedding.shape == (27, 64)
    assert model.blocks[2].relative_position.y_embedding.shape == (127, 64)
    # set_policies protocol (utils/misc.py:140-143)
    n = 0
    for cls in (modules.SimpleSTGTGate, modules.TokenDeltaGate, modules.TokenGate):
        for gate in model.modules_of_type(cls):
            gate.policy = policies.TokenNormTopK(k=2048)
            n += 1
    assert n == 36 + 8
    model.counting()
    assert all(m.count_mode for m in model.extended_modules())
    model.reset()
    assert isinstance(model.total_counts(), base.Counts)


def test_api_names_match_reference():
    assert blocks.LN_EPS == 1e-6
    for name in ("Block", "EventfulTokenwiseBlock", "EventfulMatmul1Block", "EventfulBlock"):
        assert hasattr(blocks, name)
    for name in ("SimpleSTGTGate", "TokenBuffer", "TokenGate", "TokenDeltaGate", "MatmulBuffer",
                 "MatmulDeltaAccumulator"):
        assert hasattr(modules, name)
    for name in ("TokenNormThreshold", "TokenNormTopK", "TokenNormTopFraction"):
        assert hasattr(policies, name)
    for name in ("CountedAdd", "CountedBias", "CountedConv", "CountedEinsum", "CountedLinear", "CountedMatmul"):
        assert hasattr(counting, name)
    for name in ("DropPath", "PositionEncoding", "RelativePositionEmbedding", "expand_col_index",
                 "expand_row_index"):
        assert hasattr(utils, name)
    for name in ("Counts", "ExtendedModule", "numeric_tuple", "dict_csv_header", "dict_csv_line", "dict_string"):
        assert hasattr(base, name)
    with pytest.raises(AssertionError):
        blocks.EventfulBlock(dim=32, heads=2, input_size=(8, 8), mlp_ratio=4, window_size=(4, 4))
    with pytest.raises(AssertionError):
        policies.TokenNormTopFraction(1.5)
    idx = torch.arange(6).view(2, 3)
    assert utils.expand_row_index(idx, (2, 4, 5, 7)).shape == (2, 4, 3, 7)
    assert utils.expand_col_index(idx, (2, 4, 5, 7)).shape == (2, 4, 5, 3)


def test_counts_arithmetic():
    c = base.Counts()
    c["x"] += 4
    d = (2 * c + c - 1) / 2
    assert d["x"] == 5.5 and (10 - c)["x"] == 6 and (-c)["x"] == -4
    assert base.dict_csv_header({"b": 1, "a": 2}) == "a,b" and base.dict_csv_line({"b": 1, "a": 2.5}) == "2.5,1"
    assert base.numeric_tuple(3, 2) == (3, 3) and base.numeric_tuple([1, 2], 2) == (1, 2)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_reference_models_construct_on_top_of_this_package():
    """models/vivit.py of the reference builds unchanged on this package (import-level drop-in)."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.append(REFERENCE)  # appended: `eventful_transformer` must still resolve to this package
    try:
        import importlib

        vivit = importlib.import_module("models.vivit")
        assert vivit.ViTBackbone is backbones.ViTBackbone
        model = vivit.FactorizedViViT(
            classes=10, input_shape=(3, 32, 224, 224), normalize_mean=[0.5] * 3, normalize_std=[0.5] * 3,
            spatial_config=dict(depth=2, position_encoding_size=[14, 14], block_class="EventfulBlock",
                                block_config=dict(dim=768, heads=12, mlp_ratio=4)),
            spatial_views=1, temporal_config=dict(depth=1, position_encoding_size=[16], block_class="Block",
                                                  block_config=dict(dim=768, heads=12, mlp_ratio=4)),
            temporal_stride=2, temporal_views=1, tubelet_shape=(2, 16, 16))
        assert any(isinstance(m, blocks.EventfulBlock) for m in model.modules())
    finally:
        sys.path.remove(REFERENCE)
        for mod in [m for m in sys.modules if m == "models" or m.startswith("models.") or m == "utils" or m.startswith("utils.")]:
            sys.modules.pop(mod, None)


def test_model_ends_keep_the_reference_state_dict_keys():
    """et_models.FactorizedViViT / ViTDetStem expose exactly the parameter names the reference models have (the fixtures store
    the reference's own state dicts), so converted checkpoints load unchanged."""
    import numpy as np

    import et_models
    from golden_util import GOLDEN_DIR
    from model_cases import VITDET_STEM_TINY, VIVIT_TINY

    gold = np.load(os.path.join(GOLDEN_DIR, "model_vivit_tiny.npz"))
    want = {k[len("param/"):]: gold[k].shape for k in gold.files if k.startswith("param/")}
    model = et_models.FactorizedViViT(**VIVIT_TINY["model"])
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == want
    assert model.temporal_model.backbone.blocks[0]._grid == (1, 4)  # 1-D token axis of the temporal sub-model
    gold = np.load(os.path.join(GOLDEN_DIR, "model_vitdet_stem_tiny.npz"))
    want = {k[len("param/"):]: gold[k].shape for k in gold.files if k.startswith("param/")}
    cfg = VITDET_STEM_TINY
    stem = et_models.ViTDetStem(cfg["backbone_config"], cfg["input_shape"], cfg["normalize_mean"], cfg["normalize_std"],
                                cfg["patch_size"])
    assert {k: tuple(v.shape) for k, v in stem.state_dict().items()} == want


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_reference_vitdet_stem_constructs_on_top_of_this_package():
    """models/vitdet.py of the reference (detectron2 stubbed: not installed) builds its embedding + backbone on this package."""
    stubs = ("matplotlib", "matplotlib.pyplot", "detectron2", "detectron2.config", "detectron2.structures")
    added = [name for name in stubs if name not in sys.modules]
    for name in added:
        sys.modules[name] = types.ModuleType(name)
    if "detectron2.config" in added:
        sys.modules["detectron2.config"].LazyConfig = object
        sys.modules["detectron2.config"].instantiate = lambda *a, **k: None
        sys.modules["detectron2.structures"].ImageList = object
    sys.path.append(REFERENCE)
    try:
        import importlib

        vitdet = importlib.import_module("models.vitdet")
        assert vitdet.ViTBackbone is backbones.ViTBackbone
        emb = vitdet.LinearEmbedding(3, 768, (16, 16))
        assert emb.conv.weight.shape == (768, 3, 16, 16)
        bb = vitdet.ViTBackbone(input_size=(64, 64), **syn.backbone_kwargs(syn.VITDET_B, (64, 64)).__class__(
            {k: v for k, v in syn.backbone_kwargs(syn.VITDET_B, (64, 64)).items() if k != "input_size"}))
        assert len(bb.blocks) == 12
    finally:
        sys.path.remove(REFERENCE)
        for mod in [m for m in sys.modules if m == "models" or m.startswith("models.") or m == "utils" or m.startswith("utils.")]:
            sys.modules.pop(mod, None)
        for name in added:
            sys.modules.pop(name, None)
