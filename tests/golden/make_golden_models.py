#!/usr/bin/env python
"""
Generates tests/golden/model_*.npz by running the UNMODIFIED reference models (/root/reference/models/vivit.py and the
pre-backbone part of models/vitdet.py) on seeded inputs, CPU fp32.  Build container only:
    python tests/golden/make_golden_models.py
Stored: the seeded parameters (so the test loads exactly these), the input video / images, and the reference outputs.
detectron2 is not installed; models/vitdet.py is imported with detectron2 stubbed (only its embedding / preprocessing classes
and ViTBackbone are exercised -- SURVEY.md 8(c)).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ET_REFERENCE", "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot", "detectron2", "detectron2.config", "detectron2.structures"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["detectron2.config"].LazyConfig = object
sys.modules["detectron2.config"].instantiate = lambda *a, **k: None
sys.modules["detectron2.structures"].ImageList = object
sys.path.insert(0, HERE)
sys.path.insert(0, REFERENCE)

from model_cases import VIVIT_TINY, VIVIT_TINY_ATS, VITDET_STEM_TINY, seeded_state, vivit_video, vitdet_frames  # noqa: E402

from eventful_transformer import modules as ref_modules  # noqa: E402
from eventful_transformer import policies as ref_policies  # noqa: E402
from models import vitdet as ref_vitdet  # noqa: E402
from models import vivit as ref_vivit  # noqa: E402

assert ref_vivit.__file__.startswith(REFERENCE)


def set_policies(model, k):
    for cls in (ref_modules.SimpleSTGTGate, ref_modules.TokenDeltaGate, ref_modules.TokenGate):
        for gate in model.modules_of_type(cls):
            gate.policy = ref_policies.TokenNormTopK(k=k)


def make_vivit(cfg=VIVIT_TINY, name="model_vivit_tiny"):
    model = ref_vivit.FactorizedViViT(**cfg["model"]).eval()
    state = seeded_state(model.state_dict(), seed=cfg["seed"])
    model.load_state_dict(state, strict=True)
    set_policies(model, cfg["k"])
    video = vivit_video(cfg)
    blob = {f"param/{k}": v.numpy() for k, v in state.items()}
    blob["video"] = video.numpy()
    with torch.inference_mode():
        blob["probs"] = model(video).numpy()
        model.spatial_only = True
        blob["spatial"] = model(video).numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(name + ":", {k: v.shape for k, v in blob.items() if not k.startswith("param/")}, f"{os.path.getsize(path) / 1024:.0f} KiB")


def make_vitdet_stem():
    cfg = VITDET_STEM_TINY
    from eventful_transformer.backbones import ViTBackbone

    pre = ref_vitdet.ViTDetPreprocessing(cfg["input_shape"], cfg["normalize_mean"], cfg["normalize_std"])
    emb = ref_vitdet.LinearEmbedding(cfg["input_shape"][0], cfg["backbone_config"]["block_config"]["dim"], cfg["patch_size"])
    size = (cfg["input_shape"][1] // cfg["patch_size"], cfg["input_shape"][2] // cfg["patch_size"])
    backbone = ViTBackbone(input_size=size, **cfg["backbone_config"]).eval()
    holder = torch.nn.ModuleDict(dict(embedding=emb, backbone=backbone))
    state = seeded_state(holder.state_dict(), seed=cfg["seed"])
    holder.load_state_dict(state, strict=True)
    set_policies(backbone, cfg["k"])
    frames = vitdet_frames(cfg)
    blob = {f"param/{k}": v.numpy() for k, v in state.items()}
    blob["frames"] = frames.numpy()
    with torch.inference_mode():
        for t in range(frames.shape[0]):
            x = pre(ref_vitdet.as_float32(frames[t]))
            tokens = emb(x[None])
            blob[f"tokens_{t}"] = tokens.numpy()
            blob[f"out_{t}"] = backbone(tokens).numpy()
    path = os.path.join(HERE, "model_vitdet_stem_tiny.npz")
    np.savez_compressed(path, **blob)
    print("model_vitdet_stem_tiny:", {k: v.shape for k, v in blob.items() if not k.startswith("param/")}, f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    torch.set_num_threads(4)
    make_vivit()
    make_vivit(VIVIT_TINY_ATS, "model_vivit_tiny_ats")
    make_vitdet_stem()
