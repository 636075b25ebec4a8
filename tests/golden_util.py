"""Helpers shared by the golden / parity tests."""
import os

import numpy as np
import torch

import et_synthetic as syn
from cases import CASES, n_tokens  # noqa: F401

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def case_params(case, dtype=torch.float32):
    return syn.seeded_params(case["cfg"], seed=case["seed"], std=case["std"], dtype=dtype,
                             has_class_token=case.get("has_class_token", False))


def case_frames(case, dtype=torch.float32):
    return syn.token_stream(case["batch"], n_tokens(case), case["cfg"]["dim"], case["frames"],
                            seed=case["seed"] + 100, mode=case["stream"], dtype=dtype)


def oracle_for(case, params):
    import eventful_oracle as orc

    cfg = case["cfg"]
    wclass = case.get("windowed_class", "EventfulTokenwiseBlock")
    model = orc.OracleBackbone(
        params, depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"], input_size=case["input_size"],
        position_encoding_size=cfg["position_encoding_size"], mlp_ratio=cfg["mlp_ratio"],
        block_class=case["block_class"], windowed_class=wclass,
        window_indices=cfg.get("window_indices", ()), window_size=cfg.get("window_size"),
        relative_embedding_size=cfg.get("relative_embedding_size"),
        has_class_token=case.get("has_class_token", False),
        matmul_2_cast=case.get("matmul_2_cast"),
        windowed_matmul_2_cast=(None if case.get("matmul_2_cast") else "same"),
        gate_before_ln=case.get("gate_before_ln", False), stgt=case.get("stgt", False),
        pool_size=case.get("pool_size"), ats_fraction=case.get("ats_fraction"),
        windowed_pool_size=(None if case.get("pool_size") else "same"),  # configs/*/vitdet_vid/_spatial.yml
    )
    if case["policy"] is not None:
        model.set_policy(case["policy"][0], **case["policy"][1])
    return model


def subsample(t):
    return t[:, ::5, ::37].contiguous()
