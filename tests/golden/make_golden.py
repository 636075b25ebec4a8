#!/usr/bin/env python
"""
Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/eventful_transformer) on seeded synthetic inputs.

Run in the build container only (the reference tree does not exist on the GPU
box):   python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own,
so these fixtures are what pins oracle/eventful_oracle.py -- and through it the
CUDA path -- to the reference's behaviour.  Stored per case: every frame's
backbone output (fp32; sub-sampled for the 768-wide case), every policy-driven
gate's selected index set (sorted, and in the reference's own order), the reference's own op counters per
incremental frame, and a checksum of the seeded parameters.
"""

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("ET_REFERENCE", "/root/reference")

for name in ("matplotlib", "matplotlib.pyplot"):  # utils/image.py:1 imports it; unused on this path
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, HERE)
sys.path.insert(0, REFERENCE)
# The product package directory must NOT be on sys.path here: it is a regular package and would shadow
# the reference's namespace package `eventful_transformer`. Load the synthetic-input helper by file path.
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "et_synthetic", os.path.join(ROOT, "eventful-transformer_b200", "et_synthetic.py"))
syn = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(syn)
from cases import CASES, GATES, n_tokens  # noqa: E402

from eventful_transformer import backbones as ref_backbones  # noqa: E402
from eventful_transformer import modules as ref_modules  # noqa: E402
from eventful_transformer import policies as ref_policies  # noqa: E402

assert ref_backbones.__file__.startswith(REFERENCE), ref_backbones.__file__


def reference_backbone(case):
    cfg = case["cfg"]
    kw = syn.backbone_kwargs(
        cfg, case["input_size"], block_class=case["block_class"],
        windowed_class=case.get("windowed_class", "EventfulTokenwiseBlock"),
        matmul_2_cast=case.get("matmul_2_cast"), has_class_token=case.get("has_class_token", False),
        pool_size=case.get("pool_size"), ats_fraction=case.get("ats_fraction"),
    )
    if kw.get("windowed_class") is None:
        kw.pop("windowed_class", None)
    for flag in ("gate_before_ln", "stgt"):
        if case.get(flag):
            kw["block_config"][flag] = True
    model = ref_backbones.ViTBackbone(**kw)
    params = syn.seeded_params(cfg, seed=case["seed"], std=case["std"],
                               has_class_token=case.get("has_class_token", False))
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.eval()
    if case["policy"] is not None:
        kind, pk = case["policy"]
        cls = dict(topk=ref_policies.TokenNormTopK, threshold=ref_policies.TokenNormThreshold,
                   fraction=ref_policies.TokenNormTopFraction)[kind]
        for gate_cls in (ref_modules.SimpleSTGTGate, ref_modules.TokenDeltaGate, ref_modules.TokenGate):
            for gate in model.modules_of_type(gate_cls):  # utils/misc.py:140-143
                gate.policy = cls(**pk)
    return model, params


def subsample(t):
    return t[:, ::5, ::37].contiguous()


def run_case(name, case):
    torch.manual_seed(0)
    model, params = reference_backbone(case)
    frames = syn.token_stream(case["batch"], n_tokens(case), case["cfg"]["dim"], case["frames"],
                              seed=case["seed"] + 100, mode=case["stream"])
    trace = {}

    def hook_for(block, gate):
        def hook(_module, _inputs, output):
            index = output[-1]
            if index is not None:
                trace[(block, gate)] = index.clone()
        return hook

    for i, block in enumerate(model.blocks):
        for gate in GATES:
            if hasattr(block, gate):
                getattr(block, gate).register_forward_hook(hook_for(i, gate))

    blob = {}
    psum = sum(float(v.double().abs().sum()) for v in params.values())
    blob["param_abs_sum"] = np.float64(psum)
    model.counting()
    with torch.inference_mode():
        for t, x in enumerate(frames):
            trace.clear()
            model.clear_counts()
            y = model(x.clone())
            blob[f"out_{t}"] = (subsample(y) if case.get("subsample") else y).numpy().copy()
            blob[f"out_abs_sum_{t}"] = np.float64(float(y.double().abs().sum()))
            for (i, gate), index in trace.items():
                blob[f"idx_{t}_{i}_{gate}"] = np.sort(index.numpy(), axis=-1).astype(np.int32)
                blob[f"raw_{t}_{i}_{gate}"] = index.numpy().astype(np.int32)  # the reference's own order
            for i, block in enumerate(model.blocks):  # ATS: the stabilised token indices of every block (blocks.py:175-176)
                if getattr(block, "last_ats_indices", None) is not None:
                    blob[f"ats_{t}_{i}"] = block.last_ats_indices.numpy().astype(np.int32)
            counts = model.total_counts()
            for key, value in counts.items():
                blob[f"count_{t}_{key}"] = np.int64(value)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        run_case(name, case)
