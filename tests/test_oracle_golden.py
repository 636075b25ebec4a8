"""
Pins oracle/eventful_oracle.py to the reference: the fixtures under
tests/golden/ were produced by the UNMODIFIED reference (make_golden.py).
fp32 activations must be bit-identical, selected index sets identical, and the
closed-form op counters equal to the reference's own counters.
"""
import numpy as np
import pytest
import torch

import eventful_oracle as orc
from cases import CASES, GATES, n_tokens
from golden_util import case_frames, case_params, load_golden, oracle_for, subsample


def _forced_from(gold, t):
    forced = {}
    for f in gold.files:
        if f.startswith(f"raw_{t}_"):
            _, _, i, gate = f.split("_", 3)
            forced[(int(i), gate)] = torch.from_numpy(gold[f].astype(np.int64))
    return forced


def _check_ats(model, gold, t):
    """Adaptive token sampling: the stabilised index of every block equals the reference's, slot by slot."""
    seen = 0
    for (i, gate), index in model.trace:
        if gate == "ats":
            assert np.array_equal(index.numpy(), gold[f"ats_{t}_{i}"]), (t, i)
            seen += 1
    assert seen == sum(1 for f in gold.files if f.startswith(f"ats_{t}_"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_bitwise_given_reference_index_order(name):
    """Replaying the reference's own selection order, fp32 activations are bit-identical."""
    case, gold = CASES[name], load_golden(name)
    if case.get("stgt"):
        pytest.skip("SimpleSTGTGate has no forced-index input (modules.py:28)")
    model = oracle_for(case, case_params(case))
    with torch.inference_mode():
        for t, x in enumerate(case_frames(case)):
            y = model.forward(x.clone(), forced=_forced_from(gold, t))
            got = subsample(y) if case.get("subsample") else y
            assert torch.equal(got, torch.from_numpy(gold[f"out_{t}"])), (name, t)
            assert float(y.double().abs().sum()) == float(gold[f"out_abs_sum_{t}"])
            _check_ats(model, gold, t)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_fixture(name):
    """Free-running: identical index sets; activations equal up to GEMM row-order rounding."""
    case, gold = CASES[name], load_golden(name)
    params = case_params(case)
    psum = sum(float(v.double().abs().sum()) for v in params.values())
    assert psum == float(gold["param_abs_sum"]), "seeded parameter stream drifted"
    model = oracle_for(case, params)
    with torch.inference_mode():
        for t, x in enumerate(case_frames(case)):
            y = model.forward(x.clone())
            got = subsample(y) if case.get("subsample") else y
            want = torch.from_numpy(gold[f"out_{t}"])
            assert got.shape == want.shape
            tol = 2e-5 * max(1.0, float(want.abs().max()))
            assert (got - want).abs().max() <= tol, f"{name} frame {t}: {(got - want).abs().max()}"
            _check_ats(model, gold, t)
            seen = 0
            for (i, gate), index in model.trace:
                if gate == "ats":
                    continue
                key = f"idx_{t}_{i}_{gate}"
                assert key in gold.files
                assert np.array_equal(np.sort(index.numpy(), axis=-1), gold[key]), key
                seen += 1
            assert seen == sum(1 for f in gold.files if f.startswith(f"idx_{t}_"))


@pytest.mark.parametrize("name", ["tiny_vitdet", "tiny_vivit", "tiny_matmul1", "tiny_tokenwise", "small_vitdet_b"])
def test_closed_form_counters_match_reference(name):
    case, gold = CASES[name], load_golden(name)
    cfg, k, n = case["cfg"], case["policy"][1]["k"], n_tokens(case)
    total = {}
    for i in range(cfg["depth"]):
        windowed = i in cfg.get("window_indices", ())
        cls = case.get("windowed_class", "EventfulTokenwiseBlock") if windowed else case["block_class"]
        c = orc.incremental_counts(
            n, k, cfg["dim"], cfg["heads"], cfg["mlp_ratio"], cls,
            window=(cfg["window_size"] if windowed else None), grid=case["input_size"],
            rel=cfg.get("relative_embedding_size") is not None, batch=case["batch"])
        for key, v in c.items():
            total[key] = total.get(key, 0) + v
    total["add_flops"] += case["batch"] * n * cfg["dim"]  # position encoding add (utils.py:66)
    for t in range(1, case["frames"]):
        for key, v in total.items():
            ref = int(gold[f"count_{t}_{key}"]) if f"count_{t}_{key}" in gold.files else 0
            assert v == ref, (t, key, v, ref)


def test_select_topk_tie_rule():
    norm = torch.tensor([[1.0, 3.0, 2.0, 3.0, 2.0, 2.0, 0.0]])
    assert orc.select_topk(norm, 4).tolist() == [[1, 3, 2, 4]]
    assert orc.select_topk(norm, 2).tolist() == [[1, 3]]
    assert orc.select_topk(torch.zeros(1, 5), 3).tolist() == [[0, 1, 2]]
    assert orc.select_topk(norm, 7).sort().values.tolist() == [list(range(7))]
    with pytest.raises(RuntimeError):
        orc.select_topk(norm, 8)


def test_select_threshold_is_ascending_and_strict():
    norm = torch.tensor([[0.5, 1.0, 1.5, 0.2, 3.0]])
    assert orc.select_threshold(norm, 1.0).tolist() == [[2, 4]]
    assert orc.select_threshold(norm, 9.0).shape == (1, 0)
    with pytest.raises(AssertionError):
        orc.select_threshold(torch.zeros(2, 4), 0.1)


def test_invariants_first_frame_equals_dense_and_full_refresh():
    """frame 0 == dense bitwise; k == N tracks dense to fp32 rounding."""
    case = dict(CASES["tiny_vitdet"])
    params = case_params(case)
    frames = case_frames(case)
    dense = oracle_for(dict(case, block_class="Block", windowed_class="Block", policy=None), params)
    full = oracle_for(dict(case, policy=("topk", dict(k=n_tokens(case)))), params)
    with torch.inference_mode():
        for t, x in enumerate(frames):
            yd, yf = dense.forward(x.clone()), full.forward(x.clone())
            if t == 0:
                assert torch.equal(yd, yf)
            else:
                assert (yd - yf).abs().max() < 5e-5
