"""End-to-end GPU parity of ViTBackbone (bf16, CUDA kernels) vs the oracle, on the golden-fixture cases."""
import pytest
import torch

from cases import CASES, n_tokens
from golden_util import case_frames, case_params, load_golden, oracle_for, subsample
from gpu_util import DEV, build_gpu_backbone, gpu_trace, rel_err, rounded

pytestmark = pytest.mark.gpu
DT = torch.bfloat16


def run_gpu(case, params, frames, graph=False):
    model = build_gpu_backbone(case, params, DT)
    model.use_cuda_graph = graph
    outs, traces = [], []
    with torch.inference_mode():
        for x in frames:
            outs.append(model(x.to(DT).to(DEV)).float().cpu())
            traces.append(gpu_trace(model))
    return model, outs, traces


# K/V pooling is pinned in the oracle only so far: the CUDA package raises NotImplementedError for pool_size
@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if not c.get("pool_size")))
def test_backbone_matches_oracle_given_identical_index_sets(name):
    """
    Tolerance: given identical index sets, the bf16 CUDA output must be as close to the exact
    (fp32 arithmetic, same bf16-rounded weights and inputs) oracle as the reference's own bf16 arithmetic
    is, within a factor 2, and in any case within 4 % of the output range.
    """
    case = CASES[name]
    params = case_params(case)
    frames = [f.to(DT).float() for f in case_frames(case)]
    _, outs, traces = run_gpu(case, params, frames)
    exact = oracle_for(case, rounded(params, DT))
    forced_ok = not case.get("stgt")
    lowp = oracle_for(case, {k: v.to(DT) for k, v in params.items()}) if forced_ok else None
    with torch.inference_mode():
        for t, x in enumerate(frames):
            forced = traces[t] if forced_ok else None
            want = exact.forward(x.clone(), forced=forced)
            err = rel_err(outs[t], want)
            if lowp is not None:
                try:
                    ref_err = rel_err(lowp.forward(x.to(DT), forced=forced).float(), want)
                except RuntimeError:  # an op without a bf16 CPU kernel in this torch build
                    ref_err, lowp = None, None
            else:
                ref_err = None
            bound = 0.04 if ref_err is None else max(0.02, min(0.04, 2.0 * ref_err + 0.01))
            if not forced_ok and t > 0:
                bound = 0.15  # free-running selections may diverge (the method is lossy)
            assert err <= bound, f"{name} frame {t}: rel err {err:.4f} > {bound:.4f} (reference bf16 err {ref_err})"
            if forced is not None and t > 0 and case["policy"] is not None and case["policy"][0] != "threshold":
                # the CUDA selection and the oracle's free selection on its own (near-identical) inputs agree
                exact_free = {key: idx for key, idx in exact.trace}
                assert set(exact_free) == set(forced)


@pytest.mark.parametrize("name", ["tiny_vitdet", "tiny_vivit", "small_vitdet_b", "small_vitdet_tc", "tiny_dense", "vitdet_b_672"])
def test_first_frame_matches_reference_fixture(name):
    """Frame 0 needs no selection: compare straight against the committed reference outputs."""
    case, gold = CASES[name], load_golden(name)
    frames = case_frames(case)
    _, outs, _ = run_gpu(case, case_params(case), frames[:1])
    want = torch.from_numpy(gold["out_0"])
    got = subsample(outs[0]) if case.get("subsample") else outs[0]
    assert rel_err(got, want) < 0.04


@pytest.mark.parametrize("name", ["tiny_vitdet", "small_vitdet_b", "small_vitdet_tc", "tiny_vivit", "vitdet_b_672"])
def test_counters_match_reference_fixture(name):
    case, gold = CASES[name], load_golden(name)
    model = build_gpu_backbone(case, case_params(case), DT)
    model.counting()
    with torch.inference_mode():
        for t, x in enumerate(case_frames(case)):
            model.clear_counts()
            model(x.to(DT).to(DEV))
            got = model.total_counts()
            keys = {f[len(f"count_{t}_"):] for f in gold.files if f.startswith(f"count_{t}_")}
            for key in keys:
                assert int(got[key]) == int(gold[f"count_{t}_{key}"]), (t, key)
            assert {k for k, v in got.items() if v} <= keys


@pytest.mark.parametrize("name", ["tiny_vitdet", "small_vitdet_b", "small_vitdet_tc"])
def test_cuda_graph_replay_is_bit_identical_to_eager(name):
    case = dict(CASES[name], frames=6)
    params, frames = case_params(case), case_frames(dict(CASES[name], frames=6))
    _, eager, _ = run_gpu(case, params, frames, graph=False)
    model, graphed, _ = run_gpu(case, params, frames, graph=True)
    assert model._graph is not None
    for a, b in zip(eager, graphed):
        assert torch.equal(a, b)


def test_reset_restarts_the_stream_and_full_refresh_tracks_dense():
    case = CASES["small_vitdet_b"]
    params, frames = case_params(case), case_frames(case)
    full = dict(case, policy=("topk", dict(k=n_tokens(case))))
    dense = dict(case, block_class="Block", windowed_class="Block", policy=None)
    m_full, out_full, _ = run_gpu(full, params, frames)
    _, out_dense, _ = run_gpu(dense, params, frames)
    for a, b in zip(out_full, out_dense):
        assert rel_err(a, b) < 0.03  # k = N: eventful == dense up to rounding
    m_full.reset()
    with torch.inference_mode():
        again = m_full(frames[0].to(DT).to(DEV)).float().cpu()
    assert torch.equal(again, out_full[0])


def test_frame_pipeline_overlapped_copies_match_direct_calls():
    """et_pipeline.FramePipeline (upload / compute / download on two streams) returns exactly the direct results."""
    import et_pipeline

    case = dict(CASES["small_vitdet_b"], frames=6)
    params, frames = case_params(case), case_frames(case)
    _, direct, _ = run_gpu(case, params, frames)
    model = build_gpu_backbone(case, params, DT)
    host = [f.to(DT).pin_memory() for f in frames]
    pipe = et_pipeline.FramePipeline(model, tuple(host[0].shape), DT, torch.device(DEV))
    got = []
    with torch.inference_mode():
        for t, f in enumerate(host):
            prev = pipe.step(f, host[t + 1] if t + 1 < len(host) else None)
            if prev is not None:
                got.append(prev.float().clone())
        got.append(pipe.flush().float().clone())
    assert len(got) == len(direct)
    for a, b in zip(got, direct):
        assert torch.equal(a, b)
