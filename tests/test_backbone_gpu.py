"""End-to-end GPU parity of ViTBackbone (bf16, CUDA kernels) vs the oracle, on the golden-fixture cases."""
import pytest
import torch

from cases import CASES, n_tokens
from golden_util import case_frames, case_params, load_golden, oracle_for, subsample
from gpu_util import (DEV, allclose_report, ats_trace, build_gpu_backbone, elem_err, gpu_trace, record, rel_err, rounded,
                      selection_agreement)

pytestmark = pytest.mark.gpu
DT = torch.bfloat16


def run_gpu(case, params, frames, graph=False):
    model = build_gpu_backbone(case, params, DT, cast=case.get("matmul_2_cast"))
    model.use_cuda_graph = graph
    outs, traces = [], []
    with torch.inference_mode():
        for x in frames:
            outs.append(model(x.to(DT).to(DEV)).float().cpu())
            traces.append({**gpu_trace(model), **ats_trace(model)})
    return model, outs, traces


def _compare_selections(name, t, oracle, forced, case, tag):
    """
    Every policy-driven gate: the CUDA selection (replayed into the oracle as `forced`) against the selection the
    oracle's policy makes on its OWN inputs at that gate (same history, so the inputs differ only by arithmetic
    noise).  A token may be in one set and not the other only if its norm sits at the decision boundary.
    Returns {gate key: (overlap, worst boundary distance)}.
    """
    thr = case["policy"][1].get("threshold") if case["policy"][0] == "threshold" else None
    stats = {}
    for key, free_index, norm in oracle.free_trace:
        stats[key] = selection_agreement(forced[key], free_index, norm, threshold=thr)
    gates_run = {key for key in forced if key[1] != "ats"}
    assert set(stats) == gates_run, f"{name} frame {t} ({tag}): gates compared {sorted(stats)} != gates run {sorted(gates_run)}"
    # adaptive token sampling: the CUDA path's sampled token set against the oracle's own sampling on its inputs
    for i, free_index, score in oracle.ats_free:
        overlap, edge = selection_agreement(forced[(i, "ats")], free_index, score)
        differing = round((1.0 - overlap) * free_index.shape[-1])
        assert differing <= 1 and edge <= 0.15, f"{name} frame {t} block {i}: ATS sets differ in {differing} tokens, boundary distance {edge:.3f}"
    return stats


# K/V pooling and the fp16 attention-value path have their own tests (test_variants_gpu.py)
@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if not c.get("pool_size") and c.get("matmul_2_cast") != "float16"))
def test_backbone_matches_oracle_given_identical_index_sets(name):
    """
    Activations: given identical index sets, the bf16 CUDA output must be as close to the exact
    (fp32 arithmetic, same bf16-rounded weights and inputs) oracle as the reference's own bf16 arithmetic
    is, within a factor 2, and in any case within 4 % of the output range; the per-element figures are recorded.
    Selections: at every gate the CUDA index set must agree with what the oracle's policy selects on its own inputs
    up to tokens whose norm sits at the k-th-norm boundary (bf16 noise); the first gate of the network sees
    bit-identical inputs, so there a disagreement is allowed only within a few bf16 ulps of the boundary.
    """
    case = CASES[name]
    params = case_params(case)
    frames = [f.to(DT).float() for f in case_frames(case)]
    _, outs, traces = run_gpu(case, params, frames)
    exact = oracle_for(case, rounded(params, DT))
    forced_ok = not case.get("stgt")
    # the reference's own bf16 arithmetic.  (With matmul_2_cast == model dtype the reference's `.to()` is a no-op, v then
    # ALIASES the QKV buffer and the v-gate state with it (blocks.py:561-566), so its v-gate never sees a delta; that quirk
    # only exists for cast == model dtype, which no shipped config uses, and is not reproduced: judge without the cast.)
    lowp = oracle_for(dict(case, matmul_2_cast=None), {k: v.to(DT) for k, v in params.items()}) if forced_ok else None
    policy_driven = case["policy"] is not None and forced_ok
    exact.record_free = policy_driven
    if lowp is not None:
        lowp.record_free = policy_driven
    worst_overlap, worst_edge, first_gate_edge = 1.0, 0.0, 0.0
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
            max_abs, max_mixed, rms = elem_err(outs[t], want)
            record("backbone_vs_oracle", case=name, frame=t, rel_err_of_range=err, reference_bf16_rel_err=ref_err,
                   max_abs_err=max_abs, max_err_over_rms_plus_abs=max_mixed, rms=rms)
            assert err <= bound, f"{name} frame {t}: rel err {err:.4f} > {bound:.4f} (reference bf16 err {ref_err})"
            if forced_ok:  # per element: |err| <= 6 % of rms(want) + 2^-5 |want|  (bf16 has 8 significant bits)
                allclose_report(outs[t], want, rtol=2.0 ** -5, atol_rms=0.06, what=f"{name} frame {t}")
            if policy_driven and t > 0:
                judge = lowp if lowp is not None else exact
                stats = _compare_selections(name, t, judge, forced, case, "bf16 reference arithmetic" if lowp is not None else "exact")
                for key, (overlap, edge) in stats.items():
                    worst_overlap, worst_edge = min(worst_overlap, overlap), max(worst_edge, edge)
                first = stats.get((0, "qkv_gate"))
                if first is not None and lowp is not None:
                    first_gate_edge = max(first_gate_edge, first[1])
                    assert first[0] >= 0.98 and first[1] <= 2.0 ** -5, f"{name} frame {t}: first gate {first}"
                # measured on B200 (profiles/r2_parity_measured.md): overlap >= 0.89 at k >= 64, boundary distance <= 0.074
                for key, (overlap, edge) in stats.items():
                    k_sel = forced[key].shape[-1]
                    differing = round((1.0 - overlap) * k_sel)
                    assert differing <= max(2, 0.12 * k_sel) and edge <= 0.15, \
                        f"{name} frame {t} gate {key}: {differing} of {k_sel} tokens differ, boundary distance {edge:.3f}"
    if policy_driven:
        record("selection_vs_oracle", case=name, worst_overlap=worst_overlap, worst_boundary_distance=worst_edge,
               first_gate_boundary_distance=first_gate_edge)


def test_benchmarked_config_eight_streams_cuda_graph():
    """
    BASELINE configs[1] exactly as bench.py times it: ViTDet-B 1024^2, k = 2048 of 4096, EIGHT streams batched along B,
    CUDA-graph replay (frames 2 and 3 are replayed from the graph; 20.8 waves of tc_apply, the persistent GEMM and 25
    padded windows per stream all in one pass).  Streams are independent, so each stream must (i) equal the
    single-stream run of the same video within bf16 noise and (ii) match the oracle run on that stream alone
    (checked for the first and the last stream; the oracle needs ~10 s per frame and stream on the host).
    """
    name = "vitdet_b_1024"
    case = CASES[name]
    params = case_params(case)
    streams = 8
    videos = [[f.to(DT).float() for f in case_frames(dict(case, seed=case["seed"] + 31 * s))] for s in range(streams)]
    frames = [torch.cat([videos[s][t] for s in range(streams)], dim=0) for t in range(case["frames"])]
    model, outs, traces = run_gpu(dict(case, batch=streams), params, frames, graph=True)
    assert model._graph is not None, "the CUDA-graph path was not taken"
    for s in (0, streams - 1):
        exact = oracle_for(case, rounded(params, DT))
        with torch.inference_mode():
            for t in range(case["frames"]):
                forced = {key: idx[s:s + 1] for key, idx in traces[t].items()}
                want = exact.forward(videos[s][t].clone(), forced=forced)
                err = rel_err(outs[t][s:s + 1], want)
                max_abs, max_mixed, rms = elem_err(outs[t][s:s + 1], want)
                record("benchmarked_config_8_streams", stream=s, frame=t, rel_err_of_range=err, max_abs_err=max_abs,
                       max_err_over_rms_plus_abs=max_mixed, rms=rms)
                assert err <= 0.04, f"stream {s} frame {t}: rel err {err:.4f}"
                allclose_report(outs[t][s:s + 1], want, rtol=2.0 ** -5, atol_rms=0.06, what=f"stream {s} frame {t}")
    # single-stream run of stream 3's video (eager, different GEMM / attention tiling): same selections up to ties
    _, single, single_traces = run_gpu(case, params, videos[3], graph=False)
    for t in range(case["frames"]):
        assert rel_err(outs[t][3:4], single[t]) <= 0.03
        for key, idx in single_traces[t].items():
            a, b = set(idx[0].tolist()), set(traces[t][key][3].tolist())
            record("eight_streams_vs_single_stream_selection", frame=t, gate=list(key), overlap=len(a & b) / len(a))
            assert len(a & b) >= 0.85 * len(a), (t, key, len(a & b))


@pytest.mark.parametrize("name", ["tiny_vitdet", "tiny_vivit", "small_vitdet_b", "small_vitdet_tc", "tiny_dense", "vitdet_b_672",
                                  "vitdet_b_1024"])
def test_first_frame_matches_reference_fixture(name):
    """Frame 0 needs no selection: compare straight against the committed reference outputs."""
    case, gold = CASES[name], load_golden(name)
    frames = case_frames(case)
    _, outs, _ = run_gpu(case, case_params(case), frames[:1])
    want = torch.from_numpy(gold["out_0"])
    got = subsample(outs[0]) if case.get("subsample") else outs[0]
    assert rel_err(got, want) < 0.04


@pytest.mark.parametrize("name", ["tiny_vitdet", "small_vitdet_b", "small_vitdet_tc", "tiny_vivit", "vitdet_b_672", "vitdet_b_1024",
                                  "tiny_ats_eventful", "tiny_ats_dense"])
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


@pytest.mark.parametrize("name", ["tiny_vitdet", "small_vitdet_b", "small_vitdet_tc", "vitdet_b_1024"])
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
