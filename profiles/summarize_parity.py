#!/usr/bin/env python
"""Turns the measured-parity log of a GPU test run (gpurun_out/parity_measured.jsonl, written by tests/gpu_util.record)
into profiles/r2_parity_measured.md.   usage: python profiles/summarize_parity.py [log.jsonl] > profiles/r2_parity_measured.md"""
import collections
import json
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_measured.jsonl"
rows = [json.loads(line) for line in open(path) if line.strip()]
by = collections.defaultdict(list)
for r in rows:
    by[r["test"]].append(r)


def worst(rs, key):
    vals = [r[key] for r in rs if r.get(key) is not None]
    return max(vals) if vals else float("nan")


print("# Measured parity (B200, `pytest -m gpu`; every figure below was recorded by the test that asserts on it)\n")
print("All errors are max |got - want| / max |want| over the tensor unless stated otherwise.\n")

rs = by.get("fp32_index_sets", [])
if rs:
    print("## fp32 CUDA path, free-running, against the reference fixtures (tests/test_variants_gpu.py)\n")
    print("| fixture | gate index sets identical to the reference | worst overlap | worst output error |")
    print("|---|---|---|---|")
    errs = collections.defaultdict(float)
    for r in by.get("fp32_vs_reference_fixture", []):
        errs[r["case"]] = max(errs[r["case"]], r["rel_err_of_range"])
    tot_e = tot_t = 0
    for r in sorted(rs, key=lambda r: r["case"]):
        tot_e += r["exact_sets"]
        tot_t += r["total_sets"]
        print(f"| {r['case']} | {r['exact_sets']} / {r['total_sets']} | {r['worst_overlap']:.4f} | {errs[r['case']]:.2e} |")
    print(f"| **all** | **{tot_e} / {tot_t}** | | |\n")

rs = by.get("backbone_vs_oracle", [])
if rs:
    print("## bf16 CUDA path against the oracle replaying the CUDA selections (tests/test_backbone_gpu.py)\n")
    print("| fixture | worst error (of range) | reference's own bf16 arithmetic | worst \\|err\\| / (rms + \\|want\\|) |")
    print("|---|---|---|---|")
    cases = collections.defaultdict(list)
    for r in rs:
        cases[r["case"]].append(r)
    for case in sorted(cases):
        c = cases[case]
        print(f"| {case} | {worst(c, 'rel_err_of_range'):.4f} | {worst(c, 'reference_bf16_rel_err'):.4f} | {worst(c, 'max_err_over_rms_plus_abs'):.4f} |")
    print()

rs = by.get("selection_vs_oracle", [])
if rs:
    print("## bf16 selections against the reference's bf16 arithmetic on the same history (per gate; tests/test_backbone_gpu.py)\n")
    print("| fixture | worst set overlap | worst distance of a differing token from the k-th-norm boundary | first gate (identical inputs) |")
    print("|---|---|---|---|")
    for r in sorted(rs, key=lambda r: r["case"]):
        print(f"| {r['case']} | {r['worst_overlap']:.4f} | {r['worst_boundary_distance']:.4f} | {r['first_gate_boundary_distance']:.4f} |")
    print()

rs = by.get("benchmarked_config_8_streams", [])
if rs:
    print("## The benchmarked configuration: ViTDet-B 1024^2, k = 2048, 8 streams, CUDA-graph replay (streams 0 and 7 vs the oracle)\n")
    print(f"worst error {worst(rs, 'rel_err_of_range'):.4f} of range, worst |err| / (rms + |want|) {worst(rs, 'max_err_over_rms_plus_abs'):.4f}, "
          f"max |err| {worst(rs, 'max_abs_err'):.3f} at rms {rs[0]['rms']:.2f}; "
          f"selections of the 8-stream run vs a single-stream run of the same video: overlap "
          f"{min(r['overlap'] for r in by.get('eight_streams_vs_single_stream_selection', [{'overlap': float('nan')}])):.4f} (worst gate)\n")

rs = by.get("delta_accumulator_drift", [])
if rs:
    e = rs[-1]["errors"]
    print("## 32 incremental frames of the `a_n v_n - p (v_n - dV)` accumulator form vs the oracle (tests/test_attention_gpu.py)\n")
    print(f"error after frame 0 / 1 / 8 / 16 / 32: {e[0]:.4f} / {e[1]:.4f} / {e[8]:.4f} / {e[16]:.4f} / {e[32]:.4f} (bf16 accumulator, rounded once per frame as in the reference)\n")

print("## Variants and model ends\n")
print("| test | case | dtype / model | worst error |")
print("|---|---|---|---|")
for name, key in (("matmul_2_cast", "rel_err_of_range"), ("matmul_2_cast_vs_fixture", "rel_err_of_range"),
                  ("kv_pooling_fp32_vs_fixture", "rel_err_of_range"), ("kv_pooling_bf16_vs_oracle", "rel_err_of_range"),
                  ("threshold_device_count", "rel_err_of_range"), ("matmul_buffer", "rel_err"), ("matmul_delta_accumulator", "rel_err"),
                  ("factorized_vivit_vs_reference", "max_abs_prob_err"), ("factorized_vivit_vs_reference", "spatial_rel_err"),
                  ("vitdet_stem_vs_reference", "out_rel_err")):
    groups = collections.defaultdict(list)
    for r in by.get(name, []):
        groups[(r.get("case", ""), r.get("model", r.get("dtype", "")))].append(r)
    for (case, dt), g in sorted(groups.items()):
        print(f"| {name} ({key}) | {case} | {dt} | {worst(g, key):.2e} |")
