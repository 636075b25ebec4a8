#!/usr/bin/env python
"""Per-role cycle buckets of linear_tcgen05_kernel for the four ViTDet-B layer shapes (profiling build, make prof).

Run:  EVENTFUL_B200_LIB=eventful-transformer_b200/lib/libeventful_b200_prof.so python profiles/gemm_roles.py
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native
dev, dt = "cuda", torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NAMES = {0: ("TMA producer", {0: "prologue", 1: "wait empty slot", 2: "issue TMA"}),
         1: ("MMA issuer", {0: "prologue", 3: "wait first stage / (persistent) accumulator buffer", 1: "wait full stage", 2: "issue MMAs"}),
         2: ("epilogue warp 2", {0: "prologue", 4: "index lookup", 1: "wait accumulator", 2: "tcgen05.ld", 3: "bias/act/pack -> staging", 5: "staging barrier", 6: "row stores"})}
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
if len(sys.argv) > 2: native.lib().et_debug_set(9, int(sys.argv[2]))  # 1 = persistent kernel
for name, K, F, act in (("qkv", 768, 2304, 0), ("proj", 768, 768, 0), ("mlp1", 768, 3072, 1), ("mlp2", 3072, 768, 0)):
    x = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(F, K, device=dev) * 0.02).to(dt); bias = torch.randn(F, device=dev).to(dt)
    out = torch.empty(M, F, device=dev, dtype=dt)
    for _ in range(3): native.linear(x, w, bias, act=act, out=out)
    prof = torch.zeros(3 * 16, dtype=torch.int64, device=dev)
    flush.zero_(); torch.cuda.synchronize()
    native.lib().et_debug_set(7, prof.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); native.linear(x, w, bias, act=act, out=out); b.record(); torch.cuda.synchronize()
    native.lib().et_debug_set(7, 0)
    v = prof.view(3, 16).tolist()
    # number of CTAs = producer prologue count is unknown here; normalise by the MMA role's total / its per-CTA sum
    print(f"== {name}: M={M} K={K} F={F} act={act}: {a.elapsed_time(b) * 1e3:.1f} us (L2 flushed, profiling build); summed cycles over all CTAs")
    for role, (rname, names) in NAMES.items():
        tot = sum(v[role])
        print(f"  {rname}: total {tot}")
        for i, nm in names.items():
            print(f"      {nm:24s} {v[role][i]:12d}  {100.0 * v[role][i] / max(tot, 1):5.1f} %")
