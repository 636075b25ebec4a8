#!/usr/bin/env python
"""Bisects gate_select latency on the GPU: selection skipped (k=0 path is host-side, so use k=1 / threshold), sizes, dtypes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native

def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

dev, dt = "cuda", torch.bfloat16
for (B, N, D) in [(1, 4096, 768), (1, 1024, 768), (8, 4096, 768), (1, 4096, 64)]:
    x = torch.randn(B, N, D, device=dev).to(dt); p = torch.randn(B, N, D, device=dev).to(dt)
    w = torch.randn(D, device=dev).to(dt); bb = torch.randn(D, device=dev).to(dt)
    res = {}
    res["k=N/2 ln"] = t(lambda: native.gate_select(x, p=p, ln=(w, bb), k=N // 2))
    res["k=1 ln"] = t(lambda: native.gate_select(x, p=p, ln=(w, bb), k=1))
    res["k=N/2 noln"] = t(lambda: native.gate_select(x, p=p, k=N // 2))
    res["k=N/2 nop"] = t(lambda: native.gate_select(x, k=N // 2))
    res["gather"] = t(lambda: native.gate_gather(x, torch.arange(N // 2, device=dev).repeat(B, 1), p=p, ln=(w, bb)))
    res["add"] = t(lambda: native.add(x, p))
    res["empty launch (torch.empty+zero_)"] = t(lambda: torch.empty(16, device=dev).zero_())
    print((B, N, D), {k: round(v, 1) for k, v in res.items()})
