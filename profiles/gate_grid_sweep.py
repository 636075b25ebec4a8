#!/usr/bin/env python
"""gate_select + gate_gather time vs the number of CTAs per SM (et_debug_set key 10), for 1 and 8 streams."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native
dev, dt = "cuda", torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=15):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / reps * 1e3
for B in (1, 8):
    N, D, k = 4096, 768, 2048
    x = torch.randn(B, N, D, device=dev).to(dt); p = torch.randn(B, N, D, device=dev).to(dt); xb = torch.randn(B, N, D, device=dev).to(dt)
    w = torch.randn(D, device=dev).to(dt); bb = torch.randn(D, device=dev).to(dt)
    for waves in (1, 2, 3, 4, 6, 8):
        native.lib().et_debug_set(10, waves)
        a = t(lambda: native.gate_select(x, p=p, ln=(w, bb), k=k))
        b = t(lambda: native.gate_select(x, p=p, xb=xb, want_sum=True, ln=(w, bb), k=k))
        print(f"streams={B} CTAs/SM={waves}: select(LN) {a:6.1f} us   select(add+LN) {b:6.1f} us")
    native.lib().et_debug_set(10, 2)
