#!/usr/bin/env python
"""et_linear at M = 16384 / 8192 / 2048 for the four ViTDet-B layer shapes: CTA-pair kernel (cta_group::2) vs the
single-CTA kernels (auto) vs cuBLAS (torch.nn.functional.linear, no GELU), L2 flushed, CUDA events."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native
dev, dt = "cuda", torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=20):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / reps * 1e3
for M in (16384, 8192, 2048):
  for name, K, F, act in (("qkv", 768, 2304, 0), ("proj", 768, 768, 0), ("mlp1", 768, 3072, 1), ("mlp2", 3072, 768, 0)):
    x = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(F, K, device=dev) * 0.02).to(dt); bias = torch.randn(F, device=dev).to(dt)
    out = torch.empty(M, F, device=dev, dtype=dt)
    native.lib().et_debug_set(13, 2)
    single = t(lambda: native.linear(x, w, bias, act=act, out=out))
    native.lib().et_debug_set(13, 1)
    pair = t(lambda: native.linear(x, w, bias, act=act, out=out))
    native.lib().et_debug_set(13, 0)
    auto = t(lambda: native.linear(x, w, bias, act=act, out=out))
    ref = t(lambda: torch.nn.functional.linear(x, w, bias))
    fl = 2.0 * M * K * F
    print(f"M={M:6d} {name:5s} single-CTA {single:6.1f} us | CTA pair {pair:6.1f} us ({fl / pair / 1e6:5.0f} TFLOP/s) | auto {auto:6.1f} | cuBLAS {ref:6.1f} us")
