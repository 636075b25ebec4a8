#!/usr/bin/env python
"""Times et_linear for the four ViTDet-B layer shapes over tile widths and pipeline depths (L2 flushed, CUDA events)."""
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
for M in (2048, 8192, 16384):
  for name, K, F, act in (("qkv", 768, 2304, 0), ("proj", 768, 768, 0), ("mlp1", 768, 3072, 1), ("mlp2", 3072, 768, 0)):
    x = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(F, K, device=dev) * 0.02).to(dt); bias = torch.randn(F, device=dev).to(dt)
    out = torch.empty(M, F, device=dev, dtype=dt)
    res = {}
    native.lib().et_debug_set(8, 1); native.lib().et_debug_set(9, 2)
    for bn in (0, 64, 96, 128, 192, 256):
        for depth in (1, 2):
            native.lib().et_debug_set(1, bn); native.lib().et_debug_set(5, depth)
            res[(bn, depth)] = round(t(lambda: native.linear(x, w, bias, act=act, out=out)), 1)
    native.lib().et_debug_set(5, 0)
    native.lib().et_debug_set(9, 2)
    for bn in (128, 192, 256):  # 256-row CTA tiles (two accumulators share a W tile): reported as depth 3
        native.lib().et_debug_set(1, bn); native.lib().et_debug_set(8, 2)
        res[(bn, 3)] = round(t(lambda: native.linear(x, w, bias, act=act, out=out)), 1)
    native.lib().et_debug_set(8, 0); native.lib().et_debug_set(9, 1)
    for bn in (128, 192, 256):  # persistent kernel (ring across tiles, double-buffered TMEM): reported as depth 4
        native.lib().et_debug_set(1, bn)
        res[(bn, 4)] = round(t(lambda: native.linear(x, w, bias, act=act, out=out)), 1)
    native.lib().et_debug_set(1, 0); native.lib().et_debug_set(8, 1); native.lib().et_debug_set(9, 2)
    auto1 = round(t(lambda: native.linear(x, w, bias, act=act, out=out)), 1)
    native.lib().et_debug_set(8, 0); native.lib().et_debug_set(9, 0)
    auto = round(t(lambda: native.linear(x, w, bias, act=act, out=out)), 1)
    ref = round(t(lambda: torch.nn.functional.linear(x, w, bias)), 1)
    fl = 2.0 * M * K * F
    best = min(res, key=res.get)
    print(f"M={M} {name:5s} auto128 {auto1:6.1f} auto {auto:6.1f} us ({fl/auto/1e6:6.0f} TF/s)  best {best} {res[best]:6.1f} us  cuBLAS {ref:6.1f} us | " + " ".join(f"{k[0]}/{k[1]}:{v}" for k, v in sorted(res.items())))
