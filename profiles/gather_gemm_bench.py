#!/usr/bin/env python
"""et_linear on pre-gathered rows vs et_linear_gather (TMA gather4 A operand, +state advance), and et_gate_select with /
without the c output, for the ViTDet-B gate sites at 1 and 8 streams (L2 flushed, CUDA events)."""
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
N, k = 4096, 2048
for B in (1, 8):
    src = torch.randn(B, N, 768, device=dev).to(dt)
    state = torch.zeros_like(src)
    idx = torch.stack([torch.randperm(N, device=dev)[:k] for _ in range(B)])
    idx_sorted = idx.sort(dim=1)[0].contiguous()
    rows = torch.gather(src, 1, idx[..., None].expand(-1, -1, 768)).contiguous()
    for name, F, act in (("qkv", 2304, 0), ("proj", 768, 0), ("mlp1", 3072, 1)):
        w = (torch.randn(F, 768, device=dev) * 0.02).to(dt); bias = torch.randn(F, device=dev).to(dt)
        buf = torch.zeros(B, N, F, device=dev, dtype=dt)
        dense = t(lambda: native.linear(rows, w, bias, act=act, out=buf, idx=idx))
        g = t(lambda: native.linear_gather(src, idx, w, bias, act=act, out=buf, idx=idx))
        gs = t(lambda: native.linear_gather(src, idx, w, bias, state=state, act=act, out=buf, idx=idx))
        gso = t(lambda: native.linear_gather(src, idx_sorted, w, bias, state=state, act=act, out=buf, idx=idx_sorted))
        print(f"B={B} {name:5s} dense-A {dense:6.1f} us | gather4 {g:6.1f} | gather4+state {gs:6.1f} | sorted index {gso:6.1f}")
    xa = torch.randn(B, N, 768, device=dev).to(dt); xb = torch.randn(B, N, 768, device=dev).to(dt)
    p = torch.randn(B, N, 768, device=dev).to(dt); c_all = torch.empty_like(xa)
    lw = torch.ones(768, device=dev, dtype=dt); lb = torch.zeros(768, device=dev, dtype=dt)
    a0 = t(lambda: native.gate_select(xa, p=p, ln=(lw, lb), k=k))
    a1 = t(lambda: native.gate_select(xa, p=p, ln=(lw, lb), k=k, c_out=c_all))
    b0 = t(lambda: native.gate_select(xa, p=p, xb=xb, want_sum=True, ln=(lw, lb), k=k))
    b1 = t(lambda: native.gate_select(xa, p=p, xb=xb, want_sum=True, ln=(lw, lb), k=k, c_out=c_all))
    gg = t(lambda: native.gate_gather(xa, idx, p=p, ln=(lw, lb)))
    print(f"B={B} gate_select LN {a0:6.1f} us, +c_out {a1:6.1f} | add+LN {b0:6.1f}, +c_out {b1:6.1f} | gate_gather {gg:6.1f}")
