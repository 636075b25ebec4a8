#!/usr/bin/env python
"""Phase timestamps (%globaltimer, ns) inside gate_select_kernel via the et_debug_set(3, ptr) hook."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native
dev, dt = "cuda", torch.bfloat16
dbg = torch.zeros(8, dtype=torch.int64, device=dev)
for (B, N, D, k) in [(1, 4096, 768, 2048), (1, 1024, 768, 512), (4, 4096, 768, 2048)]:
    x = torch.randn(B, N, D, device=dev).to(dt); p = torch.randn(B, N, D, device=dev).to(dt)
    w = torch.randn(D, device=dev).to(dt); bb = torch.randn(D, device=dev).to(dt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for it in range(6):
        flush.zero_()
        dbg.zero_(); dbg[0] = 2 ** 62
        torch.cuda.synchronize()
        native.lib().et_debug_set(3, dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); native.gate_select(x, p=p, ln=(w, bb), k=k); e1.record()
        torch.cuda.synchronize()
        native.lib().et_debug_set(3, 0)
        d = dbg.tolist(); t0 = d[0]
        rows.append([round((v - t0) / 1e3, 2) for v in (d[1], d[2], d[3], d[4], d[6], d[7], d[5])] + [round(e0.elapsed_time(e1) * 1e3, 1)])
    print((B, N, D, k), "us since first CTA start: [norm phase done (max), select start, keys loaded, search done, counted+scanned, offsets known, end] event_us")
    for r in rows[2:]: print("   ", r)
