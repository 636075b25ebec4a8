#!/usr/bin/env python
"""Per-phase cycle buckets of tc_window2_kernel's softmax role (warp 2), averaged over all CTAs.
Needs the profiling build:  make -C eventful-transformer_b200/csrc prof
Run:  EVENTFUL_B200_LIB=eventful-transformer_b200/lib/libeventful_b200_prof.so python profiles/window_phases.py [streams]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native
from eventful_transformer import blocks
dev, dt = "cuda", torch.bfloat16
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d, h = 768, 12
blk = blocks.EventfulTokenwiseBlock(dim=d, heads=h, input_size=(64, 64), mlp_ratio=4, relative_embedding_size=(64, 64), window_size=(14, 14)).to(dev).to(dt)
for prm in blk.parameters(): prm.data.normal_(0, 0.02)
qkv = torch.randn(streams, 4096, 3 * d, device=dev).to(dt)
NAMES = ["prologue (barriers, TMEM alloc, sync)", "E table + one-hot block", "wait Q, K (TMA)", "pad patch + fence + arrive", "wait U = Q E^T",
         "bias columns (shift, pack, store)", "wait S'", "row max (TMEM pass 1)", "exp2 + P stores (TMEM pass 2)", "wait V (TMA)",
         "V patch + fence + arrive", "wait O = P V", "epilogue stores", "final sync + TMEM release"]
for _ in range(3): blk._dense_attention(qkv)
prof = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
native.lib().et_debug_set(4, prof.data_ptr())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); blk._dense_attention(qkv); b.record(); torch.cuda.synchronize()
native.lib().et_debug_set(4, 0)
ctas = streams * 25 * 2 * h
v = prof.view(8, 16)[7].tolist()
print(f"tc_window2_kernel, {streams} stream(s): {a.elapsed_time(b) * 1e3:.1f} us (profiling build), {ctas} CTAs; cycles per CTA (softmax warp 2), total {sum(v) / ctas:.0f}")
for i, nm in enumerate(NAMES):
    print(f"  {nm:42s} {v[i] / ctas:8.0f}")
