#!/usr/bin/env python
"""Per-tile clock64 stamps of one CTA of tc_apply_kernel (DELTA) via et_debug_set(4, ptr)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native
from eventful_transformer import blocks
dev, dt = "cuda", torch.bfloat16
n, d, h, k = 4096, 768, 12, 2048
blk = blocks.EventfulBlock(dim=d, heads=h, input_size=(64, 64), mlp_ratio=4, relative_embedding_size=(64, 64)).to(dev).to(dt)
for prm in blk.parameters(): prm.data.normal_(0, 0.02)
qkv = torch.randn(1, n, 3 * d, device=dev).to(dt)
idx = torch.randperm(n, device=dev)[:k].view(1, -1)
blk._attention_first(qkv, None)
for _ in range(3): blk._attention_incremental(qkv, idx)
dbg = torch.zeros(10 * 16, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
native.lib().et_debug_set(4, dbg.data_ptr())
blk._attention_incremental(qkv, idx)
torch.cuda.synchronize()
native.lib().et_debug_set(4, 0)
v = dbg.view(10, 16).tolist()
t0 = min(x for row in v for x in row if x > 0)
names = ["prod slot free", "mma kv landed", "mma S buf free", "mma a_n ready", "mma p landed", "sm S ready", "sm computed", "sm prevPV done", "sm published", "sm wb read"]
print("cycles since first stamp, tiles 0..11")
for nme, row in zip(names, v):
    print(f"{nme:16s}", [x - t0 if x else None for x in row[:12]])
