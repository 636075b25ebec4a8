#!/usr/bin/env python
"""Per-role cycle buckets of tc_stats_kernel (profiling build): cycles per 128-key tile, mean over CTAs.

Run:  EVENTFUL_B200_LIB=eventful-transformer_b200/lib/libeventful_b200_prof.so python profiles/tc_stats_roles.py
"""
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
idx = torch.randperm(n, device=dev)[:k].sort().values.view(1, -1).contiguous()
blk._attention_first(qkv, None)
for _ in range(3): blk._attention_incremental(qkv, idx)
prof = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
native.lib().et_debug_set(4, prof.data_ptr())
blk._attention_incremental(qkv, idx)
torch.cuda.synchronize()
native.lib().et_debug_set(4, 0)
ctas, tiles = (n // 128) * h, n // 128
v = prof.view(8, 16).tolist()[4:]
NAMES = {0: ("TMA producer", {0: "wait k_empty", 1: "issue"}),
         1: ("MMA issuer", {7: "wait q_full", 0: "wait k_full", 1: "wait s_empty", 2: "issue S (4 MMAs) + commits"}),
         2: ("softmax warp 2", {0: "wait s_full", 1: "2 x tcgen05.ld + release", 2: "scale / max / exp2 / sum", 3: "loop + bias_h load"})}
for role, (name, names) in NAMES.items():
    tot = sum(v[role][i] for i in names)
    print(f"  {name}: total {tot / ctas / tiles:7.0f} cycles per 128-key tile")
    for i, nm in names.items():
        print(f"      {nm:32s} {v[role][i] / ctas / tiles:7.0f}")
