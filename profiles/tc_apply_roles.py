#!/usr/bin/env python
"""Per-role cycle buckets of tc_apply_kernel (DELTA), summed over all CTAs: where each warp role spends a tile.

Needs the profiling build:  make -C eventful-transformer_b200/csrc prof
Run:  EVENTFUL_B200_LIB=eventful-transformer_b200/lib/libeventful_b200_prof.so python profiles/tc_apply_roles.py
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
ROLES = {
    0: ("producer (warp 0)", {0: "wait k_empty (S'(t-2) done)", 1: "TMA issue", 2: "wait pv_done(t-2) for V"}),
    1: ("a_n.v_n MMA issuer (warp 1)", {8: "wait v_full", 3: "wait p_ready", 4: "issue a_n.v_n", 6: "commit", 7: "wait q_full"}),
    8: ("S' MMA issuer (warp 18), per 64-key tile", {0: "wait k_full", 1: "wait s_empty", 2: "issue S' (12 MMAs per tile pair)"}),
    9: ("p.Vd MMA issuer (warp 19)", {0: "wait v_full", 1: "wait ps_full", 2: "issue p.Vd + commit"}),
    2: ("softmax warp 4 (key half 0)", None),
    3: ("state mover (warp 2)", {0: "index loads", 1: "wait pv_done(t-1)", 2: "cp.async tile t+3", 3: "wait p_ready(t)",
                                 4: "read a_n chunks + release", 5: "state stores (STG.128)"}),
}
SM = {0: "wait s_full", 1: "tcgen05.ld + release S", 2: "exp2 / round", 3: "wait pv_done + an_free (t-1)", 4: "a_n tile stores",
      5: "fence.proxy + arrive p_ready", 11: "loop overhead", 12: "epilogue: wait o_full", 13: "epilogue: tcgen05.ld",
      14: "epilogue: add + stores"}
for order in ("ascending (what top-k emits)", "random"):
    idx = torch.randperm(n, device=dev)[:k]
    if order.startswith("asc"): idx = idx.sort().values
    idx = idx.view(1, -1).contiguous()
    blk.reset()
    blk._attention_first(qkv, None)
    for _ in range(3): blk._attention_incremental(qkv, idx)
    prof = torch.zeros(12 * 16, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    native.lib().et_debug_set(4, prof.data_ptr())
    native.lib().et_debug_set(6, 1)
    blk._attention_incremental(qkv, idx)
    torch.cuda.synchronize()
    ms = native.lib().et_debug_elapsed_ms()
    native.lib().et_debug_set(4, 0)
    native.lib().et_debug_set(6, 0)
    ctas, tiles = (n // 128) * h, k // 64
    v = prof.view(12, 16).tolist()
    print(f"== index order: {order}; apply launch {ms * 1e3:.1f} us (profiling build); cycles per 64-key tile, mean over {ctas} CTAs")
    for role, (name, names) in ROLES.items():
        names = names or SM
        tot = sum(v[role])
        print(f"  {name}: total {tot / ctas / tiles:7.0f} cycles/tile")
        for i, nm in names.items():
            print(f"      {nm:32s} {v[role][i] / ctas / tiles:7.0f}")
