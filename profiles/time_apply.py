#!/usr/bin/env python
"""tc_apply alone at the benchmark shape (8 streams), CUDA events bracketing the kernel: cluster-size experiment (key 12)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native  # noqa: E402
from eventful_transformer import blocks  # noqa: E402


def main():
    dev, dt = "cuda", torch.bfloat16
    lib = native.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    for streams in (8, 1):
        n, k = 4096, 2048
        blk = blocks.EventfulBlock(dim=768, heads=12, input_size=(64, 64), mlp_ratio=4, relative_embedding_size=(64, 64)).to(dev).to(dt).eval()
        for p in blk.parameters():
            p.data.normal_(0, 0.02)
        qkv = torch.randn(streams, n, 2304, generator=g).to(dt).to(dev)
        idx = torch.stack([torch.randperm(n, generator=g)[:k].sort()[0] for _ in range(streams)]).to(dev)
        blk._attention_first(qkv, None)
        for cluster in (1, 2, 4, 8):
            lib.et_debug_set(12, cluster)
            lib.et_debug_set(6, 1)
            for _ in range(2):
                blk._attention_incremental(qkv, idx)
            total, reps = 0.0, 10
            for _ in range(reps):
                flush.zero_()
                blk._attention_incremental(qkv, idx)
                total += lib.et_debug_elapsed_ms()
            lib.et_debug_set(6, 0)
            ms = total / reps
            gbs = streams * 2.0 * 12 * n * k * 2 / (ms * 1e-3) / 1e9
            print(f"streams={streams} cluster={cluster}: tc_apply {1e3 * ms:8.1f} us  {gbs:7.1f} GB/s of A-gate state traffic ({gbs / 6531.9:.3f} of HBM peak)")
        lib.et_debug_set(12, 1)
        del blk
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
