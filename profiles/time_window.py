#!/usr/bin/env python
"""Windowed attention alone at the benchmark shape: second- vs first-generation tcgen05 kernel (CUDA events, L2 flushed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
from eventful_transformer import _native as native  # noqa: E402
from eventful_transformer import blocks  # noqa: E402


def main():
    dev, dt = "cuda", torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for streams in (1, 8):
        blk = blocks.EventfulTokenwiseBlock(dim=768, heads=12, input_size=(64, 64), mlp_ratio=4, relative_embedding_size=(64, 64),
                                            window_size=(14, 14)).to(dev).to(dt).eval()
        for p in blk.parameters():
            p.data.normal_(0, 0.02)
        qkv = torch.randn(streams, 4096, 2304, device=dev).to(dt)
        for gen in (2, 1):
            native.lib().et_debug_set(11, gen)
            for _ in range(3):
                blk._dense_attention(qkv)
            torch.cuda.synchronize()
            total, reps = 0.0, 20
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                blk._dense_attention(qkv)
                b.record()
                torch.cuda.synchronize()
                total += a.elapsed_time(b)
            flops = 4.0 * streams * 25 * 12 * 196 * 196 * 64
            print(f"streams={streams} generation={gen}: {1e3 * total / reps:8.1f} us per call  ({flops / (total / reps * 1e-3) / 1e12:6.1f} TFLOP/s useful)")
        native.lib().et_debug_set(11, 2)


if __name__ == "__main__":
    main()
