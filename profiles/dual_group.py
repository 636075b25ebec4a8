#!/usr/bin/env python
"""Stream groups on concurrent CUDA streams: G groups of 8 / G video streams, each group with its own backbone (own state,
own CUDA graph) on its own CUDA stream, stepped together.  While one group sits in a latency-bound phase (the single-CTA
selection tail of a gate, the partial last wave of a kernel), the other group's kernels can use the idle SMs.
Prints frames/s for G = 1 (the bench configuration), 2 and 4 at 8 and 16 streams in total."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))
import bench
import et_synthetic as syn

dev, dt = torch.device("cuda", 0), torch.bfloat16
torch.cuda.set_device(0)
grid, n, d, k = (64, 64), 4096, 768, 2048
for total in (8, 16):
    for groups in (1, 2, 4):
        per = total // groups
        models = [bench.make_backbone(grid, "EventfulBlock", "EventfulTokenwiseBlock", k, dev, dt) for _ in range(groups)]
        frames = [[f.to(dev) for f in bench.make_frames(list(range(g * per, (g + 1) * per)), n, d)] for g in range(groups)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(groups)]
        with torch.inference_mode():
            for g in range(groups):
                with torch.cuda.stream(streams[g]):
                    models[g].use_cuda_graph = True
                    for t in range(5):  # dense flush, warm-up, capture, replays
                        models[g](frames[g][t % bench.RING])
            torch.cuda.synchronize()
            steps = 20
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main = torch.cuda.current_stream()
            start.record(main)
            for s_ in streams:
                s_.wait_stream(main)
            for i in range(steps):
                for g in range(groups):
                    with torch.cuda.stream(streams[g]):
                        models[g](frames[g][(5 + i) % bench.RING])
            for s_ in streams:
                main.wait_stream(s_)
            stop.record(main)
            torch.cuda.synchronize()
        ms = start.elapsed_time(stop)
        print(f"{total} streams as {groups} group(s) of {per}: {steps * total / (ms * 1e-3):8.1f} frames/s  ({ms / steps:.2f} ms per step)")
        del models, frames
        torch.cuda.empty_cache()
