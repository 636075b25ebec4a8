#!/usr/bin/env python
"""
bench.py -- ViTDet-B Eventful backbone throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

A step = one incremental frame (t >= 1) of every stream of the rank (default 8 concurrent streams per GPU, batched
along B; `single_stream` in the JSON line is the one-stream latency case) through ViTBackbone.forward:
ViTDet-B, 1024x1024 input -> 4096 tokens, TokenNormTopK k = 2048, windowed EventfulTokenwiseBlock x8 +
global EventfulBlock x4, bf16, random-init weights, synthetic token video x_t = x_0 + 0.1 t eps_t.
Streams are independent (own gate state): ranks shard streams, no collective in the hot path
(NCCL only gathers outputs after the timed region).  Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "eventful-transformer_b200"))

import et_pipeline  # noqa: E402
import et_streams  # noqa: E402
import et_synthetic as syn  # noqa: E402

METRIC = "ViTDet-B Eventful backbone frames/s (1024x1024, k=2048 of 4096 tokens, incremental frames)"
UNIT = "frames/s"
RING = 8  # distinct synthetic frames cycled through (consecutive frames always differ)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("ET_BENCH_STREAMS", "8")),
                    help="concurrent video streams per GPU, batched along B (BASELINE configs[4]: 64 streams over 8 GPUs "
                         "= 8 per GPU); the single-stream latency case is reported next to it as `single_stream`")
    ap.add_argument("--k", type=int, default=2048)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip dense comparators, kernel leg, other configs and CPU baseline")
    ap.add_argument("--total-streams", type=int, default=64,
                    help="BASELINE configs[4]: total concurrent streams partitioned over the GPUs (strong-scaling leg)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        rows = [r.split(",") for r in open(self.file.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.file.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------
# models and inputs
# ------------------------------------------------------------------------------------------
def make_backbone(grid, block_class, windowed_class, k, device, dtype):
    from eventful_transformer import backbones, modules, policies

    kw = syn.backbone_kwargs(syn.VITDET_B, grid, block_class=block_class, windowed_class=windowed_class)
    model = backbones.ViTBackbone(**kw)
    model.load_state_dict(syn.seeded_params(syn.VITDET_B, seed=0, std=0.02), strict=True)
    model = model.to(device).to(dtype).eval()
    if k is not None:
        for cls in (modules.SimpleSTGTGate, modules.TokenDeltaGate, modules.TokenGate):
            for gate in model.modules_of_type(cls):
                gate.policy = policies.TokenNormTopK(k=k)
    return model


def make_frames(stream_ids, n_tokens, dim):
    """RING frames of shape (len(stream_ids), N, D); each stream's video depends only on its global id."""
    per_stream = [syn.token_stream(1, n_tokens, dim, RING, seed=et_streams.stream_seed(100, sid), mode="drift",
                                   dtype=torch.bfloat16) for sid in stream_ids]
    return [torch.cat([frames[t] for frames in per_stream], dim=0) for t in range(RING)]


def timed_steps(step, steps, dist_ctx, finalize=None):
    """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks (ms).
    `finalize` (optional) runs before the stop event, e.g. to make the launching stream wait for side-stream copies."""
    if dist_ctx is not None:
        dist_ctx.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        step(i)
    if finalize is not None:
        finalize()
    stop.record()
    torch.cuda.synchronize()
    ms = torch.tensor([start.elapsed_time(stop)], device="cuda")
    if dist_ctx is not None:
        dist_ctx.all_reduce(ms, op=dist_ctx.ReduceOp.MAX)
        dist_ctx.barrier()
    return float(ms.item())


def run_stream_model(model, frames_dev, warmup, steps, dist_ctx, use_graph):
    """Dense flush + warm-up, then the device-resident timed loop. Returns (ms_total, launches_per_step)."""
    from eventful_transformer import _native as native

    model.reset()
    model.use_cuda_graph = use_graph
    with torch.inference_mode():
        model(frames_dev[0])
        t = 1
        c0 = native.lib().et_launch_count()
        model(frames_dev[t % RING])  # first incremental frame, eager: counts the launches of one step
        per_step = native.lib().et_launch_count() - c0
        t += 1
        for _ in range(max(3, warmup) - 1):
            model(frames_dev[t % RING])
            t += 1
        base = t

        def step(i):
            model(frames_dev[(base + i) % RING])

        ms = timed_steps(step, steps, dist_ctx)
    return ms, per_step, base + steps


# ------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (reported inside the same JSON line as `configs`)
# ------------------------------------------------------------------------------------------
VIVIT_B_SPATIAL = dict(depth=12, dim=768, heads=12, mlp_ratio=4, position_encoding_size=(14, 14))  # configs/models/vivit_b_kinetics400.yml
VIVIT_B_EPIC = dict(depth=12, dim=768, heads=12, mlp_ratio=4, position_encoding_size=(20, 20))     # configs/models/vivit_b_epic_kitchens.yml


def build_model(cfg, grid, device, dtype, policy=None, block_class="EventfulBlock", windowed_class="EventfulTokenwiseBlock",
                has_class_token=False, cast=None, pool=None):
    from eventful_transformer import backbones, modules

    kw = syn.backbone_kwargs(cfg, grid, block_class=block_class, windowed_class=windowed_class, matmul_2_cast=cast,
                             has_class_token=has_class_token, pool_size=pool)
    model = backbones.ViTBackbone(**kw)
    model.load_state_dict(syn.seeded_params(cfg, seed=0, std=0.02, has_class_token=has_class_token), strict=True)
    model = model.to(device).to(dtype).eval()
    if policy is not None:
        for cls in (modules.SimpleSTGTGate, modules.TokenDeltaGate, modules.TokenGate):
            for gate in model.modules_of_type(cls):
                gate.policy = policy()
    return model


def side_config(label, cfg, grid, batch, device, dtype, policy, steps, warmup=3, has_class_token=False, cast=None, pool=None,
                stream_mode="drift"):
    """One of the other BASELINE configurations: device-resident inputs, CUDA-graph replay, incremental frames only."""
    n = grid[0] * grid[1] + int(has_class_token)
    model = build_model(cfg, grid, device, dtype, policy=policy, has_class_token=has_class_token, cast=cast, pool=pool)
    frames = [f.to(device) for f in syn.token_stream(batch, n, cfg["dim"], RING, seed=7, mode=stream_mode, dtype=dtype)]
    ms, per_step, _ = run_stream_model(model, frames, warmup, steps, None, True)
    out = dict(workload=label, value=round(steps * batch / (ms * 1e-3), 2), unit=UNIT, ms_per_step=round(ms / steps, 4),
               batch=batch, tokens=n, dtype=str(dtype).replace("torch.", ""), launches_per_step=int(per_step), steps=steps,
               cuda_graph=model._graph is not None)
    selected = {}
    for i, blk in enumerate(model.blocks):
        gate = getattr(blk, "qkv_gate", None)
        if gate is not None and gate.last_index is not None:
            selected[i] = int(gate.last_index.shape[-1])
    if selected:
        out["selected_tokens_last_frame"] = dict(min=min(selected.values()), max=max(selected.values()))
    del model, frames
    torch.cuda.empty_cache()
    return out


def other_configs(device, steps):
    """BASELINE.json configs[0], [2], [3] on this GPU (configs[1] is the headline, configs[4] the stream legs)."""
    from eventful_transformer import policies

    out = {}
    topk = lambda k: (lambda: policies.TokenNormTopK(k=k))  # noqa: E731
    with torch.inference_mode():
        # configs[0]: ViTDet-B 672^2, 1 stream, k = 512 of 1764 -- in fp32 (the dtype the reference's CPU run uses), in the
        # reference's own timed CUDA setting (fp32 model, fp16 attention-value path) and in bf16
        for tag, dt, cast in (("c0_vitdet_b_672_fp32", torch.float32, None), ("c0_vitdet_b_672_fp32_fp16av", torch.float32, "float16"),
                              ("c0_vitdet_b_672_bf16", torch.bfloat16, None)):
            out[tag] = side_config("ViTDet-B 672x672, 1 stream, TopK k=512 of 1764" + (f", matmul_2_cast={cast}" if cast else ""),
                                   syn.VITDET_B, (42, 42), 1, device, dt, topk(512), steps, cast=cast)
        # configs[2]: ViViT-B spatial sub-model, Kinetics-400 shape: 12 views batched, 196 + class token, k = 64
        out["c2_vivit_b_k400_spatial"] = side_config(
            "ViViT-B spatial encoder (K400 shape): 12 views x 197 tokens per step, TopK k=64", VIVIT_B_SPATIAL, (14, 14), 12, device,
            torch.bfloat16, topk(64), steps, has_class_token=True)
        # configs[3]: ViViT-B EPIC-Kitchens shape: 320^2 -> 400 + class token, threshold policy (batch 1, device-side counts)
        for thr in (0.2, 1.0, 5.0):  # the reference's threshold sweep values
            out[f"c3_vivit_b_epic_threshold_{thr}"] = side_config(
                f"ViViT-B spatial encoder (EPIC-Kitchens shape): 1 view x 401 tokens per step, TokenNormThreshold {thr}",
                VIVIT_B_EPIC, (20, 20), 1, device, torch.bfloat16, (lambda thr=thr: policies.TokenNormThreshold(threshold=thr)),
                steps, has_class_token=True, stream_mode="patch")  # static background + a moving block of changed tokens
    return out


def k_sweep(grid, streams, frames_dev, device, dtype, steps, dense_fps):
    """configs[4]: frames/s at k = 512 ... 4096 for one stream group on this GPU, next to the dense (ungated) model."""
    out = {}
    for k in (512, 1024, 2048, 3072, 4096):
        model = make_backbone(grid, "EventfulBlock", "EventfulTokenwiseBlock", k, device, dtype)
        ms, _, _ = run_stream_model(model, frames_dev, 3, steps, None, True)
        fps = steps * streams / (ms * 1e-3)
        out[str(k)] = dict(value=round(fps, 2), ms_per_step=round(ms / steps, 4),
                           speedup_vs_fused_dense=None if dense_fps is None else round(fps / dense_fps, 3))
        del model
        torch.cuda.empty_cache()
    return out


def partitioned_streams_leg(grid, total_streams, world, rank, device, dtype, k, steps, dist_ctx, bytes_per_stream):
    """
    configs[4]: a FIXED total of concurrent streams partitioned over the ranks (strong scaling: 64 -> 64/32/16/8 per GPU).
    Skipped (with the reason) when a rank's share does not fit its memory.
    """
    mine = et_streams.partition_streams(total_streams, world, rank)
    free, total = torch.cuda.mem_get_info(device)
    need = len(mine) * bytes_per_stream * 1.08 + (8 << 30)
    fits = torch.tensor([1.0 if need < free else 0.0], device=device)
    if dist_ctx is not None:
        dist_ctx.all_reduce(fits, op=dist_ctx.ReduceOp.MIN)
    if float(fits.item()) < 1.0:
        return dict(total_streams=total_streams, streams_per_gpu=len(mine), unavailable=
                    f"needs ~{need / 2**30:.0f} GiB per GPU for {len(mine)} resident streams, {free / 2**30:.0f} GiB free")
    n, d = grid[0] * grid[1], syn.VITDET_B["dim"]
    try:
        model = make_backbone(grid, "EventfulBlock", "EventfulTokenwiseBlock", k, device, dtype)
        ring = 4
        per_stream = [syn.token_stream(1, n, d, ring, seed=et_streams.stream_seed(100, sid), mode="drift", dtype=dtype) for sid in mine]
        frames = [torch.cat([f[t] for f in per_stream], dim=0).to(device) for t in range(ring)]
        del per_stream
        model.reset()
        model.use_cuda_graph = True
        with torch.inference_mode():
            model(frames[0])
            for t in range(1, 4):
                model(frames[t % ring])
            ms = timed_steps(lambda i: model(frames[(4 + i) % ring]), steps, dist_ctx)
        out = dict(total_streams=total_streams, streams_per_gpu=len(mine), value=round(steps * total_streams / (ms * 1e-3), 2),
                   unit=UNIT, ms_per_step=round(ms / steps, 4), steps=steps, scaling="strong",
                   resident_state_gib_per_gpu=round(torch.cuda.max_memory_allocated(device) / 2**30, 1))
    except torch.OutOfMemoryError as exc:  # never take the box down: report and move on
        out = dict(total_streams=total_streams, streams_per_gpu=len(mine), unavailable=f"out of memory: {str(exc)[:120]}")
    model = frames = None
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------
# dense comparators
# ------------------------------------------------------------------------------------------
def eager_dense_forward(x, w, grid):
    """
    The reference's dense `Block` path restated with eager torch CUDA ops (materialised logits, two
    rel-pos adds, softmax, pad/concat windows) -- the comparator the 1.8x target is quoted against
    (SURVEY.md 7.3-2).  Bench-local; not part of the product.
    """
    import torch.nn.functional as F

    cfg = syn.VITDET_B
    d, h = cfg["dim"], cfg["heads"]
    dh = d // h
    gh, gw = grid
    b = x.shape[0]
    x = x + w["pos"]
    for i in range(cfg["depth"]):
        pre = f"blocks.{i}."
        skip = x
        y = F.layer_norm(x, (d,), w[pre + "input_layer_norm.weight"], w[pre + "input_layer_norm.bias"], 1e-6)
        y = F.linear(y, w[pre + "qkv.weight"], w[pre + "qkv.bias"])
        windowed = i in cfg["window_indices"]
        if windowed:
            wh, ww = cfg["window_size"]
            ph, pw = -gh % wh, -gw % ww
            y = y.view(b, gh, gw, 3 * d)
            if ph or pw:
                pad = w[pre + "qkv.bias"].view(1, 1, 1, -1)
                y = torch.cat([y, pad.expand(b, gh, pw, 3 * d)], 2)
                y = torch.cat([y, pad.expand(b, ph, gw + pw, 3 * d)], 1)
            th, tw = gh + ph, gw + pw
            y = y.view(b, th // wh, wh, tw // ww, ww, 3 * d).transpose(2, 3).reshape(-1, wh * ww, 3 * d)
            ah, aw = wh, ww
        else:
            ah, aw = gh, gw
        q, k, v = y.view(y.shape[0], y.shape[1], 3, h, dh).permute(2, 0, 3, 1, 4)
        a = (q / 8.0) @ k.transpose(-2, -1)
        ry, rx = w[pre + "rel_y"], w[pre + "rel_x"]
        qs = q.reshape(q.shape[0], h, ah, aw, dh)
        a = a.view(a.shape[0], h, ah, aw, ah, aw)
        a += torch.einsum("abhwc,hkc->abhwk", qs, ry).unsqueeze(-1)
        a += torch.einsum("abhwc,wkc->abhwk", qs, rx).unsqueeze(-2)
        a = a.view(a.shape[0], h, ah * aw, ah * aw).softmax(dim=-1)
        y = (a @ v).permute(0, 2, 1, 3).reshape(a.shape[0], ah * aw, d)
        if windowed:
            y = y.view(b, th // wh, tw // ww, wh, ww, d).transpose(2, 3).reshape(b, th, tw, d)[:, :gh, :gw]
            y = y.flatten(1, 2)
        x = F.linear(y, w[pre + "projection.weight"], w[pre + "projection.bias"]) + skip
        skip = x
        y = F.layer_norm(x, (d,), w[pre + "mlp_layer_norm.weight"], w[pre + "mlp_layer_norm.bias"], 1e-6)
        y = F.gelu(F.linear(y, w[pre + "mlp_1.weight"], w[pre + "mlp_1.bias"]))
        x = F.linear(y, w[pre + "mlp_2.weight"], w[pre + "mlp_2.bias"]) + skip
    return x


def eager_dense_weights(model, device, dtype):
    w = {k: v.detach() for k, v in model.state_dict().items()}
    w["pos"] = model.position_encoding.sized_encoding(1)
    for i, blk in enumerate(model.blocks):
        ry, rx = blk.relative_position.tables()
        w[f"blocks.{i}.rel_y"], w[f"blocks.{i}.rel_x"] = ry.to(dtype), rx.to(dtype)
    return w


# ------------------------------------------------------------------------------------------
# per-kernel leg: isolated timings -> roofline fractions
# ------------------------------------------------------------------------------------------
def time_call(fn, reps=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()  # evict L2 between repetitions (buffer > 126 MB)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / reps


def kernel_leg(streams, n, d, k, grid, pk):
    """Each hot kernel alone on step-shaped operands; algorithmic bytes / flops per launch as in DESIGN.md."""
    from eventful_transformer import _native as native
    from eventful_transformer import blocks

    dev, dt, e = "cuda", torch.bfloat16, 2
    b, h, dh = streams, 12, 64
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dt).to(dev)  # noqa: E731
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    x, p, xb = rnd(b, n, d), rnd(b, n, d), rnd(b, n, d)
    lnw, lnb = rnd(d), rnd(d)
    idx = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(b)]).to(dev)
    out = {}

    def add(name, ms, launches, bound, work):
        peak = pk["hbm"] if bound == "hbm" else pk["tensor_burst"]
        achieved = work / (ms * 1e-3) / (1e9 if bound == "hbm" else 1e12)
        out[name] = dict(ms=round(ms, 5), per_step=launches, bound=bound, achieved=round(achieved, 2), peak=peak,
                         unit="GB/s" if bound == "hbm" else "TFLOP/s", frac=round(achieved / peak, 4))

    ms = time_call(lambda: native.gate_select(x, p=p, ln=(lnw, lnb), k=k), flush=flush)
    add("gate_select(LN+delta-norm+topk)", ms, 24, "hbm", b * (2 * n * d * e + n * 4 + k * 8))
    ms = time_call(lambda: native.gate_select(x, p=p, xb=xb, want_sum=True, ln=(lnw, lnb), k=k), flush=flush)
    add("gate_select(add+LN+delta-norm+topk)", ms, 12, "hbm", b * (4 * n * d * e + n * 4 + k * 8))
    ms = time_call(lambda: native.gate_gather(x, idx, p=p, ln=(lnw, lnb)), flush=flush)
    add("gate_gather(LN rows + state advance)", ms, 36, "hbm", b * 3 * k * d * e)
    xs = rnd(b, k, d)
    for name, fin, fout, act, per in (("linear qkv", d, 3 * d, 0, 12), ("linear proj", d, d, 0, 12),
                                      ("linear mlp1+gelu", d, 4 * d, 1, 12), ("linear mlp2", 4 * d, d, 0, 12)):
        a_in = xs if fin == d else rnd(b, k, fin)
        wt, bias = rnd(fout, fin) * 0.02, rnd(fout)
        buf = torch.zeros(b, n, fout, dtype=dt, device=dev)
        ms = time_call(lambda: native.linear(a_in, wt, bias, act=act, out=buf, idx=idx), flush=flush)
        add(name + " (tcgen05, scatter epilogue)", ms, per, "tensor", 2.0 * b * k * fin * fout)
    qkv = rnd(b, n, 3 * d)
    wblk = blocks.EventfulTokenwiseBlock(dim=d, heads=h, input_size=grid, mlp_ratio=4, relative_embedding_size=(64, 64),
                                         window_size=(14, 14)).to(dev).to(dt)
    for prm in wblk.parameters():
        prm.data.normal_(0, 0.02, generator=None)
    ms = time_call(lambda: wblk._dense_attention(qkv), flush=flush)
    nwin, w2 = wblk._window_grid()[0], 196
    add("window_attention (+relpos bias)", ms, 8, "tensor", 4.0 * b * nwin * h * w2 * w2 * dh)
    gblk = blocks.EventfulBlock(dim=d, heads=h, input_size=grid, mlp_ratio=4, relative_embedding_size=(64, 64)).to(dev).to(dt)
    for prm in gblk.parameters():
        prm.data.normal_(0, 0.02)
    gblk._attention_first(qkv, None)
    ms = time_call(lambda: gblk._attention_incremental(qkv, idx), reps=10, flush=flush)
    add("global_attention delta: whole op (rel-pos tables + v-gate + tc_stats + tc_apply)", ms, 4, "tensor",
        2.0 * b * h * n * dh * (n + 3 * k))
    # the dominant kernel of the step, alone: CUDA events recorded by the library around its launch (same stream)
    lib = native.lib()
    lib.et_debug_set(6, 1)
    acc_ms = 0.0
    reps = 10
    for _ in range(reps):
        flush.zero_()
        gblk._attention_incremental(qkv, idx)
        acc_ms += lib.et_debug_elapsed_ms()
    lib.et_debug_set(6, 0)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(tpath) and n == 4096 and k == 2048:
        table = json.load(open(tpath))
        t = table.get("tc_apply_kernel<bf16,DELTA>" if b == 1 else f"tc_apply_kernel<bf16,DELTA>@{b}streams")
        if t is not None:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    # algorithmic bytes: the selected A-gate state columns are read once and written once, 2 H N k e per stream
    add("tc_apply_kernel (A-gate + delta accumulate on tcgen05)", acc_ms / reps, 4, "hbm", b * 2.0 * h * n * k * e)
    out["tc_apply_kernel (A-gate + delta accumulate on tcgen05)"]["traffic"] = traffic
    del flush
    return out


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm (the oracle port of the reference algorithm on the host cores)
# ------------------------------------------------------------------------------------------
def cpu_reference(grid, k, steps, warmup, budget_s, threads=None):
    if threads is not None:
        torch.set_num_threads(int(threads))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import eventful_oracle as orc  # the one place bench.py executes oracle/

    cfg = syn.VITDET_B
    n = grid[0] * grid[1]
    params = syn.seeded_params(cfg, seed=0, std=0.02)
    model = orc.OracleBackbone(
        params, depth=cfg["depth"], dim=cfg["dim"], heads=cfg["heads"], input_size=grid,
        position_encoding_size=cfg["position_encoding_size"], mlp_ratio=cfg["mlp_ratio"],
        block_class=orc.EVENTFUL, windowed_class=orc.TOKENWISE, window_indices=cfg["window_indices"],
        window_size=cfg["window_size"], relative_embedding_size=cfg["relative_embedding_size"],
        matmul_2_cast="bfloat16", windowed_matmul_2_cast=None)  # configs/time/vitdet_vid/_cpu.yml
    model.set_policy("topk", k=k)
    frames = syn.token_stream(1, n, cfg["dim"], 4, seed=1, mode="drift")
    cores = torch.get_num_threads()
    with torch.inference_mode():
        t0 = time.perf_counter()
        model.forward(frames[0])
        first = time.perf_counter() - t0
        times = []
        t_begin = time.perf_counter()
        i = 0
        while i < warmup + steps:
            t0 = time.perf_counter()
            model.forward(frames[(i + 1) % 4])
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            i += 1
            if time.perf_counter() - t_begin > budget_s and times:
                break
    sec = sum(times) / len(times)
    return dict(fps=1.0 / sec, sec_per_frame=sec, first_frame_s=first, timed=len(times), cores=cores)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    grid = (args.size // 16, args.size // 16)
    n, d = grid[0] * grid[1], syn.VITDET_B["dim"]
    workload = (f"ViTDet-B {args.size}x{args.size} Eventful backbone, {args.streams} stream(s)/GPU, "
                f"TokenNormTopK k={args.k} of {n} tokens, EventfulTokenwiseBlock x8 (14x14 windows) + EventfulBlock x4")

    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        r = cpu_reference(grid, args.k, args.steps, min(args.warmup, 1), budget_s=150.0)
        line = dict(metric=METRIC, value=round(r["fps"], 4), unit=UNIT, n_gpus=args.gpus, steps=r["timed"],
                    warmup=min(args.warmup, 1), ms_per_step=round(1e3 * r["sec_per_frame"], 2), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="f32",
                    data="synthetic", impl="reference",
                    config=dict(workload=workload.replace(f"{args.streams} stream(s)/GPU", "1 stream on the host CPU")),
                    cpu_baseline=dict(value=round(r["fps"], 4), unit=UNIT, cores=r["cores"], kind="port",
                                      sample=f"1 stream, dense flush ({r['first_frame_s']:.1f} s, untimed) then "
                                             f"{r['timed']} timed incremental frames of the same workload "
                                             f"(bounded to ~150 s of CPU time)"),
                    e2e=dict(value=round(r["fps"], 4), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the eventful_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist_ctx = None
    if world > 1:
        import torch.distributed as dist

        # NCCL writes its version banner to fd 1 when the communicator is created (whatever NCCL_DEBUG the launcher
        # exported): point fd 1 at stderr until the first collective has run, so stdout carries the JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device=torch.device("cuda", local))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        dist_ctx = dist
    from eventful_transformer import _native as native

    dev, dt = torch.device("cuda", local), torch.bfloat16
    pk = peaks()
    model = make_backbone(grid, "EventfulBlock", "EventfulTokenwiseBlock", args.k, dev, dt)
    total_streams = args.streams * world  # weak scaling: a fixed stream group per GPU
    my_streams = et_streams.partition_streams(total_streams, world, rank)
    frames_host = [f.pin_memory() for f in make_frames(my_streams, n, d)]
    frames_dev = [f.to(dev) for f in frames_host]
    use_graph = not args.no_graph

    # ---- headline: device-resident inputs
    sampler = ClockSampler(local) if rank == 0 else None
    ms, per_step, t_next = run_stream_model(model, frames_dev, args.warmup, args.steps, dist_ctx, use_graph)
    clocks = sampler.stop() if sampler is not None else None
    frames_total = args.steps * args.streams * world
    value = frames_total / (ms * 1e-3)

    # ---- e2e: pinned host input -> device, backbone, feature map -> pinned host, every step
    # every step uploads its frame from pinned host memory and downloads its feature map to pinned host memory; the
    # copies run on a second stream, double buffered, so they overlap the neighbouring frames' compute (et_pipeline)
    pipe = et_pipeline.FramePipeline(model, (args.streams, n, d), dt, dev)

    e2e_warm = max(3, args.warmup)  # untimed steps through the same pipeline: staging buffers, copy stream, pinned pages

    def e2e_step(i):
        pipe.step(frames_host[(t_next + i) % RING], frames_host[(t_next + i + 1) % RING])

    def e2e_finalize():  # the timed region ends when the LAST feature map has reached host memory
        torch.cuda.current_stream().wait_stream(pipe.copy_stream)

    with torch.inference_mode():
        for i in range(e2e_warm):
            e2e_step(i)
        t_next += e2e_warm
        ms_e2e = timed_steps(e2e_step, args.steps, dist_ctx, finalize=e2e_finalize)
        out_host = pipe.flush()
    e2e_value = frames_total / (ms_e2e * 1e-3)
    io_bytes = args.streams * n * d * 2

    # ---- NCCL only collects outputs (outside the timed region)
    if dist_ctx is not None:
        gathered = et_streams.gather_stream_outputs(out_host.to(dev), my_streams, total_streams, dist_ctx)
        assert gathered.shape[0] == total_streams and bool(torch.isfinite(gathered.float()).all())

    line = dict(metric=METRIC, value=round(value, 2), unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=max(3, args.warmup), ms_per_step=round(ms / args.steps, 4), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=workload, streams_per_gpu=args.streams, cuda_graph=use_graph,
                            l2="no flush: per-frame working set (A-gate columns 0.8 GB + token state 0.6 GB per "
                               "stream) exceeds the 126 MB L2; consecutive frames use different inputs",
                            peaks=pk["source"]),
                e2e=dict(value=round(e2e_value, 2), unit=UNIT, h2d_bytes_per_step=io_bytes, d2h_bytes_per_step=io_bytes,
                         ms_per_step=round(ms_e2e / args.steps, 4)),
                gpu_launches=int(per_step * args.steps), launches_per_step=int(per_step), clocks=clocks)

    # ---- configs[4]: a fixed total of streams partitioned over the GPUs (every rank takes part; strong scaling)
    bytes_per_stream = torch.cuda.max_memory_allocated(dev) / max(1, args.streams)
    del pipe, model
    torch.cuda.empty_cache()
    if not args.quick and args.total_streams > 0:
        torch.cuda.reset_peak_memory_stats(dev)
        line["partitioned_streams"] = partitioned_streams_leg(grid, args.total_streams, world, rank, dev, dt, args.k,
                                                              max(3, args.steps // 4), dist_ctx, bytes_per_stream)

    if rank == 0 and args.streams != 1:
        # the latency case: ONE stream on this GPU (same model, own state), device-resident inputs, CUDA-graph replay
        one = make_backbone(grid, "EventfulBlock", "EventfulTokenwiseBlock", args.k, dev, dt)
        one_frames = [f[:1].contiguous() for f in frames_dev]
        k1 = max(10, args.steps)
        ms1, per1, _ = run_stream_model(one, one_frames, args.warmup, k1, None, use_graph)
        line["single_stream"] = dict(value=round(k1 / (ms1 * 1e-3), 2), unit=UNIT, ms_per_frame=round(ms1 / k1, 4),
                                     launches_per_step=int(per1), steps=k1)
        del one, one_frames
        torch.cuda.empty_cache()

    if rank == 0 and not args.quick:
        with torch.inference_mode():
            # dense comparators on the same GPU, same weights, same frames (single stream group)
            dense = make_backbone(grid, "Block", "Block", None, dev, dt)
            ms_d, _, _ = run_stream_model(dense, frames_dev, 3, max(5, args.steps // 3), None, False)
            dense_fps = max(5, args.steps // 3) * args.streams / (ms_d * 1e-3)
            w = eager_dense_weights(dense, dev, dt)
            for _ in range(2):
                eager_dense_forward(frames_dev[0], w, grid)
            reps = 5
            ms_eager = timed_steps(lambda i: eager_dense_forward(frames_dev[i % RING], w, grid), reps, None)
            eager_fps = reps * args.streams / (ms_eager * 1e-3)
            del dense, w
            torch.cuda.empty_cache()
        line["dense"] = dict(fused_block_fps=round(dense_fps, 2), eager_torch_fps=round(eager_fps, 2),
                             speedup_vs_fused_dense=round(value / world / dense_fps, 3),
                             speedup_vs_eager_dense=round(value / world / eager_fps, 3),
                             note="fused = this repo's dense Block kernels; eager = the reference's dense Block "
                                  "restated with torch CUDA ops (materialised attention), both bf16 on this GPU")
        with torch.inference_mode():
            kernels = kernel_leg(args.streams, n, d, args.k, grid, pk)
        singles = {name: v for name, v in kernels.items() if "whole op" not in name}
        top = max(singles, key=lambda name: singles[name]["ms"] * singles[name]["per_step"])
        kt = kernels[top]
        line["roofline"] = dict(kernel=top, bound=kt["bound"], achieved=kt["achieved"], peak=kt["peak"], unit=kt["unit"],
                                frac=kt["frac"], traffic=kt.get("traffic"),
                                peak_source=pk["source"] + ", burst figure (kernel timed alone, L2 flushed, CUDA events "
                                            "on the launch stream)",
                                traffic_source="ncu --set full, profiles/r2_ncu_traffic.json <- r2_ncu_tc_apply*.csv (dram__bytes_read + write, "
                                               "one launch at this stream count)",
                                share_of_step=round(kt["ms"] * kt["per_step"] / (ms / args.steps), 3))
        line["kernels"] = kernels
        if world == 1:
            # BASELINE configs[4] k-sweep and configs[0], [2], [3] on this GPU
            with torch.inference_mode():
                line["k_sweep"] = dict(streams_per_gpu=args.streams, fused_dense_fps=round(dense_fps, 2),
                                       by_k=k_sweep(grid, args.streams, frames_dev, dev, dt, max(5, args.steps // 3), dense_fps))
            line["configs"] = other_configs(dev, max(10, args.steps))
            # the reference algorithm on the host cores: at the reference's own thread setting (configs/time/vitdet_vid/_cpu.yml:
            # threads 8) and on all cores
            cores = os.cpu_count() or 1
            r8 = cpu_reference(grid, args.k, 2, 0, budget_s=25.0, threads=min(8, cores))
            r = cpu_reference(grid, args.k, 2, 0, budget_s=25.0, threads=cores)
            sample = ("1 stream: dense flush then {t} incremental frame(s) of the same ViTDet-B 1024^2 k={k} workload, fp32 with "
                      "bf16 attention-value path (configs/time/vitdet_vid/_cpu.yml)")
            line["cpu_baseline"] = dict(value=round(r["fps"], 4), unit=UNIT, cores=r["cores"], kind="port",
                                        sample=sample.format(t=r["timed"], k=args.k),
                                        at_reference_threads=dict(value=round(r8["fps"], 4), cores=r8["cores"],
                                                                  sample=sample.format(t=r8["timed"], k=args.k)))
    if rank == 0:
        print(json.dumps(line))
    if dist_ctx is not None:
        dist_ctx.destroy_process_group()
    del native


if __name__ == "__main__":
    main()
