"""GPU parity of the attention kernels vs the oracle's restatement of blocks.py / modules.py (fp32 math on
the same bf16-rounded inputs)."""
import pytest
import torch

import eventful_oracle as orc
from eventful_transformer import _native as native
from eventful_transformer import blocks
from gpu_util import DEV, one_block_oracle, record, rel_err

pytestmark = pytest.mark.gpu
DT = torch.bfloat16


def block_params(dim, heads, rel, seed, std=0.5):
    g = torch.Generator().manual_seed(seed)
    p = {"blocks.0.qkv.bias": (0.5 * torch.randn(3 * dim, generator=g)).to(DT).float()}
    if rel is not None:
        p["blocks.0.relative_position.y_embedding"] = (std * torch.randn(2 * rel[0] - 1, dim // heads, generator=g)).to(DT).float()
        p["blocks.0.relative_position.x_embedding"] = (std * torch.randn(2 * rel[1] - 1, dim // heads, generator=g)).to(DT).float()
    return p


def gpu_block(cls, dim, heads, input_size, params, rel=None, window=None):
    kw = dict(dim=dim, heads=heads, input_size=input_size, mlp_ratio=4, window_size=window)
    if rel is not None:
        kw["relative_embedding_size"] = rel
    blk = getattr(blocks, cls)(**kw)
    sd = {k.replace("blocks.0.", ""): v for k, v in params.items()}
    blk.load_state_dict(sd, strict=False)
    return blk.to(DEV).to(DT).eval()


WINDOW_CASES = [  # dim, heads, grid, window, rel size, batch
    (768, 12, (16, 16), (14, 14), (64, 64), 2),   # ViTDet-B windowed block with padding 16 -> 28
    (768, 12, (28, 28), (14, 14), (64, 64), 1),   # no padding
    (32, 2, (7, 7), (4, 4), (5, 5), 3),           # dh = 16, padding 7 -> 8
    (64, 2, (9, 5), (3, 5), (4, 4), 2),           # dh = 32, rectangular
    (128, 2, (11, 7), (5, 3), (5, 3), 2),         # dh = 64, odd window height (uneven halves), padding 11 -> 15, 7 -> 9
    (128, 2, (16, 16), (16, 16), (16, 16), 1),    # dh = 64, one 256-token window: beyond the tcgen05 window kernels (208 keys)
    (768, 12, (64, 64), (14, 14), (64, 64), 1),   # the benchmarked shape: 1024^2 -> 64 x 64 tokens, 25 windows padded 64 -> 70
]


@pytest.mark.parametrize("dim,heads,grid,window,rel,batch", WINDOW_CASES)
def test_window_attention(dim, heads, grid, window, rel, batch):
    params = block_params(dim, heads, window, seed=dim + grid[0])  # windowed blocks size their tables to the window
    n = grid[0] * grid[1]
    qkv = torch.randn(batch, n, 3 * dim, generator=torch.Generator().manual_seed(1)).to(DT)
    oracle = one_block_oracle(params, dim, heads, grid, orc.TOKENWISE, window=window, rel=rel)
    want = oracle._attention_dense(0, qkv.float())[0]
    blk = gpu_block("EventfulTokenwiseBlock", dim, heads, grid, params, rel=rel, window=window)
    got = blk._dense_attention(qkv.to(DEV))
    assert got.shape == want.shape
    assert rel_err(got.cpu(), want) < 2e-2


@pytest.mark.parametrize("dim,heads,grid,rel,cls_token", [(768, 12, (14, 14), None, True), (32, 2, (7, 7), (5, 5), False),
                                                         (768, 12, (20, 20), None, True), (64, 4, (12, 12), (12, 12), False)])
def test_small_dense_attention(dim, heads, grid, rel, cls_token):
    params = block_params(dim, heads, rel, seed=5)
    n = grid[0] * grid[1] + int(cls_token)
    qkv = torch.randn(2, n, 3 * dim, generator=torch.Generator().manual_seed(2)).to(DT)
    oracle = one_block_oracle(params, dim, heads, grid, orc.DENSE, rel=rel, has_class_token=cls_token)
    want = oracle._attention_dense(0, qkv.float())[0]
    blk = gpu_block("Block", dim, heads, grid, params, rel=rel)
    assert rel_err(blk._dense_attention(qkv.to(DEV)).cpu(), want) < 2e-2


GLOBAL_CASES = [  # dim, heads, grid, rel, extra tokens, k, batch
    (768, 12, (4, 64), (4, 64), 0, 100, 2),       # tcgen05 path (64-wide grid: bias in registers), ragged k, batch 2
    (768, 12, (8, 64), (8, 64), 0, 512, 1),       # tcgen05 path, k == N
    (128, 2, (2, 64), None, 0, 37, 3),            # tcgen05 path without rel-pos
    (768, 12, (32, 32), (64, 64), 0, 300, 1),     # interpolated rel tables, ragged k
    (768, 12, (14, 14), None, 1, 64, 3),          # ViViT shape: 197 tokens incl. class token, batch of views
    (32, 2, (7, 7), (5, 5), 0, 12, 2),            # dh = 16, N not a multiple of 8
    (128, 2, (24, 24), (24, 24), 0, 576, 1),      # k == N (refresh everything)
    (768, 12, (64, 64), (64, 64), 0, 2048, 1),    # the benchmarked shape: tc_stats / tc_apply with rel-pos at N = 4096, k = 2048
    (768, 12, (42, 42), (64, 64), 0, 512, 1),     # BASELINE configs[0]: 672^2 -> 42 x 42 tokens (N = 1764 = 13.8 query blocks,
                                                  # 42-wide grid: bias of the statistics pass from one-hot MMAs), k = 512
    (768, 12, (20, 20), None, 1, 160, 2),         # ViViT EPIC-Kitchens shape: 401 tokens incl. class token
    (768, 12, (5, 64), (5, 64), 0, 130, 2),       # 64-wide grid with N % 128 == 64 (register-bias statistics pass, ragged tile)
]


@pytest.mark.parametrize("dim,heads,grid,rel,extra,k,batch", GLOBAL_CASES)
def test_global_eventful_attention_sequence(dim, heads, grid, rel, extra, k, batch):
    """FIRST then DELTA frames vs the oracle's matmul_buffer + gates + delta_accumulator; also checks the state."""
    params = block_params(dim, heads, rel, seed=9)
    n = grid[0] * grid[1] + extra
    g = torch.Generator().manual_seed(3)
    oracle = one_block_oracle(params, dim, heads, grid, orc.EVENTFUL, rel=rel, has_class_token=bool(extra))
    blk = gpu_block("EventfulBlock", dim, heads, grid, params, rel=rel)
    qkv = torch.randn(batch, n, 3 * dim, generator=g).to(DT)
    buf_cpu, buf_gpu = qkv.float().clone(), qkv.to(DEV).clone()
    want = oracle._attention_eventful(0, buf_cpu, None)[0]
    got = blk._attention_first(buf_gpu, None)
    assert rel_err(got.cpu(), want) < 2e-2
    for t in range(1, 4):
        idx = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(batch)])
        rows = (qkv.float().gather(1, idx.unsqueeze(-1).expand(-1, -1, 3 * dim))
                + 0.5 * torch.randn(batch, k, 3 * dim, generator=g)).to(DT)
        buf_cpu.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, 3 * dim), rows.float())
        buf_gpu.scatter_(1, idx.to(DEV).unsqueeze(-1).expand(-1, -1, 3 * dim), rows.to(DEV))
        want = oracle._attention_eventful(0, buf_cpu, idx)[0]
        got = blk._attention_incremental(buf_gpu, idx.to(DEV))
        assert rel_err(got.cpu(), want) < 3e-2, t
        # state parity: A-gate reference (logical (B, H, N, N)), v-gate reference, accumulator
        st = oracle.state[0]
        assert rel_err(blk.matmul_gate.p.cpu(), st["matmul_gate"]["p"]) < 4e-2  # bias tables are rounded to bf16 (as the bf16 reference does)
        assert torch.equal(blk.v_gate.p.cpu().float(), st["v_gate"]["p"])
        assert rel_err(blk.matmul_accumulator_2.product.cpu(), st["matmul_accumulator_2"]["product"]) < 3e-2


def test_global_dense_attention_large():
    dim, heads, grid, rel = 768, 12, (32, 32), (64, 64)
    params = block_params(dim, heads, rel, seed=11)
    qkv = torch.randn(1, 1024, 3 * dim, generator=torch.Generator().manual_seed(4)).to(DT)
    oracle = one_block_oracle(params, dim, heads, grid, orc.DENSE, rel=rel)
    want = oracle._attention_dense(0, qkv.float())[0]
    blk = gpu_block("EventfulMatmul1Block", dim, heads, grid, params, rel=rel)
    got = blk._attention_incremental(qkv.to(DEV), None)
    assert rel_err(got.cpu(), want) < 2e-2


@pytest.mark.parametrize("tc", [0, 1])
def test_delta_with_static_input_is_stationary(tc):
    """
    Identical frames: a_n == p, dV == 0, so the accumulator must not move.  mma.sync kernels (tc = 0): a_n is rounded to the
    state dtype before it is used, dA is exactly zero and the output is bit-for-bit the first frame's.  tcgen05 kernels
    (tc = 1): acc += a_n . v_n - p . (v_n - dV) runs as two fp32 MMA sums whose difference is zero only up to fp32 rounding
    (~1e-7 of the products), so after the bf16 rounding of acc an element may move by one ulp when it sat on a rounding
    boundary: at most 0.1 % of the elements, each by at most one bf16 ulp.
    """
    dim, heads, grid = 768, 12, (16, 16)
    params = block_params(dim, heads, (64, 64), seed=13)
    try:
        native.lib().et_debug_set(2, tc)
        blk = gpu_block("EventfulBlock", dim, heads, grid, params, rel=(64, 64))
        qkv = torch.randn(1, 256, 3 * dim, generator=torch.Generator().manual_seed(5)).to(DT).to(DEV)
        first = blk._attention_first(qkv, None).clone()
        idx = torch.randperm(256, generator=torch.Generator().manual_seed(6))[:100].view(1, -1).to(DEV)
        again = blk._attention_incremental(qkv, idx)
    finally:
        native.lib().et_debug_set(2, 1)
    if tc == 0:
        assert torch.equal(first, again)
    else:
        diff = (first.float() - again.float()).abs()
        moved = float((diff > 0).float().mean())
        assert moved <= 1e-3 and bool((diff <= 2.0 ** -7 * first.float().abs() + 1e-6).all()), (moved, float(diff.max()))


@pytest.mark.parametrize("rel", [(16, 64), None])
def test_tensor_core_path_agrees_with_mma_sync_path(rel):
    """The tcgen05 kernels and the mma.sync kernels implement the same function: run both on identical inputs."""
    dim, heads, grid, k = 768, 12, (16, 64), 333
    n = grid[0] * grid[1]
    params = block_params(dim, heads, rel, seed=17, std=0.2)
    g = torch.Generator().manual_seed(8)
    qkv0 = torch.randn(2, n, 3 * dim, generator=g).to(DT).to(DEV)
    idx = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(2)]).to(DEV)
    qkv1 = qkv0.clone()
    qkv1.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, 3 * dim), torch.randn(2, k, 3 * dim, generator=g).to(DT).to(DEV))
    results = []
    try:
        for flag in (1, 0):
            native.lib().et_debug_set(2, flag)
            blk = gpu_block("EventfulBlock", dim, heads, grid, params, rel=rel)
            first = blk._attention_first(qkv0, None).clone()
            second = blk._attention_incremental(qkv1, idx).clone()
            results.append((first, second, blk.matmul_gate.p.clone(), blk._acc.clone()))
    finally:
        native.lib().et_debug_set(2, 1)
    for a, b in zip(*results):
        assert rel_err(a, b) < 1.5e-2


def test_tensor_core_dense_global_attention():
    dim, heads, grid, rel = 768, 12, (8, 64), (8, 64)
    params = block_params(dim, heads, rel, seed=19)
    qkv = torch.randn(2, 512, 3 * dim, generator=torch.Generator().manual_seed(9)).to(DT)
    oracle = one_block_oracle(params, dim, heads, grid, orc.DENSE, rel=rel)
    want = oracle._attention_dense(0, qkv.float())[0]
    blk = gpu_block("EventfulMatmul1Block", dim, heads, grid, params, rel=rel)
    assert rel_err(blk._attention_incremental(qkv.to(DEV), None).cpu(), want) < 2e-2


@pytest.mark.parametrize("grid,batch", [((16, 16), 2), ((64, 64), 1), ((28, 42), 1)])
def test_tensor_core_window_paths_agree_with_mma_sync_path(grid, batch):
    """Three implementations of the same function: second- and first-generation tcgen05 window kernels, mma.sync kernel."""
    dim, heads, window = 768, 12, (14, 14)
    params = block_params(dim, heads, window, seed=23, std=0.3)
    qkv = torch.randn(batch, grid[0] * grid[1], 3 * dim, generator=torch.Generator().manual_seed(10)).to(DT).to(DEV)
    outs = []
    try:
        for tc, gen in ((1, 2), (1, 1), (0, 2)):
            native.lib().et_debug_set(2, tc)
            native.lib().et_debug_set(11, gen)
            blk = gpu_block("EventfulTokenwiseBlock", dim, heads, grid, params, rel=(64, 64), window=window)
            outs.append(blk._dense_attention(qkv).clone())
    finally:
        native.lib().et_debug_set(2, 1)
        native.lib().et_debug_set(11, 2)
    assert rel_err(outs[0], outs[2]) < 1.5e-2
    assert rel_err(outs[1], outs[2]) < 1.5e-2


def test_delta_accumulator_does_not_drift_over_32_frames():
    """
    The tcgen05 apply kernel accumulates  acc += a_n . v_n - p . (v_n - dV)  in fp32 and rounds acc to bf16 once per
    frame, where the reference adds the two bf16-rounded products a_n . dV and dA . (v_n - dV) (modules.py:293-294).
    Algebraically equal; this test follows both for 32 incremental frames on the same inputs and index sets and
    requires that the distance to the exact-arithmetic oracle does not grow.
    """
    dim, heads, grid, k, frames = 768, 12, (8, 64), 160, 32
    rel = (8, 64)
    n = grid[0] * grid[1]
    params = block_params(dim, heads, rel, seed=29, std=0.2)
    g = torch.Generator().manual_seed(12)
    oracle = one_block_oracle(params, dim, heads, grid, orc.EVENTFUL, rel=rel)
    blk = gpu_block("EventfulBlock", dim, heads, grid, params, rel=rel)
    qkv = torch.randn(1, n, 3 * dim, generator=g).to(DT)
    buf_cpu, buf_gpu = qkv.float().clone(), qkv.to(DEV).clone()
    want = oracle._attention_eventful(0, buf_cpu, None)[0]
    got = blk._attention_first(buf_gpu, None)
    errs = [rel_err(got.cpu(), want)]
    for t in range(1, frames + 1):
        idx = torch.randperm(n, generator=g)[:k].view(1, k)
        rows = (buf_cpu.gather(1, idx.unsqueeze(-1).expand(-1, -1, 3 * dim))
                + 0.3 * torch.randn(1, k, 3 * dim, generator=g)).to(DT)
        buf_cpu.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, 3 * dim), rows.float())
        buf_gpu.scatter_(1, idx.to(DEV).unsqueeze(-1).expand(-1, -1, 3 * dim), rows.to(DEV))
        want = oracle._attention_eventful(0, buf_cpu, idx)[0]
        got = blk._attention_incremental(buf_gpu, idx.to(DEV))
        errs.append(rel_err(got.cpu(), want))
    record("delta_accumulator_drift", errors=[round(e, 5) for e in errs])
    assert max(errs) < 3e-2, errs
    # acc is stored in bf16 (as in the reference): rounding once per frame is a random walk ~ sqrt(frames) x 2^-9, nothing faster
    assert max(errs[-8:]) <= 2.5 * max(errs[1:9]) + 3e-3, errs
