"""GPU parity of et_linear (tcgen05 GEMM + bias/GELU + TokenBuffer scatter epilogue) vs fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

from eventful_transformer import _native as native

pytestmark = pytest.mark.gpu
DEV = "cuda"

SHAPES = [  # (M, K, n_feat)
    (2048, 768, 2304), (2048, 768, 768), (2048, 768, 3072), (2048, 3072, 768),  # ViTDet-B, k = 2048
    (512, 768, 2304), (768, 768, 3072), (4096, 768, 768),
    (100, 32, 96), (257, 72, 136), (1, 768, 768), (129, 64, 8), (300, 40, 264),  # ragged M / K / n tails
]


def reference(x, w, b, act):
    y = x.float() @ w.float().t() + (0 if b is None else b.float())
    return F.gelu(y) if act else y


def make(m, k, f, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(m, k, generator=g).to(dtype).to(DEV)
    w = (torch.randn(f, k, generator=g) / k ** 0.5).to(dtype).to(DEV)
    b = torch.randn(f, generator=g).to(dtype).to(DEV)
    return x, w, b


def check(y, ref, dtype):
    tol = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    err = (y.float() - ref).abs()
    bound = tol * ref.abs() + tol * float(ref.abs().mean())
    assert bool((err <= bound).all()), f"max err {float(err.max())} (ref scale {float(ref.abs().mean())})"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,k,f", SHAPES)
@pytest.mark.parametrize("act", [0, 1])
def test_linear_dense(m, k, f, act, dtype):
    x, w, b = make(m, k, f, dtype)
    y = native.linear(x, w, b, act=act)
    assert y.shape == (m, f) and y.dtype == dtype
    check(y, reference(x, w, b, act), dtype)


@pytest.mark.parametrize("block_n", [64, 96, 128, 192, 256])
def test_linear_every_tile_width(block_n):
    try:
        native.lib().et_debug_set(1, block_n)
        for m, k, f in [(2048, 768, 2304), (257, 72, 136), (640, 3072, 768)]:
            x, w, b = make(m, k, f, torch.bfloat16, seed=block_n)
            check(native.linear(x, w, b), reference(x, w, b, 0), torch.bfloat16)
            check(native.linear(x, w, None, act=1), reference(x, w, None, 1), torch.bfloat16)
    finally:
        native.lib().et_debug_set(1, 0)


@pytest.mark.parametrize("block_n", [128, 192, 256])
def test_linear_persistent_kernel(block_n):
    """The persistent multi-wave kernel (ring across tiles, double-buffered TMEM accumulator), forced for every shape:
    fewer tiles than SMs, many tiles per CTA, ragged M / N, GELU, scatter epilogue; repeated launches."""
    dtype = torch.bfloat16
    try:
        native.lib().et_debug_set(1, block_n)
        native.lib().et_debug_set(9, 1)
        for m, k, f in [(2048, 768, 2304), (257, 72, 136), (100, 64, 64), (640, 3072, 768), (16384, 768, 3072),
                        (16384, 3072, 768), (5000, 768, 2304)]:
            x, w, b = make(m, k, f, dtype, seed=block_n + m)
            for _ in range(2):
                check(native.linear(x, w, b), reference(x, w, b, 0), dtype)
            check(native.linear(x, w, None, act=1), reference(x, w, None, 1), dtype)
        batch, n, k_sel, kdim, f = 8, 4096, 2048, 768, 768
        x, w, b = make(batch * k_sel, kdim, f, dtype, seed=9)
        idx = torch.stack([torch.randperm(n, generator=torch.Generator().manual_seed(5 + i))[:k_sel] for i in range(batch)]).to(DEV)
        buf = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        native.linear(x.view(batch, k_sel, kdim), w, b, out=buf, idx=idx)
        native.lib().et_debug_set(9, 2)
        native.lib().et_debug_set(8, 1)
        dense = native.linear(x, w, b).view(batch, k_sel, f)
        want = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        want.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, f), dense)
        assert torch.equal(buf, want)
    finally:
        native.lib().et_debug_set(1, 0)
        native.lib().et_debug_set(8, 0)
        native.lib().et_debug_set(9, 0)


@pytest.mark.parametrize("block_n", [128, 192, 256])
def test_linear_256_row_tiles(block_n):
    """Two 128-row accumulators per CTA sharing one W tile (the multi-stream configuration), forced for small shapes:
    ragged M (tail rows in the second half, or no second half at all), GELU, and the scatter epilogue."""
    dtype = torch.bfloat16
    try:
        native.lib().et_debug_set(1, block_n)
        native.lib().et_debug_set(8, 2)
        for m, k, f in [(2048, 768, 2304), (257, 72, 136), (100, 64, 64), (640, 3072, 768), (16384, 768, 3072)]:
            x, w, b = make(m, k, f, dtype, seed=block_n + m)
            check(native.linear(x, w, b), reference(x, w, b, 0), dtype)
            check(native.linear(x, w, None, act=1), reference(x, w, None, 1), dtype)
        batch, n, k_sel, kdim, f = 3, 700, 300, 768, 768
        x, w, b = make(batch * k_sel, kdim, f, dtype, seed=9)
        idx = torch.stack([torch.randperm(n, generator=torch.Generator().manual_seed(5 + i))[:k_sel] for i in range(batch)]).to(DEV)
        buf = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        native.linear(x.view(batch, k_sel, kdim), w, b, out=buf, idx=idx)
        native.lib().et_debug_set(8, 1)
        dense = native.linear(x, w, b).view(batch, k_sel, f)
        want = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        want.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, f), dense)
        assert torch.equal(buf, want)
    finally:
        native.lib().et_debug_set(1, 0)
        native.lib().et_debug_set(8, 0)


@pytest.mark.parametrize("batch,n,k_sel,kdim,f", [(1, 4096, 2048, 768, 2304), (3, 197, 64, 768, 768), (2, 50, 7, 32, 96)])
def test_linear_scatter_epilogue(batch, n, k_sel, kdim, f):
    """out[b, idx[b, j]] = y[b, j]; untouched rows keep their previous content (TokenBuffer semantics)."""
    dtype = torch.bfloat16
    x, w, b = make(batch * k_sel, kdim, f, dtype, seed=3)
    g = torch.Generator().manual_seed(4)
    idx = torch.stack([torch.randperm(n, generator=g)[:k_sel] for _ in range(batch)]).to(DEV)
    buf = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
    out = native.linear(x.view(batch, k_sel, kdim), w, b, out=buf, idx=idx)
    assert out.data_ptr() == buf.data_ptr()
    dense = native.linear(x, w, b).view(batch, k_sel, f)
    want = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
    want.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, f), dense)
    assert torch.equal(buf, want)
    check(dense.view(-1, f), reference(x, w, b, 0), dtype)


def test_linear_rejects_mixed_dtypes_and_bad_shapes():
    x = torch.zeros(8, 64, device=DEV)
    with pytest.raises(TypeError, match="cast the model or the input"):
        native.linear(x, torch.zeros(16, 64, device=DEV, dtype=torch.bfloat16), None)
    with pytest.raises(native.NativeError, match="multiples of 8"):
        native.linear(torch.zeros(8, 12, device=DEV, dtype=torch.bfloat16),
                      torch.zeros(16, 12, device=DEV, dtype=torch.bfloat16), None)


@pytest.mark.parametrize("m,k,f,act,scatter", [(300, 768, 2304, 0, True), (2048, 768, 3072, 1, True), (512, 3072, 768, 0, True),
                                               (197, 768, 768, 0, False), (37, 32, 96, 1, True), (1, 128, 4, 0, False)])
def test_fp32_linear_matches_torch(m, k, f, act, scatter):
    """fp32 models (BASELINE configs[0]): CUDA-core SGEMM with the same bias / exact-erf GELU / TokenBuffer scatter epilogue."""
    g = torch.Generator().manual_seed(m + f)
    x = torch.randn(1, m, k, generator=g)
    w = torch.randn(f, k, generator=g) / k ** 0.5
    b = torch.randn(f, generator=g)
    want = F.linear(x, w, b)
    if act:
        want = F.gelu(want)
    if scatter:
        n = m + 50
        idx = torch.randperm(n, generator=g)[:m].view(1, m)
        buf = torch.full((1, n, f), 7.0)
        got = native.linear(x.to(DEV), w.to(DEV), b.to(DEV), act=act, out=buf.to(DEV), idx=idx.to(DEV)).cpu()
        ref = buf.clone()
        ref[0, idx[0]] = want[0]
        want = ref
        count = torch.tensor([m // 2], dtype=torch.int32)  # device-side count: only the first half of the rows is stored
        got_c = native.linear(x.to(DEV), w.to(DEV), b.to(DEV), act=act, out=buf.to(DEV), idx=idx.to(DEV), count=count.to(DEV)).cpu()
        ref_c = buf.clone()
        ref_c[0, idx[0, : m // 2]] = want[0, idx[0, : m // 2]]
        assert torch.allclose(got_c, ref_c, rtol=2e-5, atol=2e-5)
    else:
        got = native.linear(x.to(DEV), w.to(DEV), b.to(DEV), act=act).cpu()
    assert torch.allclose(got, want, rtol=2e-5, atol=2e-5), float((got - want).abs().max())


def test_device_side_count_skips_whole_tiles_on_the_tensor_core_path():
    """bf16 GEMM with a device-side row count: rows beyond it are neither computed into the buffer nor stored (tiles made only of
    such rows exit early); rows below it match the un-counted call."""
    g = torch.Generator().manual_seed(77)
    m, k, f, n = 1024, 768, 768, 1024
    x = torch.randn(1, m, k, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(f, k, generator=g) / k ** 0.5).to(torch.bfloat16).to(DEV)
    b = torch.randn(f, generator=g).to(torch.bfloat16).to(DEV)
    idx = torch.randperm(n, generator=g)[:m].view(1, m).to(DEV)
    full = native.linear(x, w, b, out=torch.zeros(1, n, f, dtype=torch.bfloat16, device=DEV), idx=idx)
    for valid in (0, 1, 130, 517, 1024):
        count = torch.tensor([valid], dtype=torch.int32, device=DEV)
        got = native.linear(x, w, b, out=torch.zeros(1, n, f, dtype=torch.bfloat16, device=DEV), idx=idx, count=count)
        want = torch.zeros_like(full)
        want[0, idx[0, :valid]] = full[0, idx[0, :valid]]
        assert torch.equal(got, want), valid


def _gather_case(batch, n, k_sel, kdim, f, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(batch, n, kdim, generator=g).to(dtype).to(DEV)
    w = (torch.randn(f, kdim, generator=g) / kdim ** 0.5).to(dtype).to(DEV)
    b = torch.randn(f, generator=g).to(dtype).to(DEV)
    idx = torch.stack([torch.randperm(n, generator=g)[:k_sel] for _ in range(batch)]).to(DEV)
    return src, w, b, idx


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("batch,n,k_sel,kdim,f", [
    (1, 4096, 2048, 768, 2304), (1, 4096, 2048, 768, 768), (1, 4096, 2048, 768, 3072),  # ViTDet-B gate sites, one stream
    (8, 4096, 2048, 768, 2304), (8, 4096, 2048, 768, 768),                              # 8 streams: persistent kernel
    (3, 197, 64, 768, 2304), (2, 300, 131, 72, 136), (1, 50, 1, 64, 8), (2, 64, 64, 40, 264),  # ragged tails, k = N
])
def test_linear_gather_equals_linear_on_gathered_rows(batch, n, k_sel, kdim, f, dtype):
    """et_linear_gather (TMA gather4 A operand + gate-state advance + scatter epilogue) is bit-identical to et_linear on the
    rows gathered beforehand (same tiles, same arithmetic), advances exactly the selected state rows and leaves the rest."""
    src, w, b, idx = _gather_case(batch, n, k_sel, kdim, f, dtype, seed=k_sel + f)
    rows = torch.gather(src, 1, idx[..., None].expand(-1, -1, kdim)).contiguous()
    for act in (0, 1):
        state = torch.full_like(src, 3.0)
        buf = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        native.linear_gather(src, idx, w, b, state=state, act=act, out=buf, idx=idx)
        want = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        native.linear(rows, w, b, act=act, out=want, idx=idx)
        assert torch.equal(buf, want)
        check(torch.gather(buf, 1, idx[..., None].expand(-1, -1, f)).reshape(-1, f),
              reference(rows.reshape(-1, kdim), w, b, act), dtype)
        want_state = torch.full_like(src, 3.0)
        want_state.scatter_(1, idx[..., None].expand(-1, -1, kdim), rows)
        assert torch.equal(state, want_state)
        # no scatter (mlp_1): dense (B, k, F) output, no state
        y = native.linear_gather(src, idx, w, b, act=act)
        assert torch.equal(y, native.linear(rows, w, b, act=act))


@pytest.mark.parametrize("block_n,persist,mh", [(64, 2, 1), (96, 2, 1), (128, 1, 1), (192, 1, 1), (256, 1, 1), (256, 2, 2), (192, 2, 2)])
def test_linear_gather_every_kernel_variant(block_n, persist, mh):
    dtype = torch.bfloat16
    lib = native.lib()
    try:
        lib.et_debug_set(1, block_n), lib.et_debug_set(9, persist), lib.et_debug_set(8, mh if persist == 2 else 0)
        for batch, n, k_sel, kdim, f in [(2, 1024, 600, 768, 768), (1, 300, 257, 72, 136), (4, 4096, 2048, 3072, 768)]:
            src, w, b, idx = _gather_case(batch, n, k_sel, kdim, f, dtype, seed=block_n + n)
            rows = torch.gather(src, 1, idx[..., None].expand(-1, -1, kdim)).contiguous()
            state = torch.zeros_like(src)
            buf = torch.zeros((batch, n, f), dtype=dtype, device=DEV)
            for _ in range(2):
                native.linear_gather(src, idx, w, b, state=state, out=buf, idx=idx)
            want = torch.zeros((batch, n, f), dtype=dtype, device=DEV)
            native.linear(rows, w, b, out=want, idx=idx)
            assert torch.equal(buf, want)
            assert torch.equal(state, torch.zeros_like(src).scatter_(1, idx[..., None].expand(-1, -1, kdim), rows))
    finally:
        lib.et_debug_set(1, 0), lib.et_debug_set(9, 0), lib.et_debug_set(8, 0)


def test_linear_gather_device_side_count():
    """Threshold policy: only the first count[b] entries of the padded index are gathered, advanced and scattered."""
    dtype, n, kdim, f = torch.bfloat16, 1024, 768, 2304
    src, w, b, idx = _gather_case(1, n, n, kdim, f, dtype, seed=3)
    for found in (0, 1, 127, 128, 700, 1024):
        count = torch.tensor([found], dtype=torch.int32, device=DEV)
        state = torch.full_like(src, 3.0)
        buf = torch.full((1, n, f), 7.0, dtype=dtype, device=DEV)
        native.linear_gather(src, idx, w, b, state=state, out=buf, idx=idx, count=count)
        sel = idx[:, :found]
        rows = torch.gather(src, 1, sel[..., None].expand(-1, -1, kdim)).contiguous()
        want = torch.full((1, n, f), 7.0, dtype=dtype, device=DEV)
        if found:
            native.linear(rows, w, b, out=want, idx=sel.contiguous())
        assert torch.equal(buf, want), found
        assert torch.equal(state, torch.full_like(src, 3.0).scatter_(1, sel[..., None].expand(-1, -1, kdim), rows)), found


def test_linear_gather_rejects_fp32_and_bad_shapes():
    src = torch.zeros(1, 16, 64, device=DEV)
    idx = torch.zeros(1, 4, dtype=torch.int64, device=DEV)
    with pytest.raises(TypeError):
        native.linear_gather(src, idx, torch.zeros(8, 64, device=DEV), None)
    src = src.bfloat16()
    with pytest.raises(ValueError):
        native.linear_gather(src, idx, torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16), None, state=torch.zeros(1, 8, 64, device=DEV, dtype=torch.bfloat16))


def test_linear_cta_pair_kernel():
    """The cta_group::2 kernel (two CTAs of a cluster per 256 x 256 tile, W split across the pair, M = 256 MMAs issued by the
    leader), forced for every shape: fewer tiles than CTA pairs, many tiles per pair, ragged M / N tails, GELU, long K,
    scatter epilogue, repeated launches (barrier phases / double-buffered tensor memory across tiles)."""
    dtype = torch.bfloat16
    lib = native.lib()
    try:
        lib.et_debug_set(13, 1)
        for m, k, f in [(2048, 768, 2304), (257, 72, 264), (300, 64, 256), (640, 3072, 768), (16384, 768, 3072),
                        (16384, 3072, 768), (5000, 768, 2304), (16384, 768, 768)]:
            x, w, b = make(m, k, f, dtype, seed=m + f)
            for _ in range(2):
                check(native.linear(x, w, b), reference(x, w, b, 0), dtype)
            check(native.linear(x, w, None, act=1), reference(x, w, None, 1), dtype)
        x, w, b = make(1000, 768, 768, torch.float16, seed=2)
        check(native.linear(x, w, b), reference(x, w, b, 0), torch.float16)
        batch, n, k_sel, kdim, f = 8, 4096, 2048, 768, 768
        x, w, b = make(batch * k_sel, kdim, f, dtype, seed=9)
        idx = torch.stack([torch.randperm(n, generator=torch.Generator().manual_seed(5 + i))[:k_sel] for i in range(batch)]).to(DEV)
        buf = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        native.linear(x.view(batch, k_sel, kdim), w, b, out=buf, idx=idx)
        lib.et_debug_set(13, 2)
        want = torch.full((batch, n, f), 7.0, dtype=dtype, device=DEV)
        native.linear(x.view(batch, k_sel, kdim), w, b, out=want, idx=idx)
        assert torch.equal(buf, want)  # same products in the same order: bit-identical to the single-CTA kernels
    finally:
        lib.et_debug_set(13, 0)
