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
