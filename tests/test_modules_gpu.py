"""GPU parity of the stand-alone product / counted operators (MatmulBuffer, MatmulDeltaAccumulator, CountedMatmul,
CountedLinear.forward_bias / forward_linear) vs the oracle's restatement of modules.py / counting.py."""
import pytest
import torch
import torch.nn.functional as F

import eventful_oracle as orc
from eventful_transformer import counting, modules
from gpu_util import DEV, record, rel_err

pytestmark = pytest.mark.gpu


def _rand(shape, g, dtype):
    return torch.randn(shape, generator=g).to(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_matmul_buffer_rows_then_columns(dtype):
    """modules.py:204-252: rows index_q, then columns index_k of the stored product are refreshed; returns the state."""
    g = torch.Generator().manual_seed(41)
    b, h, n, dh, k = 2, 3, 64, 32, 12  # rows of the (n x n) product are 16-byte multiples in every dtype
    buf, st = modules.MatmulBuffer(), {}
    buf.counting()
    q = _rand((b, h, n, dh), g, dtype)
    kt = _rand((b, h, dh, n), g, dtype)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    for t in range(3):
        iq = ik = None
        if t:
            iq = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(b)])
            ik = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(b)])
            q = q.clone()
            q.scatter_(2, iq.view(b, 1, k, 1).expand(b, h, k, dh), _rand((b, h, k, dh), g, dtype))
            kt = kt.clone()
            kt.scatter_(3, ik.view(b, 1, 1, k).expand(b, h, dh, k), _rand((b, h, dh, k), g, dtype))
        want = orc.matmul_buffer(st, q.float(), kt.float(), iq, ik)
        got = buf(q.to(DEV), kt.to(DEV), None if iq is None else iq.to(DEV), None if ik is None else ik.to(DEV))
        assert got.data_ptr() == buf.product.data_ptr()  # the state itself is returned (modules.py:248)
        err = rel_err(got.cpu(), want)
        record("matmul_buffer", dtype=str(dtype), frame=t, rel_err=err)
        assert err < tol, (t, err)
    # counters: first frame n*n*dh per (b, h); then two refreshes of k*n*dh each
    assert buf.total_counts()["matmul_flops"] == b * h * (n * n * dh + 2 * 2 * k * n * dh)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_matmul_delta_accumulator_standalone(dtype):
    """modules.py:285-295: product += a_n . dV; product += dA . (v_n - dV)."""
    g = torch.Generator().manual_seed(43)
    b, h, n, dh, k = 2, 2, 40, 32, 9
    acc, st = modules.MatmulDeltaAccumulator(), {}
    a0, v0 = _rand((b, h, n, n), g, dtype), _rand((b, h, n, dh), g, dtype)
    want = orc.delta_accumulator(st, a0.float(), v0.float(), None, None)
    got = acc(a0.to(DEV), v0.to(DEV), None, None)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_err(got.cpu(), want) < tol
    for t in range(3):
        a_n, a_d = _rand((b, h, n, k), g, dtype), 0.1 * _rand((b, h, n, k), g, dtype)
        v_n, v_d = _rand((b, h, k, dh), g, dtype), 0.1 * _rand((b, h, k, dh), g, dtype)
        want = orc.delta_accumulator(st, a_n.float(), v_n.float(), a_d.float(), v_d.float())
        got = acc(a_n.to(DEV), v_n.to(DEV), a_d.to(DEV), v_d.to(DEV))
        assert got.data_ptr() == acc.product.data_ptr()
        err = rel_err(got.cpu(), want)
        record("matmul_delta_accumulator", dtype=str(dtype), frame=t, rel_err=err)
        assert err < tol, (t, err)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_counted_matmul_strided_operands(dtype):
    """CountedMatmul (counting.py:165-175) on permuted views, the way blocks.py feeds it (q, k^T views of the QKV buffer)."""
    g = torch.Generator().manual_seed(47)
    b, n, h, dh = 2, 70, 3, 32
    qkv = _rand((b, n, 3, h, dh), g, dtype)
    q, k = qkv[:, :, 0].permute(0, 2, 1, 3), qkv[:, :, 1].permute(0, 2, 3, 1)  # (b, h, n, dh), (b, h, dh, n)
    mm = counting.CountedMatmul()
    mm.counting()
    dev = qkv.to(DEV)
    got = mm(dev[:, :, 0].permute(0, 2, 1, 3), dev[:, :, 1].permute(0, 2, 3, 1))
    want = q.float() @ k.float()
    assert rel_err(got.cpu(), want) < (1e-5 if dtype == torch.float32 else 2e-2)
    assert mm.counts["matmul_flops"] == b * h * n * n * dh


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_counted_linear_parts(dtype):
    """forward == forward_linear + bias; forward_bias maps a zero token to the bias (counting.py:127-162)."""
    g = torch.Generator().manual_seed(53)
    lin = counting.CountedLinear(128, 192)
    lin.weight.data = 0.1 * torch.randn(192, 128, generator=g)
    lin.bias.data = torch.randn(192, generator=g)
    lin = lin.to(DEV).to(dtype).eval()
    lin.counting()
    x = _rand((3, 37, 128), g, dtype).to(DEV)
    w, bvec = lin.weight.detach().float().cpu(), lin.bias.detach().float().cpu()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    with torch.inference_mode():
        full = lin(x)
        prod = lin.forward_linear(x)
        pad = lin.forward_bias(torch.zeros(1, 1, 192, dtype=dtype, device=DEV))
    assert rel_err(full.cpu(), F.linear(x.float().cpu(), w, bvec)) < tol
    assert rel_err(prod.cpu(), F.linear(x.float().cpu(), w)) < tol
    assert torch.equal(pad.flatten().float().cpu(), bvec.to(dtype).float())
    rows = 3 * 37
    assert lin.counts["linear_flops"] == 2 * rows * 128 * 192
    assert lin.counts["bias_flops"] == rows * 192 + 192
