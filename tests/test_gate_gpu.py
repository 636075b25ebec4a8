"""GPU parity of the gate kernels (et_gate_select / et_gate_gather / et_buffer_scatter / et_add) vs the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import eventful_oracle as orc
from eventful_transformer import _native as native
from eventful_transformer import modules, policies

pytestmark = pytest.mark.gpu
DEV = "cuda"
DTYPES = [torch.float32, torch.bfloat16, torch.float16]
SHAPES = [((1, 4096, 768), 2048), ((1, 1764, 768), 512), ((12, 197, 768), 64), ((2, 49, 32), 12),
          ((3, 10000, 64), 100), ((2, 6, 4096, 64), 1000), ((1, 300, 1024), 300), ((1, 64, 1280), 1)]


def dyadic(shape, seed, scale=4.0, lim=8):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-lim, lim + 1, shape, generator=g).float() / scale


def ulp_close(a, b, dtype):
    eps = {torch.float32: 2.0 ** -22, torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10}[dtype]
    return (a.float() - b.float()).abs() <= eps * b.float().abs() + 1e-30


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,k", SHAPES)
def test_topk_exact_on_dyadic_inputs(shape, k, dtype):
    """Sums of squares are exact in fp32 -> norms are bit-identical -> index ORDER must equal the oracle's."""
    c, p = dyadic(shape, 1).to(dtype), dyadic(shape, 2).to(dtype)
    keep = (torch.rand(shape[:-1], generator=torch.Generator().manual_seed(3)) < 0.5).unsqueeze(-1)
    p = torch.where(keep, c, p)  # half of the tokens unchanged -> massive ties at norm 0
    want = orc.select_topk(orc.token_norm(c - p), k)
    got, _ = native.gate_select(c.to(DEV), p=p.to(DEV), k=k)
    assert got.dtype == torch.int64 and got.shape == want.shape
    assert torch.equal(got.cpu(), want)
    # and against the live torch CUDA ops the reference would run (policies.py:63): same index SET
    ref = torch.linalg.vector_norm(c.to(DEV) - p.to(DEV), dim=-1).topk(k, sorted=False)[1]
    assert torch.equal(got.sort(dim=-1)[0], ref.sort(dim=-1)[0])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,k", [(4096, 2048), (1764, 512), (197, 64), (9000, 17)])
def test_tie_order_matches_torch_cuda_radix_select(n, k, dtype):
    """order = greater-than-kth ascending, then ties ascending. Checked against torch.topk on this GPU."""
    g = torch.Generator().manual_seed(n)
    norm = (torch.randint(0, 40, (2, n), generator=g).float() / 8).to(dtype)
    x = torch.zeros(2, n, 8 if dtype != torch.float32 else 4, dtype=dtype)
    x[..., 0] = norm
    got, _ = native.gate_select(x.to(DEV), k=k)
    ref = norm.to(DEV).topk(k, sorted=False)[1]
    assert torch.equal(got.sort(dim=-1)[0], ref.sort(dim=-1)[0]), "index sets differ from torch.topk (CUDA)"
    assert torch.equal(got.cpu(), orc.select_topk(norm, k))


@pytest.mark.parametrize("dtype", DTYPES)
def test_topk_generic_inputs_differ_only_at_one_ulp_ties(dtype):
    shape, k = (2, 4096, 768), 2048
    g = torch.Generator().manual_seed(5)
    c = torch.randn(shape, generator=g).to(dtype).to(DEV)
    p = (c.float() + 0.1 * torch.randn(shape, generator=g).to(DEV)).to(dtype)
    got, _ = native.gate_select(c, p=p, k=k)
    norm = torch.linalg.vector_norm(c - p, dim=-1)
    ref = norm.topk(k, sorted=False)[1]
    for r in range(shape[0]):
        a, b = set(got[r].tolist()), set(ref[r].tolist())
        assert len(a) == k
        kth = norm[r].float().topk(k)[0][-1]
        for t in a ^ b:
            assert bool(ulp_close(norm[r, t], kth, dtype)), (t, float(norm[r, t]), float(kth))
        assert len(a ^ b) <= 0.01 * k


@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_add_layernorm_select_and_gather(dtype):
    b, n, d, k = 2, 1024, 768, 300
    g = torch.Generator().manual_seed(7)
    xa = torch.randn(b, n, d, generator=g).to(dtype).to(DEV)
    xb = torch.randn(b, n, d, generator=g).to(dtype).to(DEV)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(dtype).to(DEV)
    bias = (0.1 * torch.randn(d, generator=g)).to(dtype).to(DEV)
    x = xa + xb
    c = F.layer_norm(x, (d,), w, bias, 1e-6)
    p = (c.float() + 0.3 * torch.randn(b, n, d, generator=g).to(DEV) * (torch.rand(b, n, 1, generator=g).to(DEV) < 0.5)).to(dtype)
    p0 = p.clone()
    c_all = torch.empty_like(xa)
    idx, xsum = native.gate_select(xa, p=p, xb=xb, want_sum=True, ln=(w, bias), eps=1e-6, k=k, c_out=c_all)
    assert torch.equal(xsum, x)
    idx_plain, _ = native.gate_select(xa, p=p, xb=xb, want_sum=True, ln=(w, bias), eps=1e-6, k=k)
    assert torch.equal(idx, idx_plain)  # emitting c does not change the selection
    norm = torch.linalg.vector_norm(c - p, dim=-1).float()
    for r in range(b):
        a, ref = set(idx[r].tolist()), set(norm[r].topk(k)[1].tolist())
        kth = norm[r].topk(k)[0][-1]
        assert len(a) == k and len(a ^ ref) <= 0.04 * k
        for t in a ^ ref:
            assert abs(float(norm[r, t]) - float(kth)) <= 0.02 * float(kth) + 1e-3
    c_t, e_t = native.gate_gather(xsum, idx, p=p, ln=(w, bias), eps=1e-6, want_delta=True)
    want_c = c.gather(1, idx.unsqueeze(-1).expand(-1, -1, d))
    tol = dict(rtol=2e-2, atol=2e-2) if dtype != torch.float32 else dict(rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(c_t, want_c, **tol)
    # the gate input emitted for every token (gather source of et_linear_gather) holds exactly the rows et_gate_gather makes
    assert torch.equal(c_all.gather(1, idx.unsqueeze(-1).expand(-1, -1, d)), c_t)
    torch.testing.assert_close(c_all, c, **tol)
    # state advanced exactly at the selected rows, untouched elsewhere; e~ = c~ - p_old exactly (in dtype)
    assert torch.equal(p.gather(1, idx.unsqueeze(-1).expand(-1, -1, d)), c_t)
    mask = torch.ones(b, n, dtype=torch.bool, device=DEV)
    mask.scatter_(1, idx, False)
    assert torch.equal(p[mask], p0[mask])
    assert torch.equal(e_t, c_t - p0.gather(1, idx.unsqueeze(-1).expand(-1, -1, d)))
    # gate_before_ln: state keeps raw rows, output is LN(raw rows)
    q = x.clone()
    out, _ = native.gate_gather(x, idx, p=q, ln=(w, bias), eps=1e-6, ln_after=True)
    torch.testing.assert_close(out, want_c, **tol)
    assert torch.equal(q, x)


@pytest.mark.parametrize("dtype", DTYPES)
def test_threshold_policy(dtype):
    n, d = 1764, 768
    c, p = dyadic((1, n, d), 11).to(dtype), dyadic((1, n, d), 12).to(dtype)
    keep = (torch.rand(1, n, 1, generator=torch.Generator().manual_seed(13)) < 0.7)
    p = torch.where(keep, c, p)
    e = (c - p)
    norm = orc.token_norm(e)
    for thr in (0.0, float(norm.float().median()), 0.2, 1e9):
        want = orc.select_threshold(norm, thr)
        got = policies.TokenNormThreshold(threshold=thr)(e.to(DEV))
        assert got.shape == want.shape and torch.equal(got.cpu(), want), thr
    with pytest.raises(AssertionError):
        policies.TokenNormThreshold(threshold=0.5)(torch.zeros(2, 8, 32, device=DEV, dtype=dtype))


def test_topk_k_out_of_range_raises_like_torch():
    x = torch.zeros(1, 16, 32, device=DEV)
    with pytest.raises(RuntimeError):
        policies.TokenNormTopK(k=17)(x)
    assert policies.TokenNormTopK(k=16)(x).sort()[0].tolist() == [list(range(16))]
    assert policies.TokenNormTopFraction(0.0)(x).shape == (1, 0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("structure", ["row", "col"])
def test_standalone_modules_match_oracle(dtype, structure):
    """TokenGate / TokenDeltaGate / TokenBuffer used directly (4-D inputs, (B, k) index) vs the oracle."""
    g = torch.Generator().manual_seed(21)
    b, h, n, d, k = 2, 3, 40, 32, 9
    frames = [(torch.randint(-8, 9, (b, h, n, d), generator=g).float() / 4).to(dtype) for _ in range(3)]
    idxs = [torch.stack([torch.randperm(n if structure == "row" else d, generator=g)[:k] for _ in range(b)]) for _ in range(3)]
    gate, st = modules.TokenDeltaGate(structure=structure), {}
    buf, bst = modules.TokenBuffer(structure=structure), {}
    for t, c in enumerate(frames):
        forced = None if t == 0 else idxs[t]
        got = gate(c.clone().to(DEV), forced_index=None if forced is None else forced.to(DEV))
        want = orc.token_gate(st, c.clone(), forced_index=forced, structure=structure, delta=True)
        for a, w in zip(got[:2], want[:2]):
            assert (a is None) == (w is None)
            if a is not None:
                assert torch.equal(a.cpu(), w)
        assert torch.equal(gate.p.cpu(), st["p"])
        upd = got[0] if t else c.to(DEV)
        out = buf(upd, None if t == 0 else idxs[t].to(DEV))
        ref = orc.token_buffer(bst, want[0] if t else c, None if t == 0 else idxs[t], structure=structure)
        assert torch.equal(out.cpu(), ref) and out.data_ptr() == buf.b.data_ptr()


@pytest.mark.parametrize("dtype", DTYPES)
def test_policy_driven_token_gate_and_stgt(dtype):
    g = torch.Generator().manual_seed(31)
    frames = [dyadic((2, 64, 32), 40 + t).to(dtype) for t in range(3)]
    gate, st = modules.TokenGate(), {}
    gate.policy = policies.TokenNormTopK(k=10)
    sg, sst = modules.SimpleSTGTGate(), {}
    sg.policy = policies.TokenNormTopK(k=10)
    pol = orc.make_policy("topk", k=10)
    for c in frames:
        got, want = gate(c.clone().to(DEV)), orc.token_gate(st, c.clone(), policy=pol)
        assert torch.equal(got[0].cpu(), want[0]) and (got[1] is None) == (want[1] is None)
        if got[1] is not None:
            assert torch.equal(got[1].cpu(), want[1])
        got, want = sg(c.clone().to(DEV)), orc.stgt_gate(sst, c.clone(), pol)
        assert torch.equal(got[0].cpu(), want[0])
        if got[1] is not None:
            assert torch.equal(got[1].cpu(), want[1])
    # save_status keeps debugging copies (policies.py:64-67)
    pol2 = policies.TokenNormTopK(k=3, save_status=True)
    e = frames[0].to(DEV)
    out = pol2(e)
    assert torch.equal(pol2.last_input, e) and torch.equal(pol2.last_output, out)
