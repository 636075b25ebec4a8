"""
ctypes binding of libeventful_b200.so (C ABI: include/eventful_b200.h).

PyTorch is used for device memory, streams and tensor metadata only; every
compute call goes through the shared library.  There is no CPU path and no
fallback: a missing library, a non-CUDA tensor or a non-sm_100 device raises.
"""

import ctypes
import os
from ctypes import c_float, c_int, c_int32, c_int64, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "EVENTFUL_B200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libeventful_b200.so")
)

ET_F32, ET_BF16, ET_F16 = 0, 1, 2
SELECT_TOPK, SELECT_THRESHOLD = 0, 1
ACT_NONE, ACT_GELU = 0, 1
ATTN_DENSE, ATTN_FIRST, ATTN_DELTA = 0, 1, 2

_DTYPES = {torch.float32: ET_F32, torch.bfloat16: ET_BF16, torch.float16: ET_F16}

_SIGNATURES = {
    "et_version": (c_int, []),
    "et_launch_count": (c_longlong, []),
    "et_last_error": (ctypes.c_char_p, []),
    "et_device_info": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "et_gate_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int64,
                               c_int64, c_int64, c_int, c_int, c_int64, c_float, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "et_gate_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p,
                               c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "et_gate_gather_cols": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                    c_int, c_void_p, c_void_p, c_void_p]),
    "et_buffer_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                  c_int64, c_int, c_int, c_void_p]),
    "et_add": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "et_sub": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "et_debug_set": (c_int, [c_int, c_longlong]),
    "et_debug_elapsed_ms": (c_float, []),
    "et_linear": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64,
                          c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "et_linear_gather": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "et_window_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                    c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "et_attn_workspace_bytes": (c_int64, [c_int64] * 9 + [c_int]),
    "et_global_attention": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                    c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "et_ats_scores": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "et_global_attention_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                         c_int, c_int, c_void_p]),
    "et_patchify": (c_int, [c_void_p, c_void_p] + [c_int64] * 8 + [c_int, c_void_p]),
    "et_pool_kv": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "et_pool_index": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p,
                              c_void_p, c_void_p]),
    "et_bmm": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int64] * 5 + [ctypes.POINTER(c_int64)] * 3
               + [c_int, c_float, c_int, c_void_p]),
}

_lib = None
_device_ok = set()


def exported_symbols():
    """Names include/eventful_b200.h declares; tests check the .so exports each of them."""
    return sorted(_SIGNATURES)


def lib():
    """Loads the shared library once.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"eventful_b200: {LIB_PATH} not found. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C eventful-transformer_b200/csrc`. "
                "There is no CPU / PyTorch fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class NativeError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        msg = lib().et_last_error().decode("utf-8", "replace")
        if "selected index k out of range" in msg:
            raise RuntimeError(msg)  # torch.topk raises RuntimeError for k > N (policies.py:63)
        raise NativeError(f"{what} failed (status {rc}): {msg}")


def require_device(t):
    """Every tensor that reaches a kernel must live on an sm_100 CUDA device."""
    if not t.is_cuda:
        raise RuntimeError(
            "eventful_b200 runs on CUDA (sm_100a) only: got a tensor on '%s'; there is no CPU fallback" % t.device
        )
    index = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if index not in _device_ok:
        major, minor, sms = c_int(), c_int(), c_int()
        _check(lib().et_device_info(index, ctypes.byref(major), ctypes.byref(minor), ctypes.byref(sms)),
               "et_device_info")
        _device_ok.add(index)
    return index


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"eventful_b200: unsupported dtype {t.dtype}") from None


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _dense(t, name):
    if not t.is_contiguous():
        raise ValueError(f"eventful_b200: {name} must be contiguous (shape {tuple(t.shape)}, strides {t.stride()})")
    return t


# ----------------------------------------------------------------------------
# workspaces that must persist between calls
# ----------------------------------------------------------------------------
_tickets = {}


def _ticket(device, rows):
    # The last-CTA ticket of et_gate_select is a self-resetting counter per gate row: launches that may run
    # concurrently must not share it, so the workspace is keyed by (device, CUDA stream).  (Kernels of one stream
    # are ordered; a captured graph keeps the tensor of the stream it was captured on.)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _tickets.get(key)
    if t is None or t.numel() < rows:
        t = torch.zeros(max(rows, 64), dtype=torch.int32, device=device)
        _tickets[key] = t
    return t


# ----------------------------------------------------------------------------
# op wrappers (torch tensors in, torch tensors out)
# ----------------------------------------------------------------------------
def gate_select(xa, p=None, xb=None, want_sum=False, ln=None, eps=1e-6, k=None, threshold=None, ticket=None,
                device_count=False, c_out=None):
    """
    Fused [add] -> [LayerNorm] -> delta norm -> selection.  xa: (..., N, D).
    Returns (index (..., k) int64, xsum or None).  With `threshold` the result is
    (index (..., n), xsum) after ONE device->host read of the count, as the
    reference's nonzero() does (policies.py:27-28) -- or, with device_count=True,
    (index padded to (..., N), xsum, count (rows,) int32 on the device) with no host
    synchronisation at all (CUDA-graph capturable; consumers take the count pointer).
    `c_out` (same shape as xa): receives the gate input c = LN(x) of every token, the gather source of
    linear_gather().
    """
    require_device(xa)
    _dense(xa, "gate input")
    n, d = xa.shape[-2], xa.shape[-1]
    lead = tuple(xa.shape[:-2])
    rows = 1
    for s in lead:
        rows *= s
    if threshold is None and int(k) == 0:  # nothing to select (e.g. TokenNormTopFraction(0.0))
        xsum = add(xa, xb) if (xb is not None and want_sum) else None
        return torch.empty(lead + (0,), dtype=torch.int64, device=xa.device), xsum
    xsum = torch.empty_like(xa) if (xb is not None and want_sum) else None
    norm = torch.empty((rows, n), dtype=torch.float32, device=xa.device)
    if ticket is None:
        ticket = _ticket(xa.device, rows)
    if threshold is None:
        idx = torch.empty(lead + (int(k),), dtype=torch.int64, device=xa.device)
        count, mode, kk, thr = None, SELECT_TOPK, int(k), 0.0
    else:
        idx = torch.empty(lead + (n,), dtype=torch.int64, device=xa.device)
        count = torch.empty((rows,), dtype=torch.int32, device=xa.device)
        mode, kk, thr = SELECT_THRESHOLD, 0, float(threshold)
    for t, name in ((p, "gate state"), (xb, "residual"), (c_out, "gate input copy")):
        if t is not None:
            _dense(t, name)
            if t.shape != xa.shape or t.dtype != xa.dtype:
                raise ValueError(f"eventful_b200: {name} shape/dtype mismatch")
    ln_w, ln_b = (None, None) if ln is None else ln
    _check(lib().et_gate_select(_p(xa), _p(xb), _p(xsum), _p(c_out), _p(ln_w), _p(ln_b), float(eps), _p(p), rows, n, d,
                                dtype_code(xa), mode, kk, thr, _p(norm), _p(idx), _p(count), _p(ticket), _stream()),
           "et_gate_select")
    if threshold is not None:
        if device_count:
            return idx, xsum, count
        found = int(count[0].item())  # host sync, like nonzero()
        idx = idx[..., :found]
    return idx, xsum


def gate_gather(x, idx, p=None, ln=None, eps=1e-6, ln_after=False, want_delta=False, full_replace=False, count=None):
    """c~ = LN?(x)[idx]; optional e~ = c~ - p[idx]; p[idx] = c~.  x: (..., N, D), idx: (..., k).
    `count` (rows,) int32 on the device: only the first count[r] indices of each row are valid."""
    require_device(x)
    _dense(x, "gate input")
    _dense(idx, "index")
    n, d, k = x.shape[-2], x.shape[-1], idx.shape[-1]
    lead = tuple(x.shape[:-2])
    if tuple(idx.shape[:-1]) != lead:
        raise ValueError("eventful_b200: index leading dims must match the input's")
    rows = 1
    for s in lead:
        rows *= s
    if p is not None:
        _dense(p, "gate state")
    c_tilde = torch.empty(lead + (k, d), dtype=x.dtype, device=x.device)
    e_tilde = torch.empty_like(c_tilde) if want_delta else None
    ln_w, ln_b = (None, None) if ln is None else ln
    if k == 0 and not full_replace:
        return c_tilde, e_tilde
    _check(lib().et_gate_gather(_p(x), _p(ln_w), _p(ln_b), float(eps), int(ln_after), _p(p), _p(idx), _p(count), rows, n,
                                d, k, dtype_code(x), _p(c_tilde), _p(e_tilde), int(full_replace), _stream()),
           "et_gate_gather")
    return c_tilde, e_tilde


def gate_gather_cols(c, idx, p=None, want_delta=False):
    """Column gate on c (..., N, M) with idx (B, k) shared by all rows of a batch entry."""
    require_device(c)
    _dense(c, "gate input")
    if p is not None:
        _dense(p, "gate state")
    n, m, k = c.shape[-2], c.shape[-1], idx.shape[-1]
    rows = c.numel() // (n * m)
    rpi = rows // (idx.numel() // k)
    c_tilde = torch.empty(tuple(c.shape[:-1]) + (k,), dtype=c.dtype, device=c.device)
    e_tilde = torch.empty_like(c_tilde) if want_delta else None
    _check(lib().et_gate_gather_cols(_p(c), _p(p), _p(_dense(idx, "index")), rows, rpi, n, m, k, dtype_code(c),
                                     _p(c_tilde), _p(e_tilde), _stream()), "et_gate_gather_cols")
    return c_tilde, e_tilde


def buffer_scatter(buf, x, idx, structure="row"):
    require_device(buf)
    _dense(buf, "buffer"), _dense(x, "update"), _dense(idx, "index")
    n, d, k = buf.shape[-2], buf.shape[-1], idx.shape[-1]
    rows = buf.numel() // (n * d)
    if k == 0:
        return buf
    rpi = rows // max(1, idx.numel() // k)
    _check(lib().et_buffer_scatter(_p(buf), _p(x), _p(idx), None, rows, rpi, n, d, k, dtype_code(buf),
                                   0 if structure == "row" else 1, _stream()), "et_buffer_scatter")
    return buf


def add(a, b, out=None):
    require_device(a)
    _dense(a, "a"), _dense(b, "b")
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError(f"eventful_b200.add: operands must match ({tuple(a.shape)} vs {tuple(b.shape)})")
    out = torch.empty_like(a) if out is None else out
    _check(lib().et_add(_p(a), _p(b), _p(out), a.numel(), dtype_code(a), _stream()), "et_add")
    return out


def sub(a, b):
    require_device(a)
    _dense(a, "a"), _dense(b, "b")
    out = torch.empty_like(a)
    _check(lib().et_sub(_p(a), _p(b), _p(out), a.numel(), dtype_code(a), _stream()), "et_sub")
    return out


def linear(x, weight, bias, act=ACT_NONE, out=None, idx=None, n_out_rows=0, count=None):
    """
    y = act(x @ W^T + b) on tcgen05.  x: (..., K).  With `idx` (B, k) and `out` (B, N, F) the rows
    are scattered into the TokenBuffer: out[b, idx[b, j]] = y[b, j].
    """
    require_device(x)
    _dense(x, "linear input"), _dense(weight, "weight")
    k_dim = x.shape[-1]
    m = x.numel() // k_dim
    f = weight.shape[0]
    if idx is None:
        if out is None:
            out = torch.empty(tuple(x.shape[:-1]) + (f,), dtype=x.dtype, device=x.device)
        kk, rows_out = 1, 0
    else:
        _dense(idx, "index"), _dense(out, "buffer")
        kk, rows_out = idx.shape[-1], out.shape[-2]
        if kk == 0:
            return out
    if weight.dtype != x.dtype or (bias is not None and bias.dtype != x.dtype):
        raise TypeError(f"eventful_b200.linear: input is {x.dtype}, weight {weight.dtype}; cast the model or the input")
    _check(lib().et_linear(_p(x), m, k_dim, _p(weight), _p(bias), f, int(act), _p(out), out.shape[-1], _p(idx),
                           _p(count), kk, rows_out, dtype_code(x), _stream()), "et_linear")
    return out


def linear_gather(src, a_idx, weight, bias, state=None, act=ACT_NONE, out=None, idx=None, count=None):
    """
    y = act(src[b, a_idx[b, j]] @ W^T + b) with the gather done by the GEMM's TMA producer (tile::gather4) and, with
    `state` (same shape as src), state[b, a_idx[b, j]] = src[b, a_idx[b, j]] advanced by the same kernel.
    src: (B, N, K), a_idx: (B, k).  `idx` / `out` as in linear() (TokenBuffer scatter); without them y is (B, k, F).
    """
    require_device(src)
    _dense(src, "gather source"), _dense(weight, "weight"), _dense(a_idx, "gather index")
    if src.dtype not in (torch.bfloat16, torch.float16):
        raise TypeError("eventful_b200.linear_gather: 16-bit dtypes only")
    n, k_dim = src.shape[-2], src.shape[-1]
    kk = a_idx.shape[-1]
    lead = tuple(src.shape[:-2])
    if tuple(a_idx.shape[:-1]) != lead:
        raise ValueError("eventful_b200: index leading dims must match the source's")
    f = weight.shape[0]
    m = a_idx.numel()
    if state is not None:
        _dense(state, "gate state")
        if state.shape != src.shape or state.dtype != src.dtype:
            raise ValueError("eventful_b200: gate state shape/dtype mismatch")
    if idx is None:
        if out is None:
            out = torch.empty(lead + (kk, f), dtype=src.dtype, device=src.device)
        rows_out = 0
    else:
        _dense(idx, "index"), _dense(out, "buffer")
        if idx.shape != a_idx.shape:
            raise ValueError("eventful_b200: scatter and gather index shapes differ")
        rows_out = out.shape[-2]
    if kk == 0:
        return out
    if weight.dtype != src.dtype or (bias is not None and bias.dtype != src.dtype):
        raise TypeError(f"eventful_b200.linear_gather: input is {src.dtype}, weight {weight.dtype}")
    _check(lib().et_linear_gather(_p(src), n, _p(a_idx), _p(state), m, k_dim, _p(weight), _p(bias), f, int(act), _p(out),
                                  out.shape[-1], _p(idx), _p(count), kk, rows_out, dtype_code(src), _stream()),
           "et_linear_gather")
    return out


def attn_workspace(b, n, gh, gw, wh, ww, heads, dh, k, has_rel, device):
    nbytes = lib().et_attn_workspace_bytes(b, n, gh, gw, wh, ww, heads, dh, k, int(has_rel))
    return torch.empty((nbytes,), dtype=torch.uint8, device=device)


def window_attention(qkv, heads, grid, window, pad_token=None, rel=None):
    """Dense (windowed) attention on a (B, N, 3D) QKV tensor -> (B, N, D)."""
    require_device(qkv)
    _dense(qkv, "qkv")
    b, n, d3 = qkv.shape
    d = d3 // 3
    dh = d // heads
    gh, gw = grid
    wh, ww = window if window is not None else (0, 0)
    out = torch.empty((b, n, d), dtype=qkv.dtype, device=qkv.device)
    need_ws = rel is not None or qkv.dtype == torch.float32
    ws = attn_workspace(b, n, gh, gw, wh, ww, heads, dh, 0, rel is not None, qkv.device) if need_ws else None
    rel_y, rel_x = (None, None) if rel is None else rel
    _check(lib().et_window_attention(_p(qkv), _p(pad_token), _p(rel_y), _p(rel_x), _p(out), _p(ws), b, n, gh, gw, wh,
                                     ww, heads, dh, dtype_code(qkv), _stream()), "et_window_attention")
    return out


def global_attention(qkv, heads, grid, mode, rel=None, idx=None, a_state=None, v_state=None, acc=None, stats=None,
                     count=None, kv_pooled=None, pool=None, state_dtype=None):
    """
    Global attention over the QKV buffer in DENSE / FIRST / DELTA mode -> fresh (B, N, D) output (model dtype).
    kv_pooled / pool: pooled keys and values (pool_kv) and the pooling ratio; count: device-side number of valid
    entries of idx per batch entry; state_dtype: element type of a_state / v_state / acc (matmul_2_cast).
    """
    require_device(qkv)
    _dense(qkv, "qkv")
    b, n, d3 = qkv.shape
    d = d3 // 3
    dh = d // heads
    gh, gw = grid
    k = 0 if idx is None else idx.shape[-1]
    if mode == ATTN_DELTA and k == 0:
        return acc.to(qkv.dtype)  # nothing selected: the accumulator is unchanged (a copy, never the state itself)
    out = torch.empty((b, n, d), dtype=qkv.dtype, device=qkv.device)
    if stats is None:
        stats = torch.empty((b, heads, n, 2), dtype=torch.float32, device=qkv.device)
    sdt = _DTYPES[qkv.dtype if state_dtype is None else state_dtype]
    ws = attn_workspace(b, n, gh, gw, 0, 0, heads, dh, 2 * k if sdt == ET_F32 else k, rel is not None, qkv.device)
    rel_y, rel_x = (None, None) if rel is None else rel
    ph, pw = (1, 1) if pool is None else pool
    for t, name in ((idx, "index"), (kv_pooled, "pooled keys"), (a_state, "A-gate state"), (v_state, "v-gate state"),
                    (acc, "accumulator")):
        if t is not None:
            _dense(t, name)
    _check(lib().et_global_attention(_p(qkv), _p(kv_pooled), ph, pw, _p(rel_y), _p(rel_x), int(mode), _p(idx), _p(count), k,
                                     _p(a_state), _p(v_state), _p(acc), _p(out), _p(stats), _p(ws), b, n, gh, gw, heads,
                                     dh, dtype_code(qkv), sdt, _stream()), "et_global_attention")
    return out


def ats_scores(qkv, heads, score_dtype=None):
    """Adaptive token sampling, scoring pass: raw[b, h, t] = softmax(q k^T / sqrt(dh))[b, h, t, 0] * |v[b, h, t]| as a
    (B, H, N) tensor of `score_dtype` (the dtype the reference holds a and v in at that point; default: the model's)."""
    require_device(qkv)
    _dense(qkv, "qkv")
    b, n, d3 = qkv.shape
    dh = d3 // 3 // heads
    score_dtype = qkv.dtype if score_dtype is None else score_dtype
    stats = torch.empty((b, heads, n, 2), dtype=torch.float32, device=qkv.device)
    raw = torch.empty((b, heads, n), dtype=torch.float32, device=qkv.device)
    _check(lib().et_ats_scores(_p(qkv), b, n, heads, dh, dtype_code(qkv), _DTYPES[score_dtype], _p(stats), _p(raw), _stream()),
           "et_ats_scores")
    return raw.to(score_dtype)  # exact: every value was rounded to score_dtype by the kernel


def global_attention_rows(qkv, q_index, heads, mode, idx=None, a_state=None, v_state=None, acc=None, stats=None, count=None,
                          state_dtype=None):
    """et_global_attention for the query tokens q_index (B, Nq) only (adaptive token sampling): keys and values are all N
    tokens of qkv (B, N, 3D); a_state (B, H, N, NP(Nq)), acc / out (B, Nq, D)."""
    require_device(qkv)
    _dense(qkv, "qkv"), _dense(q_index, "query index")
    b, n, d3 = qkv.shape
    d = d3 // 3
    dh = d // heads
    nq = q_index.shape[-1]
    if q_index.shape[0] != b or q_index.dtype != torch.int64:
        raise ValueError("eventful_b200.global_attention_rows: the query index must be (B, Nq) int64")
    k = 0 if idx is None else idx.shape[-1]
    if mode == ATTN_DELTA and k == 0:
        return acc.to(qkv.dtype)
    out = torch.empty((b, nq, d), dtype=qkv.dtype, device=qkv.device)
    if stats is None:
        stats = torch.empty((b, heads, nq, 2), dtype=torch.float32, device=qkv.device)
    sdt = _DTYPES[qkv.dtype if state_dtype is None else state_dtype]
    ws = attn_workspace(b, n, 0, 0, 0, 0, heads, dh, 2 * k if sdt == ET_F32 else k, False, qkv.device)
    for t, name in ((idx, "index"), (a_state, "A-gate state"), (v_state, "v-gate state"), (acc, "accumulator")):
        if t is not None:
            _dense(t, name)
    _check(lib().et_global_attention_rows(_p(qkv), _p(q_index), nq, int(mode), _p(idx), _p(count), k, _p(a_state), _p(v_state),
                                          _p(acc), _p(out), _p(stats), _p(ws), b, n, heads, dh, dtype_code(qkv), sdt, _stream()),
           "et_global_attention_rows")
    return out


def pool_kv(qkv, grid, pool):
    """(B, gh*gw, 3D) -> (B, Nk, 2D) = [k | v] averaged over pool cells (Block._pool_tokens, blocks.py:303-326)."""
    require_device(qkv)
    _dense(qkv, "qkv")
    b, n, d3 = qkv.shape
    d = d3 // 3
    gh, gw = grid
    if gh * gw != n:
        raise ValueError("eventful_b200.pool_kv: K/V pooling needs a token grid without extra tokens")
    out = torch.empty((b, (gh // pool[0]) * (gw // pool[1]), 2 * d), dtype=qkv.dtype, device=qkv.device)
    _check(lib().et_pool_kv(_p(qkv), _p(out), b, gh, gw, d, pool[0], pool[1], dtype_code(qkv), _stream()), "et_pool_kv")
    return out


def pool_index(idx, count, grid, pool):
    """Token indices (B, k) [+ device count] -> (pooled-cell ids (B, k) ascending unique, count (B,) int32)."""
    require_device(idx)
    _dense(idx, "index")
    b, k = idx.shape
    out = torch.empty_like(idx)
    out_count = torch.empty((b,), dtype=torch.int32, device=idx.device)
    if k == 0:
        return out, out_count.zero_()
    _check(lib().et_pool_index(_p(idx), _p(count), b, k, grid[0], grid[1], pool[0], pool[1], _p(out), _p(out_count),
                               _stream()), "et_pool_index")
    return out, out_count


def _batch2(t):
    """(tensor, (outer, inner)) with exactly two leading batch dims, without copying when the layout allows it."""
    if t.dim() == 2:
        return t.unsqueeze(0).unsqueeze(0)
    if t.dim() == 3:
        return t.unsqueeze(0)
    if t.dim() == 4:
        return t
    lead = t.shape[:-3]  # more than two batch dims: fold all but the last one (a view when possible, else a copy)
    return t.reshape((-1,) + tuple(t.shape[-3:])) if len(lead) else t


def bmm(a, b, out=None, accumulate=False, alpha=1.0):
    """Batched matmul alpha * (..., M, K) x (..., K, N) -> (..., M, N); strided views welcome (element strides are
    passed down, two batch levels natively)."""
    require_device(a)
    if a.dtype != b.dtype or a.shape[:-2] != b.shape[:-2] or a.shape[-1] != b.shape[-2]:
        raise ValueError(f"eventful_b200.bmm: operands do not match ({tuple(a.shape)} x {tuple(b.shape)})")
    m, kd, n = a.shape[-2], a.shape[-1], b.shape[-1]
    if out is None:
        out = torch.empty(tuple(a.shape[:-2]) + (m, n), dtype=a.dtype, device=a.device)
    a4, b4, o4 = _batch2(a), _batch2(b), _batch2(out)
    if o4.data_ptr() != out.data_ptr():
        raise ValueError("eventful_b200.bmm: the output must be viewable with two batch dims")
    arr = lambda t: (c_int64 * 4)(*t.stride())  # noqa: E731
    _check(lib().et_bmm(_p(a4), _p(b4), _p(o4), o4.shape[0], o4.shape[1], m, n, kd, arr(a4), arr(b4), arr(o4),
                        int(accumulate), float(alpha), dtype_code(a), _stream()), "et_bmm")
    return out


def patchify(x, patch):
    """
    Non-overlapping patches / tubelets as GEMM rows.  x (B, C, H, W) with patch (ph, pw), or x (B, T, C, H, W) with patch
    (pt, ph, pw)  ->  (B, N, C*ph*pw) or (B, T/pt, N, C*pt*ph*pw), features in the order of the flattened conv weight.
    """
    require_device(x)
    _dense(x, "image / video")
    if x.dim() == 4:
        (b, c, h, w), t = x.shape, 1
        pt, (ph, pw) = 1, patch
    else:
        b, t, c, h, w = x.shape
        pt, ph, pw = patch
    n, f = (h // ph) * (w // pw), c * pt * ph * pw
    out = torch.empty((b, n, f) if x.dim() == 4 else (b, t // pt, n, f), dtype=x.dtype, device=x.device)
    _check(lib().et_patchify(_p(x), _p(out), b, t, c, h, w, pt, ph, pw, dtype_code(x), _stream()), "et_patchify")
    return out
