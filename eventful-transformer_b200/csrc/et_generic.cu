// General-precision path of the gated update (fp32 arithmetic on the CUDA cores, any element type in memory).
//
// The tcgen05 kernels (et_gemm.cu, et_attn_tc.cu, et_attn_win_tc.cu) cover 16-bit models whose attention state has the
// model dtype.  Everything else the reference can be configured to do goes through this file:
//   * fp32 models (BASELINE configs[0]; the reference's own timed CUDA config is an fp32 model, configs/time/vitdet_vid/_cuda.yml)
//   * matmul_2_cast != model dtype: a, v, the v-gate / A-gate state and the accumulator live in fp16 / bf16 while q k^T and
//     the softmax stay in the model dtype (blocks.py:183-189,561-562,574)
//   * K/V token pooling (blocks.py:303-326): keys / values come from a pooled (B, Nk, 2D) tensor, the A-gate state is
//     (N x Nk), rel-pos tables are pooled along the key axis (utils.py:185-188)
//   * a device-side number of selected tokens per batch entry (threshold policy, pooled unique indices)
//
//   sgemm_kernel        y = act(A W^T + b) in fp32 with the TokenBuffer scatter epilogue           (counting.py:157-162)
//   gen_stats_kernel    row max / row sum of softmax(q k^T / s + rel-pos) over all keys            (blocks.py:223-226)
//   gen_apply_kernel    DENSE / FIRST: out = a v (+ state init); DELTA: A-gate + accumulator       (modules.py:187-201,285-295)
//   gen_vgate_kernel    TokenDeltaGate on the (possibly pooled, possibly cast) v rows              (modules.py:187-201)
//   pool_kv_kernel      avg_pool2d of k and v over the token grid                                  (blocks.py:303-326)
//   pool_index_kernel   token index -> sorted unique pooled-cell index, with a device-side count   (blocks.py:525-540)
#include "et_common.cuh"

namespace {

// ---------------------------------------------------------------- runtime-typed element access
__device__ __forceinline__ float ld_elem(const void* p, long long i, int dt) {
    if (dt == ET_F32) return static_cast<const float*>(p)[i];
    if (dt == ET_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
    return __half2float(static_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_elem(void* p, long long i, int dt, float v) {
    if (dt == ET_F32) static_cast<float*>(p)[i] = v;
    else if (dt == ET_BF16) static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else static_cast<__half*>(p)[i] = __float2half_rn(v);
}
__device__ __forceinline__ float round_dt(float v, int dt) {
    if (dt == ET_F32) return v;
    if (dt == ET_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    return __half2float(__float2half_rn(v));
}
inline size_t elem_size(int dt) { return dt == ET_F32 ? 4 : 2; }

// ================================================================= SGEMM with scatter epilogue
constexpr int GM = 128, GN = 128, GK = 16, GTHREADS = 256, GLD = GM + 4;

struct SgemmArgs {
    const float* A;
    const float* W;
    const float* bias;
    float* out;
    const long long* idx;
    const int* count;
    long long ld_out;
    int M, K, NF, act, k, n_out_rows;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(GTHREADS) sgemm_kernel(const SgemmArgs a) {
    et_pdl_prologue();
    __shared__ __align__(16) float As[2][GK][GLD];
    __shared__ __align__(16) float Ws[2][GK][GLD];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    // loader mapping: 128 rows x 4 quads of k per operand tile = 512 float4, two per thread
    const int lrow = tid >> 2, lq = tid & 3;
    float4 ra[2], rw[2];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + 64 * i, kk = k0 + lq * 4;
            const int m = m0 + r, n = n0 + r;
            ra[i] = (m < a.M && kk < a.K) ? *reinterpret_cast<const float4*>(a.A + (size_t)m * a.K + kk) : make_float4(0, 0, 0, 0);
            rw[i] = (n < a.NF && kk < a.K) ? *reinterpret_cast<const float4*>(a.W + (size_t)n * a.K + kk) : make_float4(0, 0, 0, 0);
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lrow + 64 * i;
            As[buf][lq * 4 + 0][r] = ra[i].x; As[buf][lq * 4 + 1][r] = ra[i].y;
            As[buf][lq * 4 + 2][r] = ra[i].z; As[buf][lq * 4 + 3][r] = ra[i].w;
            Ws[buf][lq * 4 + 0][r] = rw[i].x; Ws[buf][lq * 4 + 1][r] = rw[i].y;
            Ws[buf][lq * 4 + 2][r] = rw[i].z; Ws[buf][lq * 4 + 3][r] = rw[i].w;
        }
    };
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    fetch(0);
    stash(0);
    __syncthreads();
    const int nk = (a.K + GK - 1) / GK;
    for (int t = 0; t < nk; ++t) {
        const int buf = t & 1;
        if (t + 1 < nk) fetch((t + 1) * GK);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (t + 1 < nk) {
            stash(buf ^ 1);
            __syncthreads();
        }
    }
    // epilogue: bias, activation, rows redirected through idx (TokenBuffer scatter, modules.py:96)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= a.M) continue;
        long long orow = m;
        if (a.idx != nullptr) {
            const int b = m / a.k, j = m - b * a.k;
            if (a.count != nullptr && j >= a.count[b]) continue;
            orow = (long long)b * a.n_out_rows + a.idx[m];
        }
        float* dst = a.out + orow * a.ld_out;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int n = n0 + half * 64 + tx * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j >= a.NF) continue;
                float v = acc[i][half * 4 + j] + (a.bias != nullptr ? a.bias[n + j] : 0.f);
                if (a.act == ET_ACT_GELU) v = gelu_erf(v);
                dst[n + j] = v;
            }
        }
    }
}

// ================================================================= generic attention
constexpr int TQ = 64, TK = 64, TD = 64, LDT = 68;  // query rows / keys per tile, padded head dim, smem row stride
constexpr int ATHREADS = 256;

struct GenArgs {
    const void* q;      // (B, Nq, ldq) token-major, head h at element offset h * dh
    const void* kk;     // keys   (B, Nk, ldkv)
    const void* vv;     // values (B, Nk, ldkv)
    const void* pad;    // window pad token (3D elements: q | k | v parts) or null
    const void* rel_y;  // (ah, kh, dh)
    const void* rel_x;  // (aw, kw, dh)
    long long ldq, ldkv;
    int B, Nq, Nk, H, dh, D;
    int gh, gw, wh, ww, nwx, nwy, windowed, Wn;  // Wn = tokens a CTA group attends over (window size or Nq)
    int aw;                                      // query-grid width used for rel-pos coordinates
    int kh, kw;                                  // key grid (pooled) for rel-pos coordinates
    float rscale;
    int mode;
    const long long* idx;
    const int* count;
    int kmax;
    void* a_state;  // (B, H, Nk, NP) column-major per head, state dtype
    int NP;
    void* acc;      // (B, Nq, D) state dtype
    void* out;      // (B, Nq, D) model dtype
    float* stats;   // (B * nwin, H, Wn, 2)
    const void* dV;  // (B, kmax, D) state dtype
    const void* Vr;  // (B, kmax, D) state dtype
    int dt, sdt;     // model dtype, state dtype
    float* ats_raw;  // optional (B, H, Nq): adaptive-token-sampling raw scores a[b, h, t, 0] * |v[b, h, t]| (gen_stats_kernel)
    int ats_dt;      // dtype the reference holds a and v in at that point (model dtype, or the matmul_2_cast dtype)
    int stats_only;  // launch_generic: statistics pass only
    const long long* q_index;  // optional (B, Nq): query row t of batch entry b is row q_index[b][t] of q (adaptive token sampling)
    int q_batch_rows;          // rows per batch entry of the q tensor (Nq unless q_index selects among more rows)
};

// window-local token -> global token row, or -1 for padding
__device__ __forceinline__ int map_token(const GenArgs& a, int win, int t) {
    if (!a.windowed) return t;
    const int wy = win / a.nwx, wx = win - wy * a.nwx;
    const int ly = t / a.ww, lx = t - ly * a.ww;
    const int gy = wy * a.wh + ly, gx = wx * a.ww + lx;
    return (gy < a.gh && gx < a.gw) ? gy * a.gw + gx : -1;
}

// Stages TQ query rows (transposed: Qt[c][r]) and computes the decomposed rel-pos bias rows bh[r][ky], bw[r][kx].
__device__ __forceinline__ void load_queries(const GenArgs& a, int b, int h, int win, int q0, float* Qt, float* bh, float* bw) {
    for (int i = threadIdx.x; i < TQ * TD; i += ATHREADS) {
        const int r = i / TD, c = i - r * TD;
        float v = 0.f;
        const int t = q0 + r;
        if (t < a.Wn && c < a.dh) {
            int tok = map_token(a, win, t);
            if (a.q_index != nullptr) tok = (int)a.q_index[(long long)b * a.Nq + t];
            v = tok >= 0 ? ld_elem(a.q, ((long long)b * a.q_batch_rows + tok) * a.ldq + h * a.dh + c, a.dt)
                         : ld_elem(a.pad, h * a.dh + c, a.dt);
        }
        Qt[c * LDT + r] = v;
    }
    __syncthreads();
    if (a.rel_y == nullptr) return;
    const int ncoord = a.kh + a.kw;
    for (int i = threadIdx.x; i < TQ * ncoord; i += ATHREADS) {
        const int r = i / ncoord, co = i - r * ncoord;
        const int t = q0 + r;
        float s = 0.f;
        if (t < a.Wn) {
            const int qy = t / a.aw, qx = t - qy * a.aw;
            const bool ymode = co < a.kh;
            const void* tab = ymode ? a.rel_y : a.rel_x;
            const long long base = ymode ? ((long long)qy * a.kh + co) * a.dh : ((long long)qx * a.kw + (co - a.kh)) * a.dh;
            for (int c = 0; c < a.dh; ++c) s = fmaf(Qt[c * LDT + r], ld_elem(tab, base + c, a.dt), s);
        }
        if (co < a.kh) bh[r * (a.kh + 1) + co] = s;
        else bw[r * (a.kw + 1) + co - a.kh] = s;
    }
    __syncthreads();
}

// S[4][4] of this thread: rows ty*4.., keys tx*4.. of the staged tiles, scaled, plus bias.
__device__ __forceinline__ void score_tile(const GenArgs& a, const float* Qt, const float* Kt, const float* bh, const float* bw,
                                           const int* s_tok, int ty, int tx, float (&s)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int c = 0; c < TD; ++c) {
        const float4 qv = *reinterpret_cast<const float4*>(Qt + c * LDT + ty * 4);
        const float4 kv = *reinterpret_cast<const float4*>(Kt + c * LDT + tx * 4);
        const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int tok = s_tok[tx * 4 + j];  // key id in the key grid (window-local or global / pooled), -1 = masked
        int ky = 0, kx = 0;
        if (a.rel_y != nullptr && tok >= 0) { ky = tok / a.kw; kx = tok - ky * a.kw; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = s[i][j] * a.rscale;
            if (a.rel_y != nullptr && tok >= 0) v += bh[(ty * 4 + i) * (a.kh + 1) + ky] + bw[(ty * 4 + i) * (a.kw + 1) + kx];
            s[i][j] = tok >= 0 ? v : -INFINITY;
        }
    }
}

// Stages TK key rows (transposed) for the key ids in s_tok; `part` 1 = k, 2 = v of the pad token.
__device__ __forceinline__ void load_keys(const GenArgs& a, int b, int h, int win, const int* s_tok, float* Kt) {
    for (int i = threadIdx.x; i < TK * TD; i += ATHREADS) {
        const int j = i / TD, c = i - j * TD;
        float v = 0.f;
        const int t = s_tok[j];
        if (t >= 0 && c < a.dh) {
            const int tok = map_token(a, win, t);
            v = tok >= 0 ? ld_elem(a.kk, ((long long)b * a.Nk + tok) * a.ldkv + h * a.dh + c, a.dt)
                         : ld_elem(a.pad, a.D + h * a.dh + c, a.dt);
        }
        Kt[c * LDT + j] = v;
    }
}

// grid (ceil(Wn / 64), H, B * n_windows)
__global__ void __launch_bounds__(ATHREADS) gen_stats_kernel(const GenArgs a) {
    et_pdl_prologue();
    extern __shared__ __align__(16) float sm[];
    float* Qt = sm;
    float* Kt = Qt + TD * LDT;
    float* bh = Kt + TD * LDT;
    float* bw = bh + TQ * (a.kh + 1);
    __shared__ int s_tok[TK];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int q0 = blockIdx.x * TQ, h = blockIdx.y;
    const int nwin = a.windowed ? a.nwx * a.nwy : 1;
    const int b = blockIdx.z / nwin, win = blockIdx.z - b * nwin;
    load_queries(a, b, h, win, q0, Qt, bh, bw);
    const int nkeys = a.windowed ? a.Wn : a.Nk;
    float mrow[4], lrow[4], s_cls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) { mrow[i] = -INFINITY; lrow[i] = 0.f; }
    for (int key0 = 0; key0 < nkeys; key0 += TK) {
        __syncthreads();
        if (tid < TK) s_tok[tid] = key0 + tid < nkeys ? key0 + tid : -1;
        __syncthreads();
        load_keys(a, b, h, win, s_tok, Kt);
        __syncthreads();
        float s[4][4];
        score_tile(a, Qt, Kt, bh, bw, s_tok, ty, tx, s);
        if (key0 == 0 && tx == 0) {  // logit of key 0 (the class token) for the ATS scores
#pragma unroll
            for (int i = 0; i < 4; ++i) s_cls[i] = s[i][0];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float mnew = fmaxf(mrow[i], mx);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) rs += expf(s[i][j] - mnew);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            lrow[i] = lrow[i] * expf(mrow[i] - mnew) + rs;
            mrow[i] = mnew;
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = q0 + ty * 4 + i;
            if (t < a.Wn) {
                float* st = a.stats + (((size_t)blockIdx.z * a.H + h) * a.Wn + t) * 2;
                st[0] = mrow[i];
                st[1] = lrow[i];
                if (a.ats_raw != nullptr) {
                    // Block._adaptive_token_sampling (blocks.py:154-155): class_scores * |v|, every factor rounded where the
                    // reference holds it in a 16-bit dtype (softmax in the model dtype, then a and v cast to ats_dt)
                    const float a0 = round_dt(round_dt(expf(s_cls[i] - mrow[i]) / lrow[i], a.dt), a.ats_dt);
                    float ss = 0.f;
                    for (int c = 0; c < a.dh; ++c) {
                        const float v = round_dt(ld_elem(a.vv, ((long long)b * a.Nk + t) * a.ldkv + h * a.dh + c, a.dt), a.ats_dt);
                        ss = fmaf(v, v, ss);
                    }
                    a.ats_raw[((size_t)b * a.H + h) * a.Nq + t] = round_dt(a0 * round_dt(sqrtf(ss), a.ats_dt), a.ats_dt);
                }
            }
        }
    }
}

// grid (ceil(Wn / 64), H, B * n_windows).  Keys: all (DENSE / FIRST) or the selected ones (DELTA).
template <int MODE>
__global__ void __launch_bounds__(ATHREADS) gen_apply_kernel(const GenArgs a) {
    et_pdl_prologue();
    extern __shared__ __align__(16) float sm[];
    float* Qt = sm;
    float* Kt = Qt + TD * LDT;
    float* V1 = Kt + TD * LDT;                                 // [j][c]: v (DENSE / FIRST) or dV (DELTA)
    float* Pn = V1 + TK * LDT;                                 // [j][r]: a_n
    float* V2 = Pn + TK * LDT;                                 // DELTA: v_n - dV
    float* Pd = V2 + (MODE == ET_ATTN_DELTA ? TK * LDT : 0);   // DELTA: old state, then dA
    float* bh = Pd + (MODE == ET_ATTN_DELTA ? TK * LDT : 0);
    float* bw = bh + TQ * (a.kh + 1);
    __shared__ int s_tok[TK];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int q0 = blockIdx.x * TQ, h = blockIdx.y;
    const int nwin = a.windowed ? a.nwx * a.nwy : 1;
    const int b = blockIdx.z / nwin, win = blockIdx.z - b * nwin;
    load_queries(a, b, h, win, q0, Qt, bh, bw);
    int nkeys = a.windowed ? a.Wn : a.Nk;
    if (MODE == ET_ATTN_DELTA) nkeys = a.count != nullptr ? min(a.kmax, a.count[b]) : a.kmax;
    float mrow[4], linv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = min(a.Wn - 1, q0 + ty * 4 + i);
        const float* st = a.stats + (((size_t)blockIdx.z * a.H + h) * a.Wn + t) * 2;
        mrow[i] = st[0];
        linv[i] = 1.f / st[1];
    }
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    const size_t a_head = ((size_t)b * a.H + h) * (size_t)a.Nk * a.NP;

    for (int key0 = 0; key0 < nkeys; key0 += TK) {
        __syncthreads();
        if (tid < TK) {
            const int j = key0 + tid;
            int tok = -1;
            if (j < nkeys) tok = MODE == ET_ATTN_DELTA ? (int)a.idx[(size_t)b * a.kmax + j] : j;
            s_tok[tid] = tok;
        }
        __syncthreads();
        load_keys(a, b, h, win, s_tok, Kt);
        for (int i = tid; i < TK * TD; i += ATHREADS) {
            const int j = i / TD, c = i - j * TD;
            const int t = s_tok[j];
            float v1 = 0.f, v2 = 0.f;
            if (t >= 0 && c < a.dh) {
                if (MODE == ET_ATTN_DELTA) {
                    const long long at = ((long long)b * a.kmax + key0 + j) * a.D + h * a.dh + c;
                    v1 = ld_elem(a.dV, at, a.sdt);
                    v2 = ld_elem(a.Vr, at, a.sdt);
                } else {
                    const int tok = map_token(a, win, t);
                    v1 = tok >= 0 ? ld_elem(a.vv, ((long long)b * a.Nk + tok) * a.ldkv + h * a.dh + c, a.dt)
                                  : ld_elem(a.pad, 2 * a.D + h * a.dh + c, a.dt);
                    v1 = round_dt(v1, a.sdt);  // _cast_matmul_2 (blocks.py:183-189)
                }
            }
            V1[j * LDT + c] = v1;
            if (MODE == ET_ATTN_DELTA) V2[j * LDT + c] = v2;
        }
        if (MODE == ET_ATTN_DELTA) {  // previous attention values of the selected columns (contiguous in the row index)
            for (int i = tid; i < TK * TQ; i += ATHREADS) {
                const int j = i / TQ, r = i - j * TQ;
                const int t = s_tok[j];
                Pd[j * LDT + r] = (t >= 0 && q0 + r < a.Nq) ? ld_elem(a.a_state, a_head + (size_t)t * a.NP + q0 + r, a.sdt) : 0.f;
            }
        }
        __syncthreads();
        float s[4][4];
        score_tile(a, Qt, Kt, bh, bw, s_tok, ty, tx, s);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float an[4], ad[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // softmax value in the model dtype (blocks.py:226,522), then cast to the state dtype (blocks.py:561)
                float v = s[i][j] == -INFINITY ? 0.f : expf(s[i][j] - mrow[i]) * linv[i];
                v = round_dt(round_dt(v, a.dt), a.sdt);
                an[i] = v;
                if (MODE == ET_ATTN_DELTA) ad[i] = round_dt(v - Pd[(tx * 4 + j) * LDT + ty * 4 + i], a.sdt);  // modules.py:196
            }
            *reinterpret_cast<float4*>(Pn + (tx * 4 + j) * LDT + ty * 4) = make_float4(an[0], an[1], an[2], an[3]);
            if (MODE == ET_ATTN_DELTA)
                *reinterpret_cast<float4*>(Pd + (tx * 4 + j) * LDT + ty * 4) = make_float4(ad[0], ad[1], ad[2], ad[3]);
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < TK; ++j) {
            const float4 p4 = *reinterpret_cast<const float4*>(Pn + j * LDT + ty * 4);
            const float4 v4 = *reinterpret_cast<const float4*>(V1 + j * LDT + tx * 4);
            const float pa[4] = {p4.x, p4.y, p4.z, p4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[i][c] = fmaf(pa[i], va[c], o[i][c]);
            if (MODE == ET_ATTN_DELTA) {
                const float4 d4 = *reinterpret_cast<const float4*>(Pd + j * LDT + ty * 4);
                const float4 w4 = *reinterpret_cast<const float4*>(V2 + j * LDT + tx * 4);
                const float da[4] = {d4.x, d4.y, d4.z, d4.w}, wa[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[i][c] = fmaf(da[i], wa[c], o[i][c]);
            }
        }
        if (MODE != ET_ATTN_DENSE) {  // p[:, idx] = a_n (modules.py:200) / state initialisation
            for (int i = tid; i < TK * TQ; i += ATHREADS) {
                const int j = i / TQ, r = i - j * TQ;
                const int t = s_tok[j];
                if (t >= 0 && q0 + r < a.Nq) st_elem(a.a_state, a_head + (size_t)t * a.NP + q0 + r, a.sdt, Pn[j * LDT + r]);
            }
        }
    }
    // accumulator update and merged-head output; windows: crop + recombine (blocks.py:346-376)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = q0 + ty * 4 + i;
        if (t >= a.Wn) continue;
        const int tok = map_token(a, win, t);
        if (tok < 0) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int col = tx * 4 + c;
            if (col >= a.dh) continue;
            const long long at = ((long long)b * a.Nq + tok) * a.D + h * a.dh + col;
            float v = o[i][c];
            if (MODE == ET_ATTN_DELTA) v += ld_elem(a.acc, at, a.sdt);
            v = round_dt(v, a.sdt);
            if (MODE != ET_ATTN_DENSE) st_elem(a.acc, at, a.sdt, v);
            st_elem(a.out, at, a.dt, v);  // _uncast_matmul_2 (blocks.py:393-396)
        }
    }
}

// v-gate with a forced index on (possibly pooled) v rows cast to the state dtype.
//   DELTA: dV = v - p, Vr = v - dV, p = v at the selected rows.  FIRST (idx == null): p = v for every row.
__global__ void __launch_bounds__(256) gen_vgate_kernel(const void* v, long long ldv, void* v_state, const long long* idx,
                                                        const int* count, void* dV, void* Vr, int Nk, int D, int kmax,
                                                        long long total, int dt, int sdt) {
    et_pdl_prologue();
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(gi % D);
        const long long row = gi / D;
        const long long b = row / kmax, j = row - b * kmax;
        if (count != nullptr && j >= count[b]) continue;
        const long long tok = idx != nullptr ? idx[row] : j;
        const float vn = round_dt(ld_elem(v, (b * Nk + tok) * ldv + c, dt), sdt);
        const long long st = (b * Nk + tok) * D + c;
        if (dV != nullptr) {
            const float d = round_dt(vn - ld_elem(v_state, st, sdt), sdt);
            st_elem(dV, row * D + c, sdt, d);
            st_elem(Vr, row * D + c, sdt, round_dt(vn - d, sdt));
        }
        st_elem(v_state, st, sdt, vn);
    }
}

// avg_pool2d of the k and v parts of the QKV buffer over the token grid: out (B, Nk, 2D) = [k_pooled | v_pooled]
__global__ void __launch_bounds__(256) pool_kv_kernel(const void* qkv, void* out, int gh, int gw, int ph, int pw, int D, long long total,
                                                      int dt) {
    et_pdl_prologue();
    const int kh = gh / ph, kw = gw / pw;
    const float inv = 1.f / (float)(ph * pw);
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(gi % (2 * D));
        const long long cell = gi / (2 * D);
        const long long b = cell / (kh * kw);
        const int ci = (int)(cell - b * kh * kw), cy = ci / kw, cx = ci - cy * kw;
        float s = 0.f;
        for (int dy = 0; dy < ph; ++dy)
            for (int dx = 0; dx < pw; ++dx) {
                const long long tok = (long long)(cy * ph + dy) * gw + cx * pw + dx;
                s += ld_elem(qkv, (b * gh * gw + tok) * 3 * D + D + c, dt);
            }
        st_elem(out, gi, dt, s * inv);
    }
}

// one CTA per batch entry: token indices -> pooled cells -> ascending unique list + count
__global__ void __launch_bounds__(256) pool_index_kernel(const long long* idx, const int* count_in, int k, int gw, int ph, int pw,
                                                         int ncells, long long* out_idx, int* out_count) {
    et_pdl_prologue();
    extern __shared__ unsigned int flags[];  // ncells bits, then 256 partial counts
    const int words = (ncells + 31) / 32;
    unsigned int* partial = flags + words;
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int w = tid; w < words; w += 256) flags[w] = 0u;
    __syncthreads();
    const int n = count_in != nullptr ? min(k, count_in[b]) : k;
    const int kwid = gw / pw;
    for (int j = tid; j < n; j += 256) {
        const long long t = idx[(size_t)b * k + j];
        const int cell = (int)((t / gw) / ph) * kwid + (int)((t % gw) / pw);
        atomicOr(&flags[cell >> 5], 1u << (cell & 31));
    }
    __syncthreads();
    // each thread owns a contiguous range of words; exclusive scan of the popcounts
    const int per = (words + 255) / 256;
    const int w0 = tid * per, w1 = min(words, w0 + per);
    unsigned int mine = 0;
    for (int w = w0; w < w1; ++w) mine += __popc(flags[w]);
    partial[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        unsigned int run = 0;
        for (int i = 0; i < 256; ++i) { const unsigned int c = partial[i]; partial[i] = run; run += c; }
        out_count[b] = (int)run;
    }
    __syncthreads();
    unsigned int pos = partial[tid];
    for (int w = w0; w < w1; ++w) {
        unsigned int bits = flags[w];
        while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            out_idx[(size_t)b * k + pos++] = (long long)w * 32 + bit;
        }
    }
}

int stats_smem(int kh, int kw) { return (2 * TD * LDT + TQ * (kh + 1 + kw + 1)) * (int)sizeof(float); }
int apply_smem(int mode, int kh, int kw) {
    return ((mode == ET_ATTN_DELTA ? 6 : 4) * TD * LDT + TQ * (kh + 1 + kw + 1)) * (int)sizeof(float);
}

int launch_generic(const GenArgs& a, cudaStream_t s) {
    const int nwin = a.windowed ? a.nwx * a.nwy : 1;
    const dim3 grid((a.Wn + TQ - 1) / TQ, a.H, a.B * nwin);
    const int kh = a.rel_y ? a.kh : 0, kw = a.rel_y ? a.kw : 0;
    int rc;
    const int sa = stats_smem(kh, kw);
    if ((rc = et_raise_smem(gen_stats_kernel, sa))) return rc;
    et_launch(gen_stats_kernel, grid, dim3(ATHREADS), sa, s, a);
    ET_COUNT_LAUNCH(1);
    if (a.stats_only) return ET_OK;
    const int sb = apply_smem(a.mode, kh, kw);
    if (a.mode == ET_ATTN_DELTA) {
        if ((rc = et_raise_smem(gen_apply_kernel<ET_ATTN_DELTA>, sb))) return rc;
        et_launch(gen_apply_kernel<ET_ATTN_DELTA>, grid, dim3(ATHREADS), sb, s, a);
    } else if (a.mode == ET_ATTN_FIRST) {
        if ((rc = et_raise_smem(gen_apply_kernel<ET_ATTN_FIRST>, sb))) return rc;
        et_launch(gen_apply_kernel<ET_ATTN_FIRST>, grid, dim3(ATHREADS), sb, s, a);
    } else {
        if ((rc = et_raise_smem(gen_apply_kernel<ET_ATTN_DENSE>, sb))) return rc;
        et_launch(gen_apply_kernel<ET_ATTN_DENSE>, grid, dim3(ATHREADS), sb, s, a);
    }
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}

int grid_1d(long long total) {
    const long long blocks = (total + 255) / 256, cap = (long long)et_sm_count() * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

// ---------------------------------------------------------------- entry points used by et_gemm.cu / et_attn.cu
int et_generic_linear(const void* A, int64_t M, int64_t K, const void* W, const void* bias, int64_t n_feat, int act, void* out,
                      int64_t ld_out, const int64_t* idx, const int32_t* count, int64_t k, int64_t n_out_rows,
                      cudaStream_t stream) {
    ET_CHECK_ARG(K % 4 == 0, "et_linear (fp32): K = %lld must be a multiple of 4", (long long)K);
    SgemmArgs a;
    a.A = static_cast<const float*>(A); a.W = static_cast<const float*>(W); a.bias = static_cast<const float*>(bias);
    a.out = static_cast<float*>(out); a.idx = reinterpret_cast<const long long*>(idx); a.count = count; a.ld_out = ld_out;
    a.M = (int)M; a.K = (int)K; a.NF = (int)n_feat; a.act = act; a.k = (int)(idx ? k : 1); a.n_out_rows = (int)n_out_rows;
    const dim3 grid((unsigned)((n_feat + GN - 1) / GN), (unsigned)((M + GM - 1) / GM));
    ET_CHECK_ARG(grid.y <= 65535, "et_linear (fp32): M too large");
    et_launch(sgemm_kernel, grid, dim3(GTHREADS), 0, stream, a);
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}

// Window attention in fp32 arithmetic (dense, two passes over the window's keys).
int et_generic_window_attention(const void* qkv, const void* pad_token, const void* rel_y, const void* rel_x, void* out,
                                float* stats, int B, int N, int gh, int gw, int wh, int ww, int H, int dh, int dtype,
                                cudaStream_t s) {
    GenArgs a = {};
    const int D = H * dh;
    a.q = qkv; a.ldq = 3LL * D; a.ldkv = 3LL * D; a.pad = pad_token; a.rel_y = rel_y; a.rel_x = rel_x;
    a.kk = static_cast<const char*>(qkv) + (size_t)D * elem_size(dtype);
    a.vv = static_cast<const char*>(qkv) + (size_t)2 * D * elem_size(dtype);
    a.B = B; a.Nq = N; a.Nk = N; a.H = H; a.dh = dh; a.D = D; a.gh = gh; a.gw = gw; a.wh = wh; a.ww = ww;
    a.windowed = wh > 0;
    if (a.windowed) {
        a.nwy = (gh + wh - 1) / wh; a.nwx = (gw + ww - 1) / ww; a.Wn = wh * ww; a.aw = ww; a.kh = wh; a.kw = ww;
    } else {
        a.nwx = a.nwy = 1; a.Wn = N; a.aw = gw; a.kh = gh; a.kw = gw;
    }
    a.rscale = 1.0f / sqrtf((float)dh);
    a.mode = ET_ATTN_DENSE; a.out = out; a.stats = stats; a.dt = dtype; a.sdt = dtype; a.NP = 0; a.q_batch_rows = N;
    return launch_generic(a, s);
}

// Global attention (DENSE / FIRST / DELTA) with optional pooled K/V, separate state dtype and device-side counts.
// `q_index` != null (adaptive token sampling): the Nq query rows are the tokens q_index[b][0 .. Nq) of qkv instead of all N;
// keys and values are still the N tokens of qkv, the state / accumulator / output have Nq rows.
static int generic_global_impl(const void* qkv, const void* kv_pooled, int pool_h, int pool_w, const void* rel_y,
                               const void* rel_x, int mode, const int64_t* idx, const int32_t* count, int k, void* a_state,
                               void* v_state, void* acc, void* out, float* stats, void* ws, int B, int N, int gh, int gw,
                               int H, int dh, int dtype, int state_dtype, cudaStream_t s, const long long* q_index, int Nq) {
    GenArgs a = {};
    const int D = H * dh;
    a.q = qkv; a.ldq = 3LL * D; a.rel_y = rel_y; a.rel_x = rel_x;
    a.B = B; a.Nq = N; a.H = H; a.dh = dh; a.D = D; a.gh = gh; a.gw = gw; a.windowed = 0; a.nwx = a.nwy = 1; a.Wn = N; a.aw = gw;
    a.q_batch_rows = N;
    if (q_index != nullptr) { a.q_index = q_index; a.Nq = Nq; a.Wn = Nq; }
    if (kv_pooled != nullptr) {
        a.kk = kv_pooled;
        a.vv = static_cast<const char*>(kv_pooled) + (size_t)D * elem_size(dtype);
        a.ldkv = 2LL * D; a.kh = gh / pool_h; a.kw = gw / pool_w; a.Nk = a.kh * a.kw;
    } else {
        a.kk = static_cast<const char*>(qkv) + (size_t)D * elem_size(dtype);
        a.vv = static_cast<const char*>(qkv) + (size_t)2 * D * elem_size(dtype);
        a.ldkv = 3LL * D; a.kh = gh; a.kw = gw; a.Nk = N;
    }
    a.rscale = 1.0f / sqrtf((float)dh);
    a.mode = mode; a.idx = reinterpret_cast<const long long*>(idx); a.count = count; a.kmax = mode == ET_ATTN_DELTA ? k : 0;
    a.a_state = a_state; a.NP = (a.Nq + 7) / 8 * 8; a.acc = acc; a.out = out; a.stats = stats; a.dt = dtype; a.sdt = state_dtype;
    if (mode == ET_ATTN_DELTA) {
        if (k == 0) return ET_OK;
        char* w = static_cast<char*>(ws);
        void* dV = w;
        void* Vr = w + (((size_t)B * k * D * elem_size(state_dtype)) + 255) / 256 * 256;
        const long long total = (long long)B * k * D;
        et_launch(gen_vgate_kernel, dim3(grid_1d(total)), dim3(256), 0, s, a.vv, a.ldkv, v_state, a.idx, count, dV, Vr, a.Nk, D, k,
                  total, dtype, state_dtype);
        ET_COUNT_LAUNCH(1);
        a.dV = dV; a.Vr = Vr;
    } else if (mode == ET_ATTN_FIRST) {
        const long long total = (long long)B * a.Nk * D;
        et_launch(gen_vgate_kernel, dim3(grid_1d(total)), dim3(256), 0, s, a.vv, a.ldkv, v_state, (const long long*)nullptr,
                  (const int*)nullptr, (void*)nullptr, (void*)nullptr, a.Nk, D, a.Nk, total, dtype, state_dtype);
        ET_COUNT_LAUNCH(1);
    }
    return launch_generic(a, s);
}

int et_generic_global_attention(const void* qkv, const void* kv_pooled, int pool_h, int pool_w, const void* rel_y,
                                const void* rel_x, int mode, const int64_t* idx, const int32_t* count, int k, void* a_state,
                                void* v_state, void* acc, void* out, float* stats, void* ws, int B, int N, int gh, int gw,
                                int H, int dh, int dtype, int state_dtype, cudaStream_t s) {
    return generic_global_impl(qkv, kv_pooled, pool_h, pool_w, rel_y, rel_x, mode, idx, count, k, a_state, v_state, acc, out, stats,
                               ws, B, N, gh, gw, H, dh, dtype, state_dtype, s, nullptr, 0);
}

extern "C" {

int et_ats_scores(const void* qkv, int64_t B, int64_t N, int64_t heads, int64_t dh, int dtype, int score_dtype, float* row_stats,
                  float* raw_scores, void* stream) {
    ET_CHECK_ARG(qkv && row_stats && raw_scores, "et_ats_scores: null pointer");
    ET_CHECK_ARG(B > 0 && N > 0 && heads > 0 && dh > 0 && dh <= 64, "et_ats_scores: bad shape (head dim <= 64)");
    ET_CHECK_ARG((dtype == ET_F32 || dtype == ET_BF16 || dtype == ET_F16) &&
                     (score_dtype == ET_F32 || score_dtype == ET_BF16 || score_dtype == ET_F16), "et_ats_scores: bad dtype");
    GenArgs a = {};
    const int D = (int)(heads * dh);
    a.q = qkv; a.ldq = 3LL * D;
    a.kk = static_cast<const char*>(qkv) + (size_t)D * elem_size(dtype);
    a.vv = static_cast<const char*>(qkv) + (size_t)2 * D * elem_size(dtype);
    a.ldkv = 3LL * D;
    a.B = (int)B; a.Nq = a.Nk = a.Wn = (int)N; a.H = (int)heads; a.dh = (int)dh; a.D = D; a.nwx = a.nwy = 1;
    a.rscale = 1.0f / sqrtf((float)dh);
    a.mode = ET_ATTN_DENSE; a.stats = row_stats; a.dt = dtype; a.sdt = score_dtype; a.q_batch_rows = (int)N;
    a.ats_raw = raw_scores; a.ats_dt = score_dtype; a.stats_only = 1;
    int rc = launch_generic(a, et_stream(stream));
    if (rc) return rc;
    ET_CHECK_LAUNCH("et_ats_scores");
    return ET_OK;
}

int et_global_attention_rows(const void* qkv, const int64_t* q_index, int64_t Nq, int mode, const int64_t* idx,
                             const int32_t* count, int64_t k, void* a_state, void* v_state, void* acc, void* out,
                             float* row_stats, void* workspace, int64_t B, int64_t N, int64_t heads, int64_t dh, int dtype,
                             int state_dtype, void* stream) {
    ET_CHECK_ARG(q_index && qkv && out && row_stats, "et_global_attention_rows: null pointer");
    ET_CHECK_ARG((dtype == ET_F32 || dtype == ET_BF16 || dtype == ET_F16) &&
                     (state_dtype == ET_F32 || state_dtype == ET_BF16 || state_dtype == ET_F16), "et_global_attention_rows: bad dtype");
    ET_CHECK_ARG(mode == ET_ATTN_DENSE || mode == ET_ATTN_FIRST || mode == ET_ATTN_DELTA, "et_global_attention_rows: bad mode");
    ET_CHECK_ARG(B > 0 && N > 0 && Nq > 0 && Nq <= N && heads > 0 && dh > 0 && dh <= 64,
                 "et_global_attention_rows: bad shape (Nq <= N, head dim <= 64)");
    ET_CHECK_ARG(mode == ET_ATTN_DENSE || (a_state && v_state && acc), "et_global_attention_rows: state pointers required");
    ET_CHECK_ARG(mode != ET_ATTN_DELTA || (idx != nullptr && k >= 0 && k <= N && workspace != nullptr),
                 "et_global_attention_rows: DELTA needs idx, k <= N and the workspace");
    int rc = generic_global_impl(qkv, nullptr, 1, 1, nullptr, nullptr, mode, idx, count, (int)k, a_state, v_state, acc, out,
                                 row_stats, workspace, (int)B, (int)N, 0, 0, (int)heads, (int)dh, dtype, state_dtype,
                                 et_stream(stream), reinterpret_cast<const long long*>(q_index), (int)Nq);
    if (rc) return rc;
    ET_CHECK_LAUNCH("et_global_attention_rows");
    return ET_OK;
}

int et_pool_kv(const void* qkv, void* out, int64_t B, int64_t gh, int64_t gw, int64_t D, int64_t pool_h, int64_t pool_w,
               int dtype, void* stream) {
    ET_CHECK_ARG(qkv && out, "et_pool_kv: null pointer");
    ET_CHECK_ARG(pool_h > 0 && pool_w > 0 && gh % pool_h == 0 && gw % pool_w == 0, "et_pool_kv: grid %lld x %lld not divisible by pool %lld x %lld",
                 (long long)gh, (long long)gw, (long long)pool_h, (long long)pool_w);
    ET_CHECK_ARG(dtype == ET_F32 || dtype == ET_BF16 || dtype == ET_F16, "et_pool_kv: bad dtype");
    const long long total = (long long)B * (gh / pool_h) * (gw / pool_w) * 2 * D;
    if (total == 0) return ET_OK;
    et_launch(pool_kv_kernel, dim3(grid_1d(total)), dim3(256), 0, et_stream(stream), qkv, out, (int)gh, (int)gw, (int)pool_h, (int)pool_w,
              (int)D, total, dtype);
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_pool_kv");
    return ET_OK;
}

int et_pool_index(const int64_t* idx, const int32_t* count_in, int64_t B, int64_t k, int64_t gh, int64_t gw, int64_t pool_h,
                  int64_t pool_w, int64_t* out_idx, int32_t* out_count, void* stream) {
    ET_CHECK_ARG(idx && out_idx && out_count, "et_pool_index: null pointer");
    ET_CHECK_ARG(pool_h > 0 && pool_w > 0 && gh % pool_h == 0 && gw % pool_w == 0, "et_pool_index: grid not divisible by the pool size");
    if (B == 0) return ET_OK;
    const int ncells = (int)((gh / pool_h) * (gw / pool_w));
    const int smem = ((ncells + 31) / 32 + 256) * (int)sizeof(unsigned int);
    int rc = et_raise_smem(pool_index_kernel, smem);
    if (rc) return rc;
    et_launch(pool_index_kernel, dim3((unsigned)B), dim3(256), smem, et_stream(stream), reinterpret_cast<const long long*>(idx), count_in,
              (int)k, (int)gw, (int)pool_h, (int)pool_w, ncells, reinterpret_cast<long long*>(out_idx), out_count);
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_pool_index");
    return ET_OK;
}

}  // extern "C"
