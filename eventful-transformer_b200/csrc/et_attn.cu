// Attention kernels of the gated update path (bf16 / fp16, fp32 accumulate).
//
//   relpos_bias_kernel : decomposed relative-position bias tables from the UNSCALED q
//                        (eventful_transformer/utils.py:139-171), rounded to dtype like the einsum outputs
//   window_attention   : dense windowed / small-global attention, single pass, online softmax
//                        (Block._forward_attention, blocks.py:205-240 incl. window partition/recombine)
//   attn_stats         : phase A of global attention: row max and row sum of softmax(q k^T / s + bias)
//   attn_apply         : phase B: DENSE / FIRST  : out = a . v (and state initialisation)
//                                 DELTA          : A-gate + MatmulDeltaAccumulator on the selected columns
//   vgate / vstate     : TokenDeltaGate on v rows (forced index), modules.py:187-201
//
// Tensor-core work here uses mma.sync.m16n8k16 (legacy HMMA path): round 1 establishes the fused
// algorithm and its parity; the tcgen05/TMEM version of attn_apply is the next optimisation step.
#include <type_traits>

#include "et_common.cuh"

int et_tc_global_attention(const void* qkv, const void* sel, void* onehot, void* onehot_all, const void* bias_h, const void* bias_w, int mode,
                           const long long* idx, int k, void* a_state, void* acc, void* out, float* stats, int B, int N,
                           int NP, int H, int gh, int gw, int is_bf16, cudaStream_t stream);
int et_tc_window_attention(const void* qkv, const void* pad_token, void* bias_comb, void* out, int B, int N, int gh, int gw,
                           int wh, int ww, int H, int has_bias, int is_bf16, cudaStream_t s);
int et_generic_window_attention(const void* qkv, const void* pad_token, const void* rel_y, const void* rel_x, void* out,
                                float* stats, int B, int N, int gh, int gw, int wh, int ww, int H, int dh, int dtype,
                                cudaStream_t s);
int et_generic_global_attention(const void* qkv, const void* kv_pooled, int pool_h, int pool_w, const void* rel_y,
                                const void* rel_x, int mode, const int64_t* idx, const int32_t* count, int k, void* a_state,
                                void* v_state, void* acc, void* out, float* stats, void* ws, int B, int N, int gh, int gw,
                                int H, int dh, int dtype, int state_dtype, cudaStream_t s);
bool et_tc_window2_applies(int wh, int ww, int dh);
int et_tc_window2_attention(const void* qkv, const void* pad_token, const void* rel_y, const void* rel_x, void* out, int B, int N,
                            int gh, int gw, int wh, int ww, int H, int is_bf16, cudaStream_t s);
int g_attn_win_gen = 2;  // et_debug_set(11, 1) selects the first-generation tcgen05 window kernel (tests compare the two)
int g_attn_tc = 1;  // et_debug_set(2, 0) forces the mma.sync kernels (tests compare the two paths)

namespace {

constexpr int kAttnThreads = 128;  // 4 warps x 16 query rows
constexpr int BQ = 64;             // query rows per CTA
constexpr int BKV = 64;            // keys per tile
constexpr float kLog2e = 1.4426950408889634f;

struct AttnArgs {
    const void* qkv;
    const void* pad_token;
    const void* bias_h;
    const void* bias_w;
    void* out;
    int B, N, NP, gh, gw, wh, ww, H, nwx, nwy, Wn, windowed;  // NP = N rounded up to 8 (a_state row stride)
    float rscale;
    // eventful part
    const long long* idx;
    const int* count;  // device-side number of valid entries of idx per batch entry (or null: all k)
    int k, mode;
    void* a_state;
    void* acc;
    const void* dV;
    const void* Vd;
    float* stats;
};

// ---------------------------------------------------------------- small PTX helpers
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_addr(p)));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (std::is_same_v<T, __nv_bfloat16>) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    T v[2] = {ElemTraits<T>::from_float(lo), ElemTraits<T>::from_float(hi)};
    return *reinterpret_cast<uint32_t*>(v);
}
template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return ElemTraits<T>::to_float(*p); }

// ---------------------------------------------------------------- token mapping
// Local token t of window `win` -> row of the (B, N, 3D) QKV buffer, or -1 for a padding token.
struct TokenMap {
    int N, gw, gh, wh, ww, nwx, windowed, extra;
    __device__ __forceinline__ int token(int win, int t) const {
        if (!windowed) return t;
        const int wy = win / nwx, wx = win - wy * nwx;
        const int ly = t / ww, lx = t - ly * ww;
        const int gy = wy * wh + ly, gx = wx * ww + lx;
        return (gy < gh && gx < gw) ? gy * gw + gx : -1;
    }
};
__device__ __forceinline__ TokenMap make_map(const AttnArgs& a) {
    TokenMap m;
    m.N = a.N; m.gw = a.gw; m.gh = a.gh; m.wh = a.wh; m.ww = a.ww; m.nwx = a.nwx; m.windowed = a.windowed; m.extra = 0;
    return m;
}

// Stage `rows` rows of DH elements into smem (row stride LD), one 16-byte cp.async per chunk.
// row_ptr(i) returns the global pointer of row i or nullptr (zero fill).
template <typename T, int DH, int LD, typename F>
__device__ __forceinline__ void stage_rows(T* dst, int rows, F row_ptr) {
    constexpr int CH = DH / 8;
    for (int c = threadIdx.x; c < rows * CH; c += kAttnThreads) {
        const int r = c / CH, ch = c - r * CH;
        const T* src = row_ptr(r);
        T* d = dst + r * LD + ch * 8;
        if (src != nullptr) cp_async16(d, src + ch * 8);
        else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
    }
}

// S[16 x 64] (per warp) = Q[16 x DH] . K[64 x DH]^T ; Q fragments preloaded.
template <typename T, int DH, int LD>
__device__ __forceinline__ void qk_tile(float (&s)[8][4], const uint32_t (&qf)[DH / 16][4], const T* Ks, int lane) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
            uint32_t b[4];
            const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
            const int col = ks * 16 + ((lane >> 3) & 1) * 8;
            ldsm_x4(b, Ks + row * LD + col);
            mma16816<T>(s[np * 2], qf[ks], b[0], b[1]);
            mma16816<T>(s[np * 2 + 1], qf[ks], b[2], b[3]);
        }
    }
}

// O[16 x DH] += P[16 x 64] . V[64 x DH] ; P given as A fragments (4 k-steps).
template <typename T, int DH, int LD>
__device__ __forceinline__ void pv_tile(float (&o)[DH / 8][4], const uint32_t (&pf)[4][4], const T* Vs, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < DH / 16; ++np) {
            uint32_t b[4];
            const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int col = np * 16 + (lane >> 4) * 8;
            ldsm_x4_trans(b, Vs + row * LD + col);
            mma16816<T>(o[np * 2], pf[ks], b[0], b[1]);
            mma16816<T>(o[np * 2 + 1], pf[ks], b[2], b[3]);
        }
    }
}

template <typename T, int DH, int LD>
__device__ __forceinline__ void load_q_frags(uint32_t (&qf)[DH / 16][4], const T* Qs, int warp, int lane) {
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = ks * 16 + (lane >> 4) * 8;
        ldsm_x4(qf[ks], Qs + row * LD + col);
    }
}

// Adds the decomposed rel-pos bias to a score tile.  bias rows live in smem: bh[r][ky], bw[r][kx].
template <typename T>
__device__ __forceinline__ void add_bias(float (&s)[8][4], const T* bh, const T* bw, int ldh, int ldw, int row_lo,
                                         int key0, int kw, int lane) {
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row_lo + g + (i >> 1) * 8;
            const int key = key0 + nt * 8 + tq * 2 + (i & 1);
            const int ky = key / kw, kx = key - ky * kw;
            s[nt][i] += ldf(bh + r * ldh + ky) + ldf(bw + r * ldw + kx);
        }
    }
}

// ---------------------------------------------------------------- rel-pos bias tables
// grid (wh + ww, H, B * n_windows); y-lines compute bias_h, x-lines compute bias_w.
// One line = all tokens sharing a query coordinate, so they share one (coords x DH) relative table:
// bias[token][coord] = q[token] . table[coord]  -> a small tensor-core GEMM per line (mma.sync m16n8k16).
template <typename T, int DH>
__global__ void __launch_bounds__(kAttnThreads) relpos_bias_kernel(const AttnArgs a, const T* rel_y, const T* rel_x,
                                                                    T* bias_h, T* bias_w, int ld_pad, float scale,
                                                                    int rows_pad, int combined) {
    et_pdl_prologue();
    constexpr int LD = DH + 8;
    __shared__ __align__(16) T Qs[BQ * LD];
    __shared__ __align__(16) T Ts[BKV * LD];
    __shared__ __align__(16) T Cs[BQ * LD];  // output tile staging (16-bit dtypes)
    const TokenMap map = make_map(a);
    const int line = blockIdx.x, h = blockIdx.y, bw = blockIdx.z;
    const int nwin = a.windowed ? a.nwx * a.nwy : 1;
    const int b = bw / nwin, win = bw - b * nwin;
    const int lh = a.windowed ? a.wh : a.gh, lw = a.windowed ? a.ww : a.gw;
    const bool ymode = line < lh;
    const int fixed = ymode ? line : line - lh;      // ly or lx
    const int ntok = ymode ? lw : lh;                // tokens on this line
    const int nout = ymode ? lh : lw;                // key coordinates
    const T* table = (ymode ? rel_y + (size_t)fixed * lh * DH : rel_x + (size_t)fixed * lw * DH);
    const int D = a.H * DH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    T* dst = ymode ? bias_h : bias_w;
    for (int tok0 = 0; tok0 < ntok; tok0 += BQ) {
        __syncthreads();
        stage_rows<T, DH, LD>(Qs, BQ, [&](int r) -> const T* {
            const int j = tok0 + r;
            if (j >= ntok) return nullptr;
            const int tok = map.token(win, ymode ? fixed * lw + j : j * lw + fixed);
            return tok >= 0 ? static_cast<const T*>(a.qkv) + ((size_t)b * a.N + tok) * 3 * D + h * DH
                            : static_cast<const T*>(a.pad_token) + h * DH;
        });
        for (int out0 = 0; out0 < nout; out0 += BKV) {
            __syncthreads();
            stage_rows<T, DH, LD>(Ts, BKV, [&](int r) -> const T* {
                return out0 + r < nout ? table + (size_t)(out0 + r) * DH : nullptr;
            });
            cp_async_wait_all();
            __syncthreads();
            uint32_t qf[DH / 16][4];
            load_q_frags<T, DH, LD>(qf, Qs, warp, lane);
            float s[8][4];
            qk_tile<T, DH, LD>(s, qf, Ts, lane);
            // ld_pad: tensor-core layouts with padded rows. combined: one array [bias_h | bias_w | 0] per token
            // (pre-zeroed by the caller, bias_w == bias_h); otherwise the pad columns are zero-filled here.
            const int ld = ld_pad > 0 ? ld_pad : nout;
            const int rows = rows_pad > 0 ? rows_pad : a.Wn;
            const int coff = (combined && !ymode) ? lh : 0;
            const int limit = combined ? nout : ld;
            if constexpr (sizeof(T) == 2) {
                if (((ld | coff) & 7) == 0) {
                    // 16-bit tables with 16-byte aligned rows: the 64 x 64 tile is staged in shared memory and written as
                    // 16-byte chunks (a token's 64 coordinates are one 128-byte row) instead of one 2-byte store per element
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                        for (int i = 0; i < 4; i += 2) {
                            const int r = warp * 16 + g + (i >> 1) * 8;
                            const int c = nt * 8 + tq * 2;
                            const float v0 = out0 + c < nout ? s[nt][i] * scale : 0.f;
                            const float v1 = out0 + c + 1 < nout ? s[nt][i + 1] * scale : 0.f;
                            *reinterpret_cast<uint32_t*>(Cs + r * LD + c) = pack2_rn<T>(v0, v1);
                        }
                    __syncthreads();
                    for (int ch = threadIdx.x; ch < BQ * (BKV / 8); ch += kAttnThreads) {
                        const int r = ch >> 3, c8 = (ch & 7) * 8;
                        const int j = tok0 + r, kc0 = out0 + c8;
                        if (j < ntok && kc0 < limit) {
                            const int t = ymode ? fixed * lw + j : j * lw + fixed;
                            T* row = dst + (((size_t)bw * a.H + h) * rows + t) * ld + coff + kc0;
                            if (kc0 + 8 <= limit) {
                                st16(row, ld16(Cs + r * LD + c8));
                            } else {
                                for (int e = 0; kc0 + e < limit; ++e) row[e] = Cs[r * LD + c8 + e];
                            }
                        }
                    }
                    continue;  // the next iteration starts with __syncthreads() before Ts / Cs are rewritten
                }
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = tok0 + warp * 16 + g + (i >> 1) * 8;
                    const int kc = out0 + nt * 8 + tq * 2 + (i & 1);
                    if (j < ntok && kc < limit) {
                        const int t = ymode ? fixed * lw + j : j * lw + fixed;
                        dst[(((size_t)bw * a.H + h) * rows + t) * ld + coff + kc] =
                            ElemTraits<T>::from_float(kc < nout ? s[nt][i] * scale : 0.f);
                    }
                }
        }
    }
}

// ---------------------------------------------------------------- rel-pos bias, windowed blocks, one CTA per (window, head)
// Stages the window's q rows and both relative tables in shared memory once, then each warp walks lines
// (all tokens sharing ly, or sharing lx): a 16 x 16 x DH product per line with ldmatrix row gathers.
// Output: compact tensor-core layout (Bw, H, 256, 32): columns [0, wh) = 8 bias_h, [wh, wh + ww) = 8 bias_w, zeros
// elsewhere (wh + ww <= 32).  The tile is assembled in shared memory and written as one contiguous 16 KB block, so the
// scratch needs no memset and the attention kernel reads 64 bytes per query row.
template <typename T, int DH>
__global__ void __launch_bounds__(kAttnThreads) relpos_window_kernel(const AttnArgs a, const T* rel_y, const T* rel_x, T* comb,
                                                                      float scale) {
    et_pdl_prologue();
    constexpr int LD = DH + 8;
    extern __shared__ __align__(16) uint8_t smraw[];
    const int wh = a.wh, ww = a.ww, Wn = a.Wn;
    T* Qs = reinterpret_cast<T*>(smraw);            // [Wn + 1][LD], last row = zeros
    T* Ty = Qs + (Wn + 1) * LD;                     // [wh * wh + 1][LD]
    T* Tx = Ty + (wh * wh + 1) * LD;                // [ww * ww + 1][LD]
    T* Cs = Tx + (ww * ww + 1) * LD;                // [256][32] output tile
    for (int c = threadIdx.x; c < 256 * 32 * (int)sizeof(T) / 16; c += kAttnThreads)
        reinterpret_cast<uint4*>(Cs)[c] = make_uint4(0, 0, 0, 0);
    const TokenMap map = make_map(a);
    const int bw = blockIdx.x, h = blockIdx.y;
    const int nwin = a.nwx * a.nwy;
    const int b = bw / nwin, win = bw - b * nwin;
    const int D = a.H * DH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    stage_rows<T, DH, LD>(Qs, Wn + 1, [&](int r) -> const T* {
        if (r >= Wn) return nullptr;
        const int tok = map.token(win, r);
        return tok >= 0 ? static_cast<const T*>(a.qkv) + ((size_t)b * a.N + tok) * 3 * D + h * DH
                        : static_cast<const T*>(a.pad_token) + h * DH;
    });
    stage_rows<T, DH, LD>(Ty, wh * wh + 1, [&](int r) -> const T* { return r < wh * wh ? rel_y + (size_t)r * DH : nullptr; });
    stage_rows<T, DH, LD>(Tx, ww * ww + 1, [&](int r) -> const T* { return r < ww * ww ? rel_x + (size_t)r * DH : nullptr; });
    cp_async_wait_all();
    __syncthreads();
    for (int line = warp; line < wh + ww; line += kAttnThreads / 32) {
        const bool ymode = line < wh;
        const int fixed = ymode ? line : line - wh;
        const int ntok = ymode ? ww : wh, nout = ymode ? wh : ww;
        const T* tab = ymode ? Ty : Tx;
        const int tab_rows = ymode ? wh * wh : ww * ww;
        for (int j0 = 0; j0 < ntok; j0 += 16) {
            for (int o0 = 0; o0 < nout; o0 += 16) {
                float acc[2][4] = {};
#pragma unroll
                for (int ks = 0; ks < DH / 16; ++ks) {
                    uint32_t af[4], bf[4];
                    const int j = j0 + (lane & 15);
                    const int t = ymode ? fixed * ww + j : j * ww + fixed;
                    ldsm_x4(af, Qs + (j < ntok ? t : Wn) * LD + ks * 16 + (lane >> 4) * 8);
                    const int kc = o0 + (lane & 7) + (lane >> 4) * 8;
                    ldsm_x4(bf, tab + (kc < nout ? fixed * nout + kc : tab_rows) * LD + ks * 16 + ((lane >> 3) & 1) * 8);
                    mma16816<T>(acc[0], af, bf[0], bf[1]);
                    mma16816<T>(acc[1], af, bf[2], bf[3]);
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int j = j0 + g + (i >> 1) * 8;
                        const int kc = o0 + nt * 8 + tq * 2 + (i & 1);
                        if (j < ntok && kc < nout) {
                            const int t = ymode ? fixed * ww + j : j * ww + fixed;
                            Cs[t * 32 + (ymode ? 0 : wh) + kc] = ElemTraits<T>::from_float(acc[nt][i] * scale);
                        }
                    }
            }
        }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(comb + ((size_t)bw * a.H + h) * 256 * 32);
    for (int c = threadIdx.x; c < 256 * 32 * (int)sizeof(T) / 16; c += kAttnThreads) dst[c] = reinterpret_cast<const uint4*>(Cs)[c];
}

// ---------------------------------------------------------------- windowed / small dense attention
// grid (ceil(Wn / 64), H, B * n_windows)
template <typename T, int DH>
__global__ void __launch_bounds__(kAttnThreads) window_attention_kernel(const AttnArgs a) {
    et_pdl_prologue();
    constexpr int LD = DH + 8;
    extern __shared__ __align__(16) uint8_t smraw[];
    T* Qs = reinterpret_cast<T*>(smraw);
    T* Ks = Qs + BQ * LD;
    T* Vs = Ks + BKV * LD;
    T* bh = Vs + BKV * LD;
    const int lh = a.windowed ? a.wh : a.gh, lw = a.windowed ? a.ww : a.gw;
    const int ldh = lh + 1, ldw = lw + 1;
    T* bwS = bh + BQ * ldh;
    const bool has_bias = a.bias_h != nullptr;

    const TokenMap map = make_map(a);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int qb = blockIdx.x, h = blockIdx.y, bwi = blockIdx.z;
    const int nwin = a.windowed ? a.nwx * a.nwy : 1;
    const int b = bwi / nwin, win = bwi - b * nwin;
    const int D = a.H * DH;
    const T* qkv = static_cast<const T*>(a.qkv);
    const T* pad = static_cast<const T*>(a.pad_token);
    const int q0 = qb * BQ;

    auto row_ptr = [&](int t, int part) -> const T* {
        if (t >= a.Wn) return nullptr;
        const int tok = map.token(win, t);
        if (tok < 0) return pad + part * D + h * DH;
        return qkv + ((size_t)b * a.N + tok) * 3 * D + part * D + h * DH;
    };
    stage_rows<T, DH, LD>(Qs, BQ, [&](int r) { return row_ptr(q0 + r, 0); });
    if (has_bias) {
        const T* gh_ = static_cast<const T*>(a.bias_h) + (((size_t)bwi * a.H + h) * a.Wn) * lh;
        const T* gw_ = static_cast<const T*>(a.bias_w) + (((size_t)bwi * a.H + h) * a.Wn) * lw;
        for (int i = threadIdx.x; i < BQ * lh; i += kAttnThreads) {
            const int r = i / lh, c = i - r * lh;
            bh[r * ldh + c] = (q0 + r < a.Wn) ? gh_[(size_t)(q0 + r) * lh + c] : ElemTraits<T>::from_float(0.f);
        }
        for (int i = threadIdx.x; i < BQ * lw; i += kAttnThreads) {
            const int r = i / lw, c = i - r * lw;
            bwS[r * ldw + c] = (q0 + r < a.Wn) ? gw_[(size_t)(q0 + r) * lw + c] : ElemTraits<T>::from_float(0.f);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    uint32_t qf[DH / 16][4];
    load_q_frags<T, DH, LD>(qf, Qs, warp, lane);

    float o[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

    for (int key0 = 0; key0 < a.Wn; key0 += BKV) {
        __syncthreads();  // previous tile fully consumed
        stage_rows<T, DH, LD>(Ks, BKV, [&](int r) { return row_ptr(key0 + r, 1); });
        stage_rows<T, DH, LD>(Vs, BKV, [&](int r) { return row_ptr(key0 + r, 2); });
        cp_async_wait_all();
        __syncthreads();
        float s[8][4];
        qk_tile<T, DH, LD>(s, qf, Ks, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) s[nt][i] *= a.rscale;
        if (has_bias) add_bias<T>(s, bh, bwS, ldh, ldw, warp * 16, key0, lw, lane);
        // mask keys beyond the window, online softmax
        float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int key = key0 + nt * 8 + tq * 2 + (i & 1);
                if (key >= a.Wn) s[nt][i] = -INFINITY;
                mx[i >> 1] = fmaxf(mx[i >> 1], s[nt][i]);
            }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
            mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        }
        float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            corr[hh] = exp2f((mrow[hh] - mx[hh]) * kLog2e);
            mrow[hh] = mx[hh];
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                p[i] = exp2f((s[nt][i] - mx[i >> 1]) * kLog2e);
                rs[i >> 1] += p[i];
            }
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack2<T>(p[0], p[1]);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack2<T>(p[2], p[3]);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 1);
            rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 2);
            lrow[hh] = lrow[hh] * corr[hh] + rs[hh];
        }
#pragma unroll
        for (int i = 0; i < DH / 8; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0];
            o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        pv_tile<T, DH, LD>(o, pf, Vs, lane);
    }
    // normalise and write out[b, token, h*DH + c] (padding rows are cropped, blocks.py:371-373)
    T* out = static_cast<T*>(a.out);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int t = q0 + warp * 16 + g + hh * 8;
        if (t >= a.Wn) continue;
        const int tok = map.token(win, t);
        if (tok < 0) continue;
        const float inv = 1.f / lrow[hh];
        T* dst = out + ((size_t)b * a.N + tok) * D + h * DH;
#pragma unroll
        for (int nt = 0; nt < DH / 8; ++nt)
            *reinterpret_cast<uint32_t*>(dst + nt * 8 + tq * 2) = pack2<T>(o[nt][hh * 2] * inv, o[nt][hh * 2 + 1] * inv);
    }
}

// ---------------------------------------------------------------- global attention, phase A
// grid (N / 64 rounded up, H, B): row max / row sum of the softmax over ALL keys.
template <typename T, int DH>
__global__ void __launch_bounds__(kAttnThreads) attn_stats_kernel(const AttnArgs a) {
    et_pdl_prologue();
    constexpr int LD = DH + 8;
    extern __shared__ __align__(16) uint8_t smraw[];
    T* Qs = reinterpret_cast<T*>(smraw);
    T* Ks = Qs + BQ * LD;
    T* bh = Ks + BKV * LD;
    const int ldh = a.gh + 1, ldw = a.gw + 1;
    T* bwS = bh + BQ * ldh;
    const bool has_bias = a.bias_h != nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int D = a.H * DH;
    const T* qkv = static_cast<const T*>(a.qkv);
    auto row_ptr = [&](int t, int part) -> const T* {
        return t < a.N ? qkv + ((size_t)b * a.N + t) * 3 * D + part * D + h * DH : nullptr;
    };
    stage_rows<T, DH, LD>(Qs, BQ, [&](int r) { return row_ptr(q0 + r, 0); });
    if (has_bias) {
        const T* gh_ = static_cast<const T*>(a.bias_h) + (((size_t)b * a.H + h) * a.N) * a.gh;
        const T* gw_ = static_cast<const T*>(a.bias_w) + (((size_t)b * a.H + h) * a.N) * a.gw;
        for (int i = threadIdx.x; i < BQ * a.gh; i += kAttnThreads) {
            const int r = i / a.gh, c = i - r * a.gh;
            bh[r * ldh + c] = (q0 + r < a.N) ? gh_[(size_t)(q0 + r) * a.gh + c] : ElemTraits<T>::from_float(0.f);
        }
        for (int i = threadIdx.x; i < BQ * a.gw; i += kAttnThreads) {
            const int r = i / a.gw, c = i - r * a.gw;
            bwS[r * ldw + c] = (q0 + r < a.N) ? gw_[(size_t)(q0 + r) * a.gw + c] : ElemTraits<T>::from_float(0.f);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    uint32_t qf[DH / 16][4];
    load_q_frags<T, DH, LD>(qf, Qs, warp, lane);
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    for (int key0 = 0; key0 < a.N; key0 += BKV) {
        __syncthreads();
        stage_rows<T, DH, LD>(Ks, BKV, [&](int r) { return row_ptr(key0 + r, 1); });
        cp_async_wait_all();
        __syncthreads();
        float s[8][4];
        qk_tile<T, DH, LD>(s, qf, Ks, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) s[nt][i] *= a.rscale;
        if (has_bias) add_bias<T>(s, bh, bwS, ldh, ldw, warp * 16, key0, a.gw, lane);
        float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int key = key0 + nt * 8 + tq * 2 + (i & 1);
                if (key >= a.N) s[nt][i] = -INFINITY;
                mx[i >> 1] = fmaxf(mx[i >> 1], s[nt][i]);
            }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
            mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) rs[i >> 1] += exp2f((s[nt][i] - mx[i >> 1]) * kLog2e);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 1);
            rs[hh] += __shfl_xor_sync(0xffffffffu, rs[hh], 2);
            lrow[hh] = lrow[hh] * exp2f((mrow[hh] - mx[hh]) * kLog2e) + rs[hh];
            mrow[hh] = mx[hh];
        }
    }
    if (tq == 0) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int t = q0 + warp * 16 + g + hh * 8;
            if (t < a.N) {
                float* st = a.stats + (((size_t)b * a.H + h) * a.N + t) * 2;
                st[0] = mrow[hh];
                st[1] = lrow[hh];
            }
        }
    }
}

// ---------------------------------------------------------------- global attention, phase B
// grid (N / 64 rounded up, H, B).  Key tiles run over all tokens (DENSE / FIRST) or over the
// k gate-selected tokens (DELTA).  a_state is column-major per head: a_state[b][h][col][row].
template <typename T, int DH, int MODE>
__global__ void __launch_bounds__(kAttnThreads) attn_apply_kernel(const AttnArgs a) {
    et_pdl_prologue();
    constexpr int LD = DH + 8;
    constexpr int LDP = BQ + 8;
    extern __shared__ __align__(16) uint8_t smraw[];
    T* Qs = reinterpret_cast<T*>(smraw);
    T* Ks = Qs + BQ * LD;
    T* V1 = Ks + BKV * LD;   // V (DENSE / FIRST) or delta-V (DELTA)
    T* V2 = V1 + BKV * LD;   // v_n - delta-V (DELTA)
    T* Ps = V2 + BKV * LD;   // a_state tile, [key j][row r]
    T* bh = Ps + BKV * LDP;
    const int ldh = a.gh + 1, ldw = a.gw + 1;
    T* bwS = bh + BQ * ldh;
    __shared__ int s_tok[BKV];
    const bool has_bias = a.bias_h != nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int D = a.H * DH;
    const T* qkv = static_cast<const T*>(a.qkv);
    const int nkeys = (MODE == ET_ATTN_DELTA) ? (a.count != nullptr ? min(a.k, a.count[b]) : a.k) : a.N;
    auto row_ptr = [&](int t, int part) -> const T* {
        return t < a.N ? qkv + ((size_t)b * a.N + t) * 3 * D + part * D + h * DH : nullptr;
    };
    stage_rows<T, DH, LD>(Qs, BQ, [&](int r) { return row_ptr(q0 + r, 0); });
    if (has_bias) {
        const T* gh_ = static_cast<const T*>(a.bias_h) + (((size_t)b * a.H + h) * a.N) * a.gh;
        const T* gw_ = static_cast<const T*>(a.bias_w) + (((size_t)b * a.H + h) * a.N) * a.gw;
        for (int i = threadIdx.x; i < BQ * a.gh; i += kAttnThreads) {
            const int r = i / a.gh, c = i - r * a.gh;
            bh[r * ldh + c] = (q0 + r < a.N) ? gh_[(size_t)(q0 + r) * a.gh + c] : ElemTraits<T>::from_float(0.f);
        }
        for (int i = threadIdx.x; i < BQ * a.gw; i += kAttnThreads) {
            const int r = i / a.gw, c = i - r * a.gw;
            bwS[r * ldw + c] = (q0 + r < a.N) ? gw_[(size_t)(q0 + r) * a.gw + c] : ElemTraits<T>::from_float(0.f);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    uint32_t qf[DH / 16][4];
    load_q_frags<T, DH, LD>(qf, Qs, warp, lane);

    float mrow[2], linv[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int t = min(a.N - 1, q0 + warp * 16 + g + hh * 8);
        const float* st = a.stats + (((size_t)b * a.H + h) * a.N + t) * 2;
        mrow[hh] = st[0];
        linv[hh] = 1.f / st[1];
    }
    float o[DH / 8][4];
#pragma unroll
    for (int i = 0; i < DH / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

    T* a_state = static_cast<T*>(a.a_state);
    const size_t a_head = ((size_t)b * a.H + h) * (size_t)a.N * a.NP;

    for (int key0 = 0; key0 < nkeys; key0 += BKV) {
        __syncthreads();
        if (threadIdx.x < BKV) {
            const int j = key0 + threadIdx.x;
            int tok = -1;
            if (j < nkeys) tok = (MODE == ET_ATTN_DELTA) ? (int)a.idx[(size_t)b * a.k + j] : j;
            s_tok[threadIdx.x] = tok;
        }
        __syncthreads();
        stage_rows<T, DH, LD>(Ks, BKV, [&](int r) { return s_tok[r] >= 0 ? row_ptr(s_tok[r], 1) : nullptr; });
        if (MODE == ET_ATTN_DELTA) {
            const T* dV = static_cast<const T*>(a.dV);
            const T* Vd = static_cast<const T*>(a.Vd);
            stage_rows<T, DH, LD>(V1, BKV, [&](int r) {
                return s_tok[r] >= 0 ? dV + ((size_t)b * a.k + key0 + r) * D + h * DH : nullptr;
            });
            stage_rows<T, DH, LD>(V2, BKV, [&](int r) {
                return s_tok[r] >= 0 ? Vd + ((size_t)b * a.k + key0 + r) * D + h * DH : nullptr;
            });
            // previous attention values of the selected columns: contiguous 64-row segments
            for (int c = threadIdx.x; c < BKV * (BQ / 8); c += kAttnThreads) {
                const int j = c / (BQ / 8), ch = c - j * (BQ / 8);
                T* d = Ps + j * LDP + ch * 8;
                if (s_tok[j] >= 0 && q0 + ch * 8 < a.N)
                    cp_async16(d, a_state + a_head + (size_t)s_tok[j] * a.NP + q0 + ch * 8);
                else
                    *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
            }
        } else {
            stage_rows<T, DH, LD>(V1, BKV, [&](int r) { return s_tok[r] >= 0 ? row_ptr(s_tok[r], 2) : nullptr; });
        }
        cp_async_wait_all();
        __syncthreads();

        float s[8][4];
        qk_tile<T, DH, LD>(s, qf, Ks, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) s[nt][i] *= a.rscale;
        if (has_bias) {
            // bias needs the key's grid position (gathered in DELTA mode)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = warp * 16 + g + (i >> 1) * 8;
                    const int tok = s_tok[nt * 8 + tq * 2 + (i & 1)];
                    if (tok >= 0) {
                        const int ky = tok / a.gw, kx = tok - ky * a.gw;
                        s[nt][i] += ldf(bh + r * ldh + ky) + ldf(bwS + r * ldw + kx);
                    }
                }
        }
        uint32_t pn[4][4], pd[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float av[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool on = s_tok[nt * 8 + tq * 2 + (i & 1)] >= 0;
                // normalised attention value, rounded to dtype exactly as it is stored in the gate state
                av[i] = on ? round_to<T>(exp2f((s[nt][i] - mrow[i >> 1]) * kLog2e) * linv[i >> 1]) : 0.f;
            }
            pn[nt >> 1][(nt & 1) * 2 + 0] = pack2<T>(av[0], av[1]);
            pn[nt >> 1][(nt & 1) * 2 + 1] = pack2<T>(av[2], av[3]);
            if (MODE == ET_ATTN_DELTA) {
                float dv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int col = nt * 8 + tq * 2 + (i & 1);
                    const int r = warp * 16 + g + (i >> 1) * 8;
                    T* cell = Ps + col * LDP + r;
                    dv[i] = av[i] - ldf(cell);           // a_n - p  (modules.py:196)
                    *cell = ElemTraits<T>::from_float(av[i]);  // p[:, idx] = a_n (modules.py:200)
                }
                pd[nt >> 1][(nt & 1) * 2 + 0] = pack2<T>(dv[0], dv[1]);
                pd[nt >> 1][(nt & 1) * 2 + 1] = pack2<T>(dv[2], dv[3]);
            } else if (MODE == ET_ATTN_FIRST) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int col = nt * 8 + tq * 2 + (i & 1);
                    const int r = warp * 16 + g + (i >> 1) * 8;
                    Ps[col * LDP + r] = ElemTraits<T>::from_float(av[i]);
                }
            }
        }
        pv_tile<T, DH, LD>(o, pn, V1, lane);                       // a_n . dV      (or a . v)
        if (MODE == ET_ATTN_DELTA) pv_tile<T, DH, LD>(o, pd, V2, lane);  // dA . (v_n - dV)
        if (MODE != ET_ATTN_DENSE) {
            __syncthreads();
            for (int c = threadIdx.x; c < BKV * (BQ / 8); c += kAttnThreads) {
                const int j = c / (BQ / 8), ch = c - j * (BQ / 8);
                if (s_tok[j] >= 0 && q0 + ch * 8 < a.N)
                    st16(a_state + a_head + (size_t)s_tok[j] * a.NP + q0 + ch * 8,
                         *reinterpret_cast<const uint4*>(Ps + j * LDP + ch * 8));
            }
        }
    }
    // accumulate into the MatmulDeltaAccumulator state and emit the merged-head output
    T* acc = static_cast<T*>(a.acc);
    T* out = static_cast<T*>(a.out);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int t = q0 + warp * 16 + g + hh * 8;
        if (t >= a.N) continue;
        const size_t off = ((size_t)b * a.N + t) * D + h * DH;
#pragma unroll
        for (int nt = 0; nt < DH / 8; ++nt) {
            float v0 = o[nt][hh * 2], v1 = o[nt][hh * 2 + 1];
            const size_t at = off + nt * 8 + tq * 2;
            if (MODE == ET_ATTN_DELTA) {
                v0 += ldf(acc + at);
                v1 += ldf(acc + at + 1);
            }
            const uint32_t packed = pack2<T>(v0, v1);
            if (MODE != ET_ATTN_DENSE) *reinterpret_cast<uint32_t*>(acc + at) = packed;
            *reinterpret_cast<uint32_t*>(out + at) = packed;
        }
    }
}

// ---------------------------------------------------------------- v gate
// DELTA: for each selected row: dV = v - p_v, Vd = v - dV, p_v = v.   FIRST: p_v = v for every token.
template <typename T>
__global__ void __launch_bounds__(256) vgate_kernel(const T* qkv, T* v_state, const long long* idx, const int* count, T* Ksel, T* dV,
                                                    T* Vd, int N, int D, int k, long long total_vec, int emit_vn) {
    et_pdl_prologue();
    const int nch = D / 8;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total_vec;
         gi += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(gi % nch);
        const long long row = gi / nch;
        const long long b = row / k;
        if (count != nullptr && row - b * k >= count[b]) continue;
        const long long tok = idx != nullptr ? idx[row] : row % k;
        const size_t src = ((size_t)b * N + tok) * 3 * D + 2 * D + (size_t)ch * 8;
        const size_t st = ((size_t)b * N + tok) * D + (size_t)ch * 8;
        const uint4 vraw = ld16(qkv + src);
        if (Ksel != nullptr) st16(Ksel + (size_t)row * D + (size_t)ch * 8, ld16(qkv + src - D));  // compact k rows
        if (dV != nullptr) {
            float vn[8], pv[8], d[8], r[8];
            unpack16<T>(vraw, vn);
            unpack16<T>(ld16(v_state + st), pv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                d[i] = round_to<T>(vn[i] - pv[i]);
                r[i] = vn[i] - d[i];
            }
            // tensor-core path consumes v_n (acc += a_n . v_n - p . (v_n - dV)); the mma.sync path consumes dV
            st16(dV + (size_t)row * D + (size_t)ch * 8, emit_vn ? vraw : pack16<T>(d));
            st16(Vd + (size_t)row * D + (size_t)ch * 8, pack16<T>(r));
        }
        st16(v_state + st, vraw);
    }
}

template <typename T, int DH>
constexpr int apply_smem(int gh, int gw) {
    return (BQ * (DH + 8) + 3 * BKV * (DH + 8) + BKV * (BQ + 8) + BQ * (gh + 1 + gw + 1)) * (int)sizeof(T) + 16;
}

// Raises the dynamic-smem limit of a kernel once per (kernel, device, size high-water mark).
template <typename K>
int set_smem(K kernel, int bytes) { return et_raise_smem(kernel, bytes); }

// every sub-buffer of the workspace starts on a 16-byte boundary (element counts rounded up to 8)
inline size_t align8(size_t n) { return (n + 7) / 8 * 8; }

// Bias tables and the v-gate deltas live in a CALLER-PROVIDED workspace (et_attn_workspace_bytes);
// the library never allocates.
template <typename T, int DH>
int run_window(const AttnArgs& a, const void* rel_y, const void* rel_x, void* bias_ws, cudaStream_t s) {
    const int lh = a.windowed ? a.wh : a.gh, lw = a.windowed ? a.ww : a.gw;
    const int nwin = a.windowed ? a.nwx * a.nwy : 1;
    AttnArgs args = a;
    // second-generation tensor-core path (et_attn_win2_tc.cu): half a window per CTA, two CTAs per SM, rel-pos in the CTA
    if (a.windowed && g_attn_tc && g_attn_win_gen == 2 && et_tc_window2_applies(a.wh, a.ww, DH))
        return et_tc_window2_attention(a.qkv, a.pad_token, rel_y, rel_x, a.out, a.B, a.N, a.gh, a.gw, a.wh, a.ww, a.H,
                                       std::is_same_v<T, __nv_bfloat16> ? 1 : 0, s);
    // first-generation tensor-core path: real windows of at most 208 tokens, dh = 64, rel-pos coordinates fitting one 64-column block
    if (a.windowed && DH == 64 && a.Wn <= 208 && (rel_y == nullptr || lh + lw <= 32) && g_attn_tc) {
        if (rel_y != nullptr) {
            T* comb = static_cast<T*>(bias_ws);
            const int smem_rp = ((a.Wn + 1) + (lh * lh + 1) + (lw * lw + 1)) * (DH + 8) * (int)sizeof(T) + 256 * 32 * (int)sizeof(T) + 16;
            int rc = set_smem(relpos_window_kernel<T, DH>, smem_rp);
            if (rc) return rc;
            et_launch(relpos_window_kernel<T, DH>, dim3(dim3(a.B * nwin, a.H)), dim3(kAttnThreads), smem_rp, s, a, static_cast<const T*>(rel_y), static_cast<const T*>(rel_x), comb, 8.f);
            ET_COUNT_LAUNCH(1);
        }
        return et_tc_window_attention(a.qkv, a.pad_token, bias_ws, a.out, a.B, a.N, a.gh, a.gw, a.wh, a.ww, a.H,
                                      rel_y != nullptr ? 1 : 0, std::is_same_v<T, __nv_bfloat16> ? 1 : 0, s);
    }
    if (rel_y != nullptr) {
        T* bh = static_cast<T*>(bias_ws);
        T* bw = bh + align8((size_t)a.B * nwin * a.H * a.Wn * lh);
        et_launch(relpos_bias_kernel<T, DH>, dim3(dim3(lh + lw, a.H, a.B * nwin)), dim3(kAttnThreads), 0, s, a, static_cast<const T*>(rel_y), static_cast<const T*>(rel_x), bh, bw, 0, 1.f, 0, 0);
        ET_COUNT_LAUNCH(1);
        args.bias_h = bh;
        args.bias_w = bw;
    }
    const int smem = (BQ * (DH + 8) + 2 * BKV * (DH + 8) + BQ * (lh + 1 + lw + 1)) * (int)sizeof(T) + 16;
    int rc = set_smem(window_attention_kernel<T, DH>, smem);
    if (rc) return rc;
    et_launch(window_attention_kernel<T, DH>, dim3(dim3((a.Wn + BQ - 1) / BQ, a.H, a.B * nwin)), dim3(kAttnThreads), smem, s, args);
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}

template <typename T, int DH>
int run_global(const AttnArgs& a, const void* rel_y, const void* rel_x, void* v_state, void* ws, cudaStream_t s) {
    AttnArgs args = a;
    const int D = a.H * DH;
    // tensor-core path: dh = 64, at least one full 128-row query block (ragged blocks / tiles are masked), rel-pos bias
    // for token grids up to 64 x 64 (one-hot key coordinates are 64 + 64 columns of the augmented operands)
    const bool use_tc = DH == 64 && a.N >= 128 && (rel_y == nullptr || (a.gw <= 64 && a.gh <= 64)) && g_attn_tc &&
                        a.count == nullptr;  // a device-side key count runs on the mma.sync kernels
    // workspace layout: [bias_h | bias_w | K_sel | dV | Vd | onehot]; the tc path pads bias rows to 64 columns
    const size_t ldh = use_tc ? 64 : a.gh, ldw = use_tc ? 64 : a.gw;
    T* bh = static_cast<T*>(ws);
    T* bw = bh + (rel_y ? align8((size_t)a.B * a.H * a.N * ldh) : 0);
    T* Ksel = bw + (rel_y ? align8((size_t)a.B * a.H * a.N * ldw) : 0);
    T* dV = Ksel + (size_t)a.B * a.k * D;
    T* Vd = dV + (size_t)a.B * a.k * D;
    T* onehot = Vd + (size_t)a.B * a.k * D;
    T* onehot_all = onehot + (size_t)((long long)a.B * a.k > a.N ? (long long)a.B * a.k : a.N) * 128;
    if (rel_y != nullptr) {
        et_launch(relpos_bias_kernel<T, DH>, dim3(dim3(a.gh + a.gw, a.H, a.B)), dim3(kAttnThreads), 0, s, a, static_cast<const T*>(rel_y), static_cast<const T*>(rel_x), bh, bw, use_tc ? 64 : 0, use_tc ? 8.f : 1.f, 0, 0);
        ET_COUNT_LAUNCH(1);
        args.bias_h = bh;
        args.bias_w = bw;
    }
    const T* qkv = static_cast<const T*>(a.qkv);
    if (a.mode == ET_ATTN_DELTA) {
        const long long total = (long long)a.B * a.k * (D / 8);
        if (total > 0)
            et_launch(vgate_kernel<T>, dim3((int)((total + 255) / 256)), dim3(256), 0, s, qkv, static_cast<T*>(v_state), a.idx, a.count,
                                                                      use_tc ? Ksel : nullptr, dV, Vd, a.N, D, a.k, total,
                                                                      use_tc ? 1 : 0);
        ET_COUNT_LAUNCH(1);
        args.dV = dV;
        args.Vd = Vd;
    } else if (a.mode == ET_ATTN_FIRST) {
        const long long total = (long long)a.B * a.N * (D / 8);
        et_launch(vgate_kernel<T>, dim3((int)((total + 255) / 256)), dim3(256), 0, s, qkv, static_cast<T*>(v_state), nullptr, nullptr, nullptr,
                                                                  nullptr, nullptr, a.N, D, a.N, total, 0);
        ET_COUNT_LAUNCH(1);
    }
    if (use_tc) {
        if (a.mode == ET_ATTN_DELTA && a.k == 0) return ET_OK;
        return et_tc_global_attention(a.qkv, Ksel, onehot, onehot_all, args.bias_h, args.bias_w, a.mode, a.idx, a.k, a.a_state, a.acc, a.out,
                                      a.stats, a.B, a.N, a.NP, a.H, a.gh, a.gw, std::is_same_v<T, __nv_bfloat16> ? 1 : 0, s);
    }
    const dim3 grid((a.N + BQ - 1) / BQ, a.H, a.B);
    const int smem_a = (BQ * (DH + 8) + BKV * (DH + 8) + BQ * (a.gh + 1 + a.gw + 1)) * (int)sizeof(T) + 16;
    int rc = set_smem(attn_stats_kernel<T, DH>, smem_a);
    if (rc) return rc;
    et_launch(attn_stats_kernel<T, DH>, dim3(grid), dim3(kAttnThreads), smem_a, s, args);
    ET_COUNT_LAUNCH(1);
    const int smem_b = apply_smem<T, DH>(a.gh, a.gw);
    if (a.mode == ET_ATTN_DELTA) {
        if ((rc = set_smem(attn_apply_kernel<T, DH, ET_ATTN_DELTA>, smem_b))) return rc;
        if (a.k > 0) et_launch(attn_apply_kernel<T, DH, ET_ATTN_DELTA>, dim3(grid), dim3(kAttnThreads), smem_b, s, args);
        ET_COUNT_LAUNCH(1);
    } else if (a.mode == ET_ATTN_FIRST) {
        if ((rc = set_smem(attn_apply_kernel<T, DH, ET_ATTN_FIRST>, smem_b))) return rc;
        et_launch(attn_apply_kernel<T, DH, ET_ATTN_FIRST>, dim3(grid), dim3(kAttnThreads), smem_b, s, args);
        ET_COUNT_LAUNCH(1);
    } else {
        if ((rc = set_smem(attn_apply_kernel<T, DH, ET_ATTN_DENSE>, smem_b))) return rc;
        et_launch(attn_apply_kernel<T, DH, ET_ATTN_DENSE>, dim3(grid), dim3(kAttnThreads), smem_b, s, args);
        ET_COUNT_LAUNCH(1);
    }
    return ET_OK;
}

#define ET_DISPATCH_DH(dh, DH, ...)                                                              \
    switch (dh) {                                                                                \
        case 16: { constexpr int DH = 16; __VA_ARGS__; break; }                                  \
        case 32: { constexpr int DH = 32; __VA_ARGS__; break; }                                  \
        case 64: { constexpr int DH = 64; __VA_ARGS__; break; }                                  \
        default: return et_fail(ET_ERR_UNSUPPORTED, "head dim %d not supported (16, 32, 64)", (int)(dh)); \
    }

}  // namespace

extern "C" {

// Bytes of caller-provided scratch needed by et_window_attention / et_global_attention (an upper bound over every
// code path: tensor-core layouts, the mma.sync tables, and the general-precision path's fp32 statistics / v-gate deltas).
int64_t et_attn_workspace_bytes(int64_t B, int64_t N, int64_t gh, int64_t gw, int64_t wh, int64_t ww, int64_t heads,
                                int64_t dh, int64_t k, int has_relpos) {
    int64_t elems = 0, extra = 0;
    if (wh > 0) {
        const int64_t nw = ((gh + wh - 1) / wh) * ((gw + ww - 1) / ww);
        if (has_relpos) elems += align8(B * nw * heads * wh * ww * wh) + align8(B * nw * heads * wh * ww * ww);
        if (has_relpos) elems += B * nw * heads * 256 * 64 + 256 * 64;  // tensor-core layout: combined bias rows + one-hot block
        extra = B * nw * heads * wh * ww * 2 * 4;                       // general path: (max, sum) per query row, fp32
    } else {
        // upper bound over both layouts (the tensor-core path pads bias rows to 64 columns and adds a one-hot scratch)
        if (has_relpos) elems += align8(B * heads * N * (gh > 64 ? gh : 64)) + align8(B * heads * N * (gw > 64 ? gw : 64));
        elems += 3 * align8(B * k * heads * dh);
        elems += ((B * k > N) ? B * k : N) * 128 + N * 128;  // one-hot key coordinates: selected keys, and all keys
        extra = B * heads * N * 2 * 4;  // statistics of the small single-window case
    }
    return elems * 2 + extra + 512;
}

int et_window_attention(const void* qkv, const void* pad_token, const void* rel_y, const void* rel_x, void* out,
                           void* workspace, int64_t B, int64_t N, int64_t gh, int64_t gw, int64_t wh, int64_t ww,
                           int64_t heads, int64_t dh, int dtype, void* stream) {
    ET_CHECK_ARG(qkv && out, "et_window_attention: null pointer");
    ET_CHECK_ARG(dtype == ET_BF16 || dtype == ET_F16 || dtype == ET_F32, "et_window_attention: bad dtype %d", dtype);
    ET_CHECK_ARG((rel_y == nullptr) == (rel_x == nullptr), "et_window_attention: rel_y / rel_x both or neither");
    ET_CHECK_ARG((rel_y == nullptr && dtype != ET_F32) || workspace != nullptr, "et_window_attention: workspace required");
    if (dtype == ET_F32) {  // general-precision path (et_generic.cu): fp32 arithmetic on the CUDA cores
        ET_CHECK_ARG(dh <= 64, "et_window_attention (fp32): head dim %lld > 64", (long long)dh);
        ET_CHECK_ARG(wh == 0 || (gh * gw == N && ww > 0), "et_window_attention: windowed attention needs N == gh * gw");
        ET_CHECK_ARG(wh == 0 || pad_token != nullptr || (gh % wh == 0 && gw % ww == 0), "et_window_attention: padding needs pad_token");
        ET_CHECK_ARG(wh > 0 || rel_y == nullptr || gh * gw == N, "et_window_attention: rel-pos needs N == gh * gw");
        // the statistics live at the END of the workspace (the front is the 16-bit paths' bias scratch)
        const int64_t total = et_attn_workspace_bytes(B, N, gh, gw, wh, ww, heads, dh, 0, rel_y != nullptr);
        const int64_t nw = wh > 0 ? ((gh + wh - 1) / wh) * ((gw + ww - 1) / ww) : 1;
        const int64_t need = B * nw * heads * (wh > 0 ? wh * ww : N) * 2 * 4;
        float* stats = reinterpret_cast<float*>(static_cast<char*>(workspace) + (total - need - 256) / 16 * 16);
        int rc = et_generic_window_attention(qkv, pad_token, rel_y, rel_x, out, stats, (int)B, (int)N, (int)gh, (int)gw, (int)wh,
                                             (int)ww, (int)heads, (int)dh, dtype, et_stream(stream));
        if (rc) return rc;
        ET_CHECK_LAUNCH("et_window_attention");
        return ET_OK;
    }
    ET_CHECK_ARG(et_aligned16(qkv) && et_aligned16(out) && et_aligned16(pad_token) && et_aligned16(workspace),
                 "et_window_attention: pointers must be 16-byte aligned");
    AttnArgs a = {};
    a.qkv = qkv; a.pad_token = pad_token; a.out = out;
    a.B = (int)B; a.N = (int)N; a.gh = (int)gh; a.gw = (int)gw; a.wh = (int)wh; a.ww = (int)ww; a.H = (int)heads;
    a.windowed = wh > 0;
    if (a.windowed) {
        ET_CHECK_ARG(gh * gw == N && ww > 0, "et_window_attention: windowed attention needs N == gh * gw");
        a.nwy = (int)((gh + wh - 1) / wh); a.nwx = (int)((gw + ww - 1) / ww); a.Wn = (int)(wh * ww);
        ET_CHECK_ARG(pad_token != nullptr || (gh % wh == 0 && gw % ww == 0), "et_window_attention: padding needs pad_token");
    } else {
        a.nwx = a.nwy = 1; a.Wn = (int)N;
        ET_CHECK_ARG(rel_y == nullptr || gh * gw == N, "et_window_attention: rel-pos needs N == gh * gw");
    }
    a.rscale = 1.0f / sqrtf((float)dh);
    int rc = ET_OK;
    ET_DISPATCH_DTYPE(dtype, T, {
        if constexpr (sizeof(T) == 2) {
            ET_DISPATCH_DH(dh, DH, rc = run_window<T, DH>(a, rel_y, rel_x, workspace, et_stream(stream)));
        }
    });
    if (rc) return rc;
    ET_CHECK_LAUNCH("et_window_attention");
    return ET_OK;
}

int et_global_attention(const void* qkv, const void* kv_pooled, int64_t pool_h, int64_t pool_w, const void* rel_y,
                        const void* rel_x, int mode, const int64_t* idx, const int32_t* count, int64_t k, void* a_state,
                        void* v_state, void* acc, void* out, float* row_stats, void* workspace, int64_t B, int64_t N,
                        int64_t gh, int64_t gw, int64_t heads, int64_t dh, int dtype, int state_dtype, void* stream) {
    ET_CHECK_ARG(qkv && out && row_stats, "et_global_attention: null pointer");
    ET_CHECK_ARG(dtype == ET_BF16 || dtype == ET_F16 || dtype == ET_F32, "et_global_attention: bad dtype %d", dtype);
    ET_CHECK_ARG(state_dtype == ET_BF16 || state_dtype == ET_F16 || state_dtype == ET_F32, "et_global_attention: bad state dtype %d", state_dtype);
    ET_CHECK_ARG(mode == ET_ATTN_DENSE || mode == ET_ATTN_FIRST || mode == ET_ATTN_DELTA, "et_global_attention: bad mode");
    ET_CHECK_ARG((rel_y == nullptr) == (rel_x == nullptr), "et_global_attention: rel_y / rel_x both or neither");
    ET_CHECK_ARG(rel_y == nullptr || gh * gw == N, "et_global_attention: rel-pos needs N == gh * gw");
    ET_CHECK_ARG(mode == ET_ATTN_DENSE || (a_state && v_state && acc), "et_global_attention: state pointers required");
    ET_CHECK_ARG(mode != ET_ATTN_DELTA || (idx != nullptr && k >= 0 && k <= N), "et_global_attention: DELTA needs idx, k <= N");
    ET_CHECK_ARG((mode != ET_ATTN_DELTA && rel_y == nullptr) || workspace != nullptr, "et_global_attention: workspace required");
    ET_CHECK_ARG(count == nullptr || mode == ET_ATTN_DELTA, "et_global_attention: count applies to DELTA mode");
    if (kv_pooled != nullptr)
        ET_CHECK_ARG(pool_h > 0 && pool_w > 0 && gh * gw == N && gh % pool_h == 0 && gw % pool_w == 0,
                     "et_global_attention: pooled keys need a %lld x %lld grid divisible by the pool size", (long long)gh, (long long)gw);
    if (dtype == ET_F32 || state_dtype != dtype || kv_pooled != nullptr) {
        // general-precision path (et_generic.cu): fp32 models, matmul_2_cast != model dtype, pooled keys / values
        ET_CHECK_ARG(dh <= 64, "et_global_attention (general path): head dim %lld > 64", (long long)dh);
        int rc = et_generic_global_attention(qkv, kv_pooled, (int)pool_h, (int)pool_w, rel_y, rel_x, mode, idx, count, (int)k, a_state,
                                             v_state, acc, out, row_stats, workspace, (int)B, (int)N, (int)gh, (int)gw, (int)heads,
                                             (int)dh, dtype, state_dtype, et_stream(stream));
        if (rc) return rc;
        ET_CHECK_LAUNCH("et_global_attention");
        return ET_OK;
    }
    AttnArgs a = {};
    a.qkv = qkv; a.out = out; a.B = (int)B; a.N = (int)N; a.NP = (int)((N + 7) / 8 * 8); a.gh = (int)gh; a.gw = (int)gw;
    a.H = (int)heads;
    a.windowed = 0; a.nwx = a.nwy = 1; a.Wn = (int)N; a.rscale = 1.0f / sqrtf((float)dh);
    a.idx = reinterpret_cast<const long long*>(idx); a.count = count; a.k = (int)(mode == ET_ATTN_DELTA ? k : 0); a.mode = mode;
    a.a_state = a_state; a.acc = acc; a.stats = row_stats;
    if (rel_y == nullptr) { a.gh = 0; a.gw = 0; }
    int rc = ET_OK;
    ET_DISPATCH_DTYPE(dtype, T, {
        if constexpr (sizeof(T) == 2) {
            ET_DISPATCH_DH(dh, DH, rc = run_global<T, DH>(a, rel_y, rel_x, v_state, workspace, et_stream(stream)));
        }
    });
    if (rc) return rc;
    ET_CHECK_LAUNCH("et_global_attention");
    return ET_OK;
}

}  // extern "C"
