// Global attention of the Eventful blocks on 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
// Fast path for dh = 64, N % 128 == 0 and (no rel-pos | 64-wide token grid); everything else takes the
// mma.sync kernels in et_attn.cu.  Statistics are kept in the log2 domain: m2 = max(x) * log2(e).
//
//   tc_stats_kernel  (phase A)  S = Q K^T per 128 x 128 tile in TMEM (double buffered), softmax warps read
//                               their own row with tcgen05.ld, add the rel-pos bias held in registers and keep
//                               the running row max / row sum.  Output: (m2, l) per row.
//   tc_apply_kernel  (phase B)  per 128-row query block and 64-key tile of the SELECTED keys:
//                               S = Q K_sel^T (TMEM)  ->  a_n = exp2(S - m2) / l (bf16, exactly as stored)
//                               A-gate: dA = a_n - a_state[:, idx]; a_state[:, idx] = a_n
//                                       (column-major state: one 256-byte cp.async.bulk per column each way)
//                               accumulate: O += a_n . dV + dA . (v_n - dV)  (two tcgen05 MMAs, P from smem,
//                                       V tiles as MN-major B operands straight from TMA)
//                               epilogue: acc += O; out = acc.
//                    FIRST / DENSE modes run the same pipeline over all keys with V from the QKV buffer.
// Warp roles (192 threads): warp 0 = TMA / bulk-copy producer + state write-back, warp 1 = TMEM allocator and
// single-thread MMA issuer, warps 2-5 = softmax / gate / epilogue (one query row per thread).
#include "et_tcgen05.cuh"

using namespace et_tc;

namespace {

constexpr int kThreads = 192;
constexpr int QROWS = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct TcArgs {
    const void* bias_h;
    const void* bias_w;
    float* stats;
    const long long* idx;
    void* a_state;
    void* acc;
    void* out;
    int B, N, NP, H, D, gh, gw, k, is_bf16, sel_rows;
    float c1;  // (1 / sqrt(dh)) * log2(e)
};

template <bool BF16>
__device__ __forceinline__ float elem_to_float(uint16_t raw) {
    if constexpr (BF16) return __uint_as_float((uint32_t)raw << 16);
    else return __half2float(__ushort_as_half(raw));
}
template <bool BF16>
__device__ __forceinline__ uint16_t float_to_elem(float v) {
    if constexpr (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
    else return __half_as_ushort(__float2half_rn(v));
}

__device__ __forceinline__ void named_sync_softmax() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ============================================================================================= phase A
constexpr int ST_KEYS = 128;
constexpr int ST_STAGES = 4;
constexpr int ST_TILE = ST_KEYS * 64 * 2;  // 16 KB
constexpr int ST_SMEM = QROWS * 64 * 2 + ST_STAGES * ST_TILE + 256 + 1024;

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1) tc_stats_kernel(const __grid_constant__ CUtensorMap tm_qkv, const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Qs = smem;
    uint8_t* Ks = smem + QROWS * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Ks + ST_STAGES * ST_TILE);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + ST_STAGES;
    uint64_t* s_full = k_empty + ST_STAGES;
    uint64_t* s_empty = s_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * QROWS, h = blockIdx.y, b = blockIdx.z;
    const int T = a.N / ST_KEYS;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_qkv) : "memory");
        mbar_init(smem_u32(q_full), 1);
        for (int s = 0; s < ST_STAGES; ++s) {
            mbar_init(smem_u32(&k_full[s]), 1);
            mbar_init(smem_u32(&k_empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(smem_u32(q_full), QROWS * 128);
            tma_load_2d(smem_u32(Qs), &tm_qkv, smem_u32(q_full), h * 64, b * a.N + q0);
            for (int t = 0; t < T; ++t) {
                const int s = t % ST_STAGES;
                mbar_wait(smem_u32(&k_empty[s]), ((t / ST_STAGES) & 1) ^ 1);
                mbar_expect_tx(smem_u32(&k_full[s]), ST_TILE);
                tma_load_2d(smem_u32(Ks + s * ST_TILE), &tm_qkv, smem_u32(&k_full[s]), a.D + h * 64, b * a.N + t * ST_KEYS);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_ex(128, ST_KEYS, a.is_bf16, 0);
            mbar_wait(smem_u32(q_full), 0);
            const uint64_t dq = umma_smem_desc(smem_u32(Qs));
            for (int t = 0; t < T; ++t) {
                const int s = t % ST_STAGES, u = t & 1;
                mbar_wait(smem_u32(&k_full[s]), (t / ST_STAGES) & 1);
                mbar_wait(smem_u32(&s_empty[u]), ((t >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint64_t dk = umma_smem_desc(smem_u32(Ks + s * ST_TILE));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tcgen05_mma_f16(tmem_base + u * ST_KEYS, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc, kk > 0);
                tcgen05_commit(smem_u32(&k_empty[s]));
                tcgen05_commit(smem_u32(&s_full[u]));
            }
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const size_t grow = ((size_t)b * a.H + h) * a.N + q0 + row;
        const bool has_bias = a.bias_h != nullptr;
        // rel-pos bias of this query row, pre-scaled by log2(e): bw for the 64 key columns, bh per key image row
        float bwl[64];
        const uint16_t* bh_row = nullptr;
        if (has_bias) {
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(a.bias_w) + grow * 64);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 u4 = src[c];
                const uint16_t* e = reinterpret_cast<const uint16_t*>(&u4);
#pragma unroll
                for (int i = 0; i < 8; ++i) bwl[c * 8 + i] = elem_to_float<BF16>(e[i]) * kLog2e;
            }
            bh_row = static_cast<const uint16_t*>(a.bias_h) + grow * a.gh;
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) bwl[i] = 0.f;
        }
        float m2 = -INFINITY, l = 0.f;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (int t = 0; t < T; ++t) {
            const int u = t & 1;
            float bh0 = 0.f, bh1 = 0.f;
            if (has_bias) {  // a 128-key tile spans two image rows of the 64-wide grid
                const uint32_t pair = *reinterpret_cast<const uint32_t*>(bh_row + 2 * t);
                bh0 = elem_to_float<BF16>((uint16_t)(pair & 0xffffu)) * kLog2e;
                bh1 = elem_to_float<BF16>((uint16_t)(pair >> 16)) * kLog2e;
            }
            mbar_wait(smem_u32(&s_full[u]), (t >> 1) & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_load_32x32(taddr + (uint32_t)(u * ST_KEYS + c * 32), v);
                if (c == 3) {  // whole tile read: hand the TMEM buffer back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&s_empty[u]));
                }
                const float bh = (c < 2) ? bh0 : bh1;
                float x[32];
                float cmax = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    x[i] = fmaf(__uint_as_float(v[i]), a.c1, bwl[(c & 1) * 32 + i]) + bh;
                    cmax = fmaxf(cmax, x[i]);
                }
                const float mn = fmaxf(m2, cmax);
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) sum += exp2f(x[i] - mn);
                l = l * exp2f(m2 - mn) + sum;
                m2 = mn;
            }
        }
        a.stats[grow * 2] = m2;
        a.stats[grow * 2 + 1] = l;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ============================================================================================= phase B
constexpr int AP_KEYS = 64;
constexpr int AP_KV = AP_KEYS * 64 * 2;           // 8 KB: one K / V tile
constexpr int AP_STAGE = 3 * AP_KV;               // K, V1, V2
constexpr int AP_PT = AP_KEYS * QROWS * 2;        // 16 KB: a_state tile [key][row]
constexpr int AP_P = QROWS * AP_KEYS * 2;         // 16 KB: P tile (A operand)
constexpr int AP_BIAS_LD = 130;                   // halfwords per bias-table row (odd word count: no bank conflicts)
constexpr int AP_OFF_ST = QROWS * 128;            // 16 KB Q
constexpr int AP_OFF_PT = AP_OFF_ST + 2 * AP_STAGE;
constexpr int AP_OFF_P = AP_OFF_PT + 2 * AP_PT;
constexpr int AP_OFF_BIAS = AP_OFF_P + 4 * AP_P;
constexpr int AP_OFF_MISC = AP_OFF_BIAS + ((QROWS * AP_BIAS_LD * 2 + 1023) / 1024) * 1024;
constexpr int AP_SMEM = AP_OFF_MISC + 2048 + 1024;

template <bool BF16, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
tc_apply_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Qs = smem;
    auto Kt = [&](int u) { return smem + AP_OFF_ST + u * AP_STAGE; };
    auto V1 = [&](int u) { return smem + AP_OFF_ST + u * AP_STAGE + AP_KV; };
    auto V2 = [&](int u) { return smem + AP_OFF_ST + u * AP_STAGE + 2 * AP_KV; };
    auto Pt = [&](int u) { return reinterpret_cast<uint16_t*>(smem + AP_OFF_PT + u * AP_PT); };
    auto Pn = [&](int u) { return smem + AP_OFF_P + u * 2 * AP_P; };
    auto Pd = [&](int u) { return smem + AP_OFF_P + u * 2 * AP_P + AP_P; };
    uint16_t* bias_tab = reinterpret_cast<uint16_t*>(smem + AP_OFF_BIAS);
    int* s_tok = reinterpret_cast<int*>(smem + AP_OFF_MISC);           // [2][64]
    uint8_t* s_ky = reinterpret_cast<uint8_t*>(s_tok + 2 * AP_KEYS);    // [2][64]
    uint8_t* s_kx = s_ky + 2 * AP_KEYS;                                 // [2][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AP_OFF_MISC + 1024);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* ps_full = bars + 3;
    uint64_t* s_full = bars + 5;
    uint64_t* s_empty = bars + 7;
    uint64_t* p_ready = bars + 9;
    uint64_t* pv_done = bars + 11;
    uint64_t* ps_done = bars + 13;
    uint64_t* o_full = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * QROWS, h = blockIdx.y, b = blockIdx.z;
    const int nkeys = (MODE == ET_ATTN_DELTA) ? a.k : a.N;
    const int T = (nkeys + AP_KEYS - 1) / AP_KEYS;
    const bool has_bias = a.bias_h != nullptr;
    uint16_t* a_state = static_cast<uint16_t*>(a.a_state);
    const size_t a_head = ((size_t)b * a.H + h) * (size_t)a.N * a.NP;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_kv) : "memory");
        mbar_init(smem_u32(q_full), 1);
        mbar_init(smem_u32(o_full), 1);
        for (int u = 0; u < 2; ++u) {
            mbar_init(smem_u32(&kv_full[u]), 1);
            mbar_init(smem_u32(&ps_full[u]), 1);
            mbar_init(smem_u32(&s_full[u]), 1);
            mbar_init(smem_u32(&s_empty[u]), 4);
            mbar_init(smem_u32(&p_ready[u]), 4);
            mbar_init(smem_u32(&pv_done[u]), 1);
            mbar_init(smem_u32(&ps_done[u]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 2 * AP_KEYS;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer + state write-back
        if (lane == 0) {
            mbar_expect_tx(smem_u32(q_full), QROWS * 128);
            tma_load_2d(smem_u32(Qs), &tm_q, smem_u32(q_full), h * 64, b * a.N + q0);
        }
        auto write_back = [&](int tt) {  // a_state[:, idx of tile tt] <- a_n (softmax warps left it in Pt)
            const int u = tt & 1;
            mbar_wait(smem_u32(&ps_done[u]), (tt >> 1) & 1);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int j = lane + half * 32;
                const int tok = s_tok[u * AP_KEYS + j];
                if (tok >= 0) bulk_store(a_state + a_head + (size_t)tok * a.NP + q0, smem_u32(Pt(u) + j * QROWS), QROWS * 2);
            }
            bulk_commit();
        };
        for (int t = 0; t < T; ++t) {
            const int u = t & 1, key0 = t * AP_KEYS;
            if (t >= 2) mbar_wait(smem_u32(&pv_done[u]), ((t >> 1) & 1) ^ 1);  // tile t-2 fully consumed
            if (MODE != ET_ATTN_DENSE) bulk_wait_read_all();                    // its write-back has left smem
            int tok[2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int j = key0 + lane + half * 32;
                tok[half] = -1;
                if (j < nkeys) tok[half] = (MODE == ET_ATTN_DELTA) ? (int)a.idx[(size_t)b * a.k + j] : j;
                s_tok[u * AP_KEYS + lane + half * 32] = tok[half];
                if (has_bias && tok[half] >= 0) {
                    s_ky[u * AP_KEYS + lane + half * 32] = (uint8_t)(tok[half] / a.gw);
                    s_kx[u * AP_KEYS + lane + half * 32] = (uint8_t)(tok[half] % a.gw);
                }
            }
            __syncwarp();
            if (lane == 0) {
                const uint32_t fb = smem_u32(&kv_full[u]);
                if (MODE == ET_ATTN_DELTA) {
                    mbar_expect_tx(fb, 3 * AP_KV);
                    const int r0 = b * a.k + key0;
                    tma_load_2d(smem_u32(Kt(u)), &tm_kv, fb, h * 64, r0);
                    tma_load_2d(smem_u32(V1(u)), &tm_kv, fb, h * 64, a.sel_rows + r0);
                    tma_load_2d(smem_u32(V2(u)), &tm_kv, fb, h * 64, 2 * a.sel_rows + r0);
                    const int nvalid = min(AP_KEYS, nkeys - key0);
                    mbar_expect_tx(smem_u32(&ps_full[u]), (uint32_t)nvalid * QROWS * 2);
                } else {
                    mbar_expect_tx(fb, 2 * AP_KV);
                    tma_load_2d(smem_u32(Kt(u)), &tm_kv, fb, a.D + h * 64, b * a.N + key0);
                    tma_load_2d(smem_u32(V1(u)), &tm_kv, fb, 2 * a.D + h * 64, b * a.N + key0);
                }
            }
            __syncwarp();
            if (MODE == ET_ATTN_DELTA) {
#pragma unroll
                for (int half = 0; half < 2; ++half)
                    if (tok[half] >= 0)
                        bulk_load(smem_u32(Pt(u) + (lane + half * 32) * QROWS),
                                  a_state + a_head + (size_t)tok[half] * a.NP + q0, QROWS * 2, smem_u32(&ps_full[u]));
            }
            if (MODE != ET_ATTN_DENSE && t >= 1) write_back(t - 1);
        }
        if (MODE != ET_ATTN_DENSE && T >= 1) write_back(T - 1);
        bulk_wait_all();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_ex(128, AP_KEYS, a.is_bf16, 0);
            const uint32_t idesc_o = umma_idesc_ex(128, 64, a.is_bf16, 1);  // B = V tile, MN-major
            mbar_wait(smem_u32(q_full), 0);
            const uint64_t dq = umma_smem_desc(smem_u32(Qs));
            auto issue_s = [&](int t) {
                const int u = t & 1;
                mbar_wait(smem_u32(&kv_full[u]), (t >> 1) & 1);
                mbar_wait(smem_u32(&s_empty[u]), ((t >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint64_t dk = umma_smem_desc(smem_u32(Kt(u)));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tcgen05_mma_f16(tmem_base + u * AP_KEYS, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc_s, kk > 0);
                tcgen05_commit(smem_u32(&s_full[u]));
            };
            auto issue_pv = [&](int t) {
                const int u = t & 1;
                mbar_wait(smem_u32(&p_ready[u]), (t >> 1) & 1);
                tcgen05_fence_after();
                const uint64_t dpn = umma_smem_desc(smem_u32(Pn(u)));
                const uint64_t dv1 = umma_smem_desc_mn(smem_u32(V1(u)));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)  // 16 keys per step: +32 B in P rows, +16 rows (2048 B) in the V tile
                    tcgen05_mma_f16(tmem_o, dpn + (uint64_t)(2 * kk), dv1 + (uint64_t)(128 * kk), idesc_o, (t > 0 || kk > 0));
                if (MODE == ET_ATTN_DELTA) {
                    const uint64_t dpd = umma_smem_desc(smem_u32(Pd(u)));
                    const uint64_t dv2 = umma_smem_desc_mn(smem_u32(V2(u)));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tcgen05_mma_f16(tmem_o, dpd + (uint64_t)(2 * kk), dv2 + (uint64_t)(128 * kk), idesc_o, 1u);
                }
                tcgen05_commit(smem_u32(&pv_done[u]));
                if (t == T - 1) tcgen05_commit(smem_u32(o_full));
            };
            if (T > 0) issue_s(0);
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) issue_s(t + 1);
                issue_pv(t);
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax / gate / epilogue
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const size_t grow = ((size_t)b * a.H + h) * a.N + q0 + row;
        const int wtab = a.gh + a.gw;
        if (has_bias) {  // this CTA's 128 bias rows -> padded smem table [row][bh | bw]
            const uint16_t* gbh = static_cast<const uint16_t*>(a.bias_h) + (grow - row) * a.gh;
            const uint16_t* gbw = static_cast<const uint16_t*>(a.bias_w) + (grow - row) * a.gw;
            const int tid = threadIdx.x - 64;
            for (int i = tid; i < QROWS * a.gh; i += 128) bias_tab[(i / a.gh) * AP_BIAS_LD + i % a.gh] = gbh[i];
            for (int i = tid; i < QROWS * a.gw; i += 128) bias_tab[(i / a.gw) * AP_BIAS_LD + a.gh + i % a.gw] = gbw[i];
            named_sync_softmax();
        }
        (void)wtab;
        const float m2 = a.stats[grow * 2];
        const float linv = 1.f / a.stats[grow * 2 + 1];
        const uint16_t* brow = bias_tab + row * AP_BIAS_LD;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (int t = 0; t < T; ++t) {
            const int u = t & 1;
            const uint32_t ph = (t >> 1) & 1;
            mbar_wait(smem_u32(&s_full[u]), ph);
            tcgen05_fence_after();
            uint32_t v[64];
            {
                uint32_t lo[32], hi[32];
                tmem_load_32x32(taddr + (uint32_t)(u * AP_KEYS), lo);
                tmem_load_32x32(taddr + (uint32_t)(u * AP_KEYS + 32), hi);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    v[i] = lo[i];
                    v[32 + i] = hi[i];
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_empty[u]));
            const int* tk = s_tok + u * AP_KEYS;
            // normalised attention values of the selected columns, rounded to dtype exactly as stored in the state
            uint32_t an[32], ad[32];  // bf16/fp16 pairs: element j in half (j & 1) of word j >> 1
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                float bias = 0.f;
                if (has_bias)
                    bias = (elem_to_float<BF16>(brow[s_ky[u * AP_KEYS + j]]) +
                            elem_to_float<BF16>(brow[a.gh + s_kx[u * AP_KEYS + j]])) * kLog2e;
                const float p = exp2f(fmaf(__uint_as_float(v[j]), a.c1, bias) - m2) * linv;
                const uint32_t e = tk[j] >= 0 ? (uint32_t)float_to_elem<BF16>(p) : 0u;
                if (j & 1) an[j >> 1] |= e << 16;
                else an[j >> 1] = e;
            }
            if (MODE == ET_ATTN_DELTA) {
                mbar_wait(smem_u32(&ps_full[u]), ph);
                uint16_t* pt = Pt(u) + row;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const uint16_t cur = (uint16_t)((an[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
                    const float prev = tk[j] >= 0 ? elem_to_float<BF16>(pt[j * QROWS]) : 0.f;
                    const uint32_t d = float_to_elem<BF16>(elem_to_float<BF16>(cur) - prev);  // dA = a_n - p (modules.py:196)
                    if (j & 1) ad[j >> 1] |= d << 16;
                    else ad[j >> 1] = d;
                    pt[j * QROWS] = cur;                                                       // p[:, idx] = a_n (modules.py:200)
                }
            } else if (MODE == ET_ATTN_FIRST) {
                uint16_t* pt = Pt(u) + row;
#pragma unroll
                for (int j = 0; j < 64; ++j) pt[j * QROWS] = (uint16_t)((an[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
            }
            if (t >= 2) mbar_wait(smem_u32(&pv_done[u]), ph ^ 1);  // the P tiles of tile t-2 have been consumed
            // P tiles as K-major, 128-byte-swizzled A operands: row r, 16-byte chunk c -> c ^ (r % 8)
            {
                uint8_t* pn_row = Pn(u) + (row >> 3) * 1024 + (row & 7) * 128;
                uint8_t* pd_row = Pd(u) + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int sw = (c ^ (row & 7)) * 16;
                    *reinterpret_cast<uint4*>(pn_row + sw) = make_uint4(an[c * 4], an[c * 4 + 1], an[c * 4 + 2], an[c * 4 + 3]);
                    if (MODE == ET_ATTN_DELTA)
                        *reinterpret_cast<uint4*>(pd_row + sw) = make_uint4(ad[c * 4], ad[c * 4 + 1], ad[c * 4 + 2], ad[c * 4 + 3]);
                }
            }
            fence_proxy_async();  // generic-proxy smem writes -> visible to tcgen05.mma and the bulk stores
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&p_ready[u]));
                if (MODE != ET_ATTN_DENSE) mbar_arrive(smem_u32(&ps_done[u]));
            }
        }
        // ---- epilogue: acc += O, out = acc
        uint16_t* acc = static_cast<uint16_t*>(a.acc);
        uint16_t* out = static_cast<uint16_t*>(a.out);
        const size_t off = ((size_t)b * a.N + q0 + row) * a.D + h * 64;
        float o[64];
        if (T > 0) {
            mbar_wait(smem_u32(o_full), 0);
            tcgen05_fence_after();
            uint32_t lo[32], hi[32];
            tmem_load_32x32(taddr + (uint32_t)(2 * AP_KEYS), lo);
            tmem_load_32x32(taddr + (uint32_t)(2 * AP_KEYS + 32), hi);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                o[i] = __uint_as_float(lo[i]);
                o[32 + i] = __uint_as_float(hi[i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
            uint4 prev = make_uint4(0, 0, 0, 0);
            if (MODE == ET_ATTN_DELTA) prev = *reinterpret_cast<const uint4*>(acc + off + c * 8);
            const uint32_t pw[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float lo = o[c * 8 + 2 * i], hi = o[c * 8 + 2 * i + 1];
                if (MODE == ET_ATTN_DELTA) {
                    lo += elem_to_float<BF16>((uint16_t)(pw[i] & 0xffffu));
                    hi += elem_to_float<BF16>((uint16_t)(pw[i] >> 16));
                }
                w[i] = (uint32_t)float_to_elem<BF16>(lo) | ((uint32_t)float_to_elem<BF16>(hi) << 16);
            }
            const uint4 pk = make_uint4(w[0], w[1], w[2], w[3]);
            if (MODE != ET_ATTN_DENSE) *reinterpret_cast<uint4*>(acc + off + c * 8) = pk;
            *reinterpret_cast<uint4*>(out + off + c * 8) = pk;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

template <typename K>
int raise_smem(K kernel, int bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return et_fail(ET_ERR_CUDA, "cudaFuncSetAttribute(%d bytes): %s", bytes, cudaGetErrorString(e));
    return ET_OK;
}

template <bool BF16>
int launch_tc(const void* qkv, const void* sel, const TcArgs& a, int mode, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        int rc;
        if ((rc = raise_smem(tc_stats_kernel<BF16>, ST_SMEM))) return rc;
        if ((rc = raise_smem(tc_apply_kernel<BF16, ET_ATTN_DENSE>, AP_SMEM))) return rc;
        if ((rc = raise_smem(tc_apply_kernel<BF16, ET_ATTN_FIRST>, AP_SMEM))) return rc;
        if ((rc = raise_smem(tc_apply_kernel<BF16, ET_ATTN_DELTA>, AP_SMEM))) return rc;
        configured = true;
    }
    CUtensorMap tm128, tm64, tmsel;
    int rc;
    if ((rc = make_tmap_2d(&tm128, qkv, (long long)a.B * a.N, 3LL * a.D, 128, a.is_bf16))) return rc;
    if ((rc = make_tmap_2d(&tm64, qkv, (long long)a.B * a.N, 3LL * a.D, 64, a.is_bf16))) return rc;
    const dim3 grid(a.N / QROWS, a.H, a.B);
    tc_stats_kernel<BF16><<<grid, kThreads, ST_SMEM, s>>>(tm128, a);
    ET_COUNT_LAUNCH(1);
    if (mode == ET_ATTN_DELTA) {
        if ((rc = make_tmap_2d(&tmsel, sel, 3LL * a.sel_rows, (long long)a.D, 64, a.is_bf16))) return rc;
        tc_apply_kernel<BF16, ET_ATTN_DELTA><<<grid, kThreads, AP_SMEM, s>>>(tm128, tmsel, a);
    } else if (mode == ET_ATTN_FIRST) {
        tc_apply_kernel<BF16, ET_ATTN_FIRST><<<grid, kThreads, AP_SMEM, s>>>(tm128, tm64, a);
    } else {
        tc_apply_kernel<BF16, ET_ATTN_DENSE><<<grid, kThreads, AP_SMEM, s>>>(tm128, tm64, a);
    }
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}

}  // namespace

// Entry used by et_global_attention (et_attn.cu) when the shape qualifies for the tensor-core path.
// `sel` = workspace rows [K_sel | dV | v_n - dV], each (B * k, D); bias tables as produced by relpos_bias_kernel.
int et_tc_global_attention(const void* qkv, const void* sel, const void* bias_h, const void* bias_w, int mode,
                           const long long* idx, int k, void* a_state, void* acc, void* out, float* stats, int B, int N,
                           int NP, int H, int gh, int gw, int is_bf16, cudaStream_t stream) {
    TcArgs a;
    a.bias_h = bias_h; a.bias_w = bias_w; a.stats = stats; a.idx = idx; a.a_state = a_state; a.acc = acc; a.out = out;
    a.B = B; a.N = N; a.NP = NP; a.H = H; a.D = H * 64; a.gh = gh; a.gw = gw; a.k = k; a.is_bf16 = is_bf16;
    a.sel_rows = B * k;
    a.c1 = 0.125f * kLog2e;
    return is_bf16 ? launch_tc<true>(qkv, sel, a, mode, stream) : launch_tc<false>(qkv, sel, a, mode, stream);
}
