// Global attention of the Eventful blocks on 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
// Fast path for dh = 64, N % 128 == 0 and (no rel-pos | 64-wide token grid); everything else takes the
// mma.sync kernels in et_attn.cu.  Statistics are kept in the log2 domain: m2 = reference exponent of the row.
//
//   tc_stats_kernel  (phase A)  S = Q K^T per 128 x 128 tile in TMEM (double buffered); softmax warps read their
//                               row with tcgen05.ld, add the rel-pos bias (registers) and keep (m2, l) per row.
//   tc_apply_kernel  (phase B)  per 128-row query block and 64-key tile of the SELECTED keys:
//                               S' = Q' K'^T  with  Q' = [q | 8 bias_h(row) | 8 bias_w(row)],
//                                                   K' = [k | onehot(key y) | onehot(key x)]
//                                    so that S' / 8 = q.k / 8 + bias_h[row, ky] + bias_w[row, kx]: the decomposed
//                                    rel-pos bias (eventful_transformer/utils.py:157-166) comes out of the MMA and the
//                                    softmax threads do no table lookups;
//                               a_n = exp2(S' c1 - m2) / l, rounded to dtype exactly as stored in the gate state;
//                               A-gate + accumulator in one algebraic step.  The reference computes
//                                    acc += a_n . dV + (a_n - p) . (v_n - dV)            (modules.py:293-294)
//                                    and since dV + (v_n - dV) = v_n this equals
//                                    acc += a_n . v_n  -  p . (v_n - dV)
//                                    so the OLD state tile p = a_state[:, idx] is fed to the tensor core directly
//                                    (negated A operand, fp32 accumulate) and no thread ever reads it.  The state is
//                                    column-major: a selected column is 256 contiguous bytes per query block, moved
//                                    by 16-byte cp.async into an MN-major, 128B-swizzled A tile; a_n is written once
//                                    by the softmax threads into a tile of the same layout, which is both the A
//                                    operand of a_n . v_n and the source of the write-back a_state[:, idx] = a_n.
//                                    V tiles are MN-major B operands straight from TMA.
//                               epilogue: acc += O; out = acc.
//                    FIRST / DENSE modes run the same pipeline over all keys with V from the QKV buffer.
// Warp roles.  tc_stats (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator and single-thread MMA issuer,
// warps 2-9 = softmax: two warps per TMEM lane quarter, each thread owns one query row and half of the tile's key
// columns.  tc_apply (576 threads): the same plus eight state-mover warps (2, 3, 12-17) that own all A-gate state
// traffic; its softmax / epilogue warps are 4-11.
#include <type_traits>

#include "et_tcgen05.cuh"

using namespace et_tc;

// et_debug_set(6, 1): bracket the apply-kernel launch with CUDA events on its stream (bench.py reads the elapsed time
// of the last launch through et_debug_elapsed_ms()); never enabled inside graph capture.
int g_tc_time_apply = 0;
int g_tc_apply_cluster = 1;  // et_debug_set(12, n): tc_apply CTAs of n adjacent query blocks form a cluster (experiment)
// et_debug_set(4, device pointer to 8 x 16 u64): per-role cycle buckets, summed over all CTAs: rows 0-3 tc_apply_kernel
// (producer, MMA, softmax, mover), rows 4-6 tc_stats_kernel (producer, MMA, softmax).
// Only the profiling build (make prof: -DET_TC_PROFILE -> libeventful_b200_prof.so) writes to it.
unsigned long long* g_tc_prof = nullptr;
#ifdef ET_TC_PROFILE
#define PF_DECL long long pf_[16]; for (int i_ = 0; i_ < 16; ++i_) pf_[i_] = 0; long long pf_t_ = clock64();
#define PF(i) do { const long long n_ = clock64(); pf_[i] += n_ - pf_t_; pf_t_ = n_; } while (0)
#define PF_FLUSH(role) do { if (a.prof) for (int i_ = 0; i_ < 16; ++i_) atomicAdd(a.prof + (role) * 16 + i_, (unsigned long long)pf_[i_]); } while (0)
#else
#define PF_DECL
#define PF(i)
#define PF_FLUSH(role)
#endif
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

namespace {

constexpr int kThreads = 320;
constexpr int QROWS = 128;
constexpr float kLog2e = 1.4426950408889634f;

struct TcArgs {
    const void* bias_h;  // (B, H, N, 64): 8 x bias, zero padded to 64 columns (tc layout)
    const void* bias_w;
    float* stats;
    const long long* idx;
    void* a_state;
    void* acc;
    void* out;
    int B, N, NP, H, D, gh, gw, k, is_bf16, sel_rows, has_bias;
    float c1;  // (1 / sqrt(dh)) * log2(e)
    unsigned long long* prof;
};

template <bool BF16>
__device__ __forceinline__ float elem_to_float(uint16_t raw) {
    if constexpr (BF16) return __uint_as_float((uint32_t)raw << 16);
    else return __half2float(__ushort_as_half(raw));
}
template <bool BF16>
__device__ __forceinline__ uint32_t float_to_elem(float v) {
    if constexpr (BF16) return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    else return (uint32_t)__half_as_ushort(__float2half_rn(v));
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2_elem(float lo, float hi) {  // one F2FP; `lo` in bits 0-15
    uint32_t r;
    if constexpr (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_load_32x32(taddr, r); }
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_load_32x16(taddr, r); }

__device__ __forceinline__ void named_sync_softmax() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ============================================================================================= phase A
constexpr int ST_KEYS = 128;
constexpr int ST_STAGES = 4;
constexpr int ST_TILE = ST_KEYS * 64 * 2;  // 16 KB
constexpr int ST_PARTS = 4;                   // softmax warps per TMEM lane quarter (column parts of a 64-wide image row)
constexpr int ST_PW = 64 / ST_PARTS;          // columns per part
// Key pairs (of the 8 per thread and 64-key half tile) whose exp2 runs on the FMA pipe; -1 = the scalar MUFU-only form.
#ifndef ET_STATS_SW_PAIRS
#define ET_STATS_SW_PAIRS 0
#endif
constexpr int ST_SW_PAIRS = ET_STATS_SW_PAIRS;
// S tile buffers in tensor memory (128 columns each): 2 or 4
#ifndef ET_STATS_SBUFS
#define ET_STATS_SBUFS 2
#endif
constexpr int ST_SBUFS = ET_STATS_SBUFS;
constexpr int ST_SBUF_LOG = ST_SBUFS == 4 ? 2 : 1;
constexpr int kStThreads = 64 + ST_PARTS * 128;
// GEN = false: no rel-pos bias, or a 64-wide token grid (a 128-key tile = two image rows: the bias sits in registers).
// GEN = true : any grid up to 64 x 64: the bias comes out of the MMA as in tc_apply, S' = [q | 8 bias_h | 8 bias_w] .
//              [k | onehot(ky) | onehot(kx)]^T, three 64-column reduction blocks per operand (12 MMAs per tile instead of 4).
template <bool GEN>
struct StSmem {
    static constexpr int NKB = GEN ? 3 : 1;
    static constexpr int STAGES = GEN ? 3 : ST_STAGES;
    static constexpr int Q_BYTES = NKB * QROWS * 128;
    static constexpr int STAGE_BYTES = NKB * ST_TILE;
    static constexpr int OFF_X = Q_BYTES + STAGES * STAGE_BYTES;  // (m2, l) exchange between the column parts
    static constexpr int TOTAL = OFF_X + (ST_PARTS - 1) * QROWS * 8 + 256 + 1024;
    static_assert(TOTAL <= 227 * 1024, "tc_stats shared memory");
};

// RAGGED: the token count is not a multiple of 128 (last query block / key tile partly empty: masks and clamps compiled in)
template <bool BF16, bool GEN, bool RAGGED>
__global__ void __launch_bounds__(kStThreads, 1)
tc_stats_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_bh,
                const __grid_constant__ CUtensorMap tm_bw, const __grid_constant__ CUtensorMap tm_oh, const TcArgs a) {
    et_pdl_prologue();
    using L = StSmem<GEN>;
    constexpr int ST_STAGES = L::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Qs = smem;
    uint8_t* Ks = smem + L::Q_BYTES;
    float2* xchg = reinterpret_cast<float2*>(smem + L::OFF_X);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_X + (ST_PARTS - 1) * QROWS * 8);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + ST_STAGES;
    uint64_t* s_full = k_empty + ST_STAGES;
    uint64_t* s_empty = s_full + ST_SBUFS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + ST_SBUFS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * QROWS, h = blockIdx.y, b = blockIdx.z;
    const int T = (a.N + ST_KEYS - 1) / ST_KEYS;  // the last tile may be ragged: its columns >= N are masked below

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_qkv) : "memory");
        mbar_init(smem_u32(q_full), 1);
        for (int s = 0; s < ST_STAGES; ++s) {
            mbar_init(smem_u32(&k_full[s]), 1);
            mbar_init(smem_u32(&k_empty[s]), 1);
        }
        for (int s = 0; s < ST_SBUFS; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), ST_PARTS * 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), ST_SBUFS * ST_KEYS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            PF_DECL
            mbar_expect_tx(smem_u32(q_full), L::Q_BYTES);
            tma_load_2d(smem_u32(Qs), &tm_qkv, smem_u32(q_full), h * 64, b * a.N + q0);
            if (GEN) {
                const int brow = (b * a.H + h) * a.N + q0;
                tma_load_2d(smem_u32(Qs + QROWS * 128), &tm_bh, smem_u32(q_full), 0, brow);
                tma_load_2d(smem_u32(Qs + 2 * QROWS * 128), &tm_bw, smem_u32(q_full), 0, brow);
            }
            for (int t = 0; t < T; ++t) {
                const int s = t % ST_STAGES;
                PF(1);
                mbar_wait(smem_u32(&k_empty[s]), ((t / ST_STAGES) & 1) ^ 1);
                PF(0);
                uint8_t* dst = Ks + s * L::STAGE_BYTES;
                mbar_expect_tx(smem_u32(&k_full[s]), L::STAGE_BYTES);
                tma_load_2d(smem_u32(dst), &tm_qkv, smem_u32(&k_full[s]), a.D + h * 64, b * a.N + t * ST_KEYS);
                if (GEN) {  // one-hot coordinates of keys t * 128 ..: rows past N are zero-filled by the TMA unit
                    tma_load_2d(smem_u32(dst + ST_TILE), &tm_oh, smem_u32(&k_full[s]), 0, t * ST_KEYS);
                    tma_load_2d(smem_u32(dst + 2 * ST_TILE), &tm_oh, smem_u32(&k_full[s]), 64, t * ST_KEYS);
                }
            }
            PF(1);
            PF_FLUSH(4);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_ex(128, ST_KEYS, a.is_bf16, 0);
            PF_DECL
            mbar_wait(smem_u32(q_full), 0);
            PF(7);
            for (int t = 0; t < T; ++t) {
                const int s = t % ST_STAGES, u = t & (ST_SBUFS - 1);
                mbar_wait(smem_u32(&k_full[s]), (t / ST_STAGES) & 1);
                PF(0);
                mbar_wait(smem_u32(&s_empty[u]), ((t >> ST_SBUF_LOG) & 1) ^ 1);
                PF(1);
                tcgen05_fence_after();
#pragma unroll
                for (int kb = 0; kb < L::NKB; ++kb) {
                    const uint64_t dqb = umma_smem_desc(smem_u32(Qs + kb * QROWS * 128));
                    const uint64_t dk = umma_smem_desc(smem_u32(Ks + s * L::STAGE_BYTES + kb * ST_TILE));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tcgen05_mma_f16(tmem_base + u * ST_KEYS, dqb + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc, kb > 0 || kk > 0);
                }
                tcgen05_commit(smem_u32(&k_empty[s]));
                tcgen05_commit(smem_u32(&s_full[u]));
                PF(2);
            }
            PF_FLUSH(5);
        }
    } else {
        // ST_PARTS softmax warps per TMEM lane quarter: warp (quarter, part) owns key columns [ST_PW part, +ST_PW) of
        // each 64-wide image row.  Four parts = four softmax warps per scheduler: the exp2 / sum chains are latency-bound
        // with two (profiles/r1_tc_stats_roles.txt).
        const int quarter = warp & 3;
        const int part = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const size_t grow = ((size_t)b * a.H + h) * a.N + q0 + row;
        const size_t brow = RAGGED ? ((size_t)b * a.H + h) * a.N + min(q0 + row, a.N - 1) : grow;  // clamped in a ragged query block
        // rel-pos bias of this query row (stored as 8 x bias), in the log2 domain
        const float bscale = 0.125f * kLog2e;
        float bwl[ST_PW];
        const uint16_t* bh_row = nullptr;
        if (a.has_bias && !GEN) {
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(a.bias_w) + brow * 64 + part * ST_PW);
#pragma unroll
            for (int c = 0; c < ST_PW / 8; ++c) {
                const uint4 u4 = src[c];
                const uint32_t w[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    bwl[c * 8 + 2 * i] = elem_to_float<BF16>((uint16_t)(w[i] & 0xffffu)) * bscale;
                    bwl[c * 8 + 2 * i + 1] = elem_to_float<BF16>((uint16_t)(w[i] >> 16)) * bscale;
                }
            }
            bh_row = static_cast<const uint16_t*>(a.bias_h) + brow * 64;
        } else {
#pragma unroll
            for (int i = 0; i < ST_PW; ++i) bwl[i] = 0.f;
        }
        float m2 = -1e30f, l = 0.f;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // the row's bias_h pair of the NEXT tile is fetched one tile ahead: loaded at the point of use, its L2 round trip was
        // 20 % of the kernel's stall samples (gpurun_out/r2_stats.ncu-rep, long_scoreboard at the conversion below)
        uint32_t pair_next = (a.has_bias && !GEN) ? *reinterpret_cast<const uint32_t*>(bh_row) : 0u;
        PF_DECL
        for (int t = 0; t < T; ++t) {
            const int u = t & (ST_SBUFS - 1);
            PF(3);
            float bh2[2] = {0.f, 0.f};
            if (a.has_bias && !GEN) {  // a 128-key tile spans two image rows of the 64-wide grid
                const uint32_t pair = pair_next;
                if (t + 1 < T) pair_next = *reinterpret_cast<const uint32_t*>(bh_row + 2 * (t + 1));
                bh2[0] = elem_to_float<BF16>((uint16_t)(pair & 0xffffu)) * bscale;
                bh2[1] = elem_to_float<BF16>((uint16_t)(pair >> 16)) * bscale;
            }
            mbar_wait(smem_u32(&s_full[u]), (t >> ST_SBUF_LOG) & 1);
            PF(0);
            tcgen05_fence_after();
            uint32_t v0[ST_PW], v1[ST_PW];
            tmem_load_cols(taddr + (uint32_t)(u * ST_KEYS + part * ST_PW), v0);       // image row 2t
            tmem_load_cols(taddr + (uint32_t)(u * ST_KEYS + 64 + part * ST_PW), v1);  // image row 2t + 1
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_empty[u]));
            PF(1);
            // one 64-key half of the tile; MASKED: the ragged last tile of a token count that is not a multiple of 128
            auto half_tile = [&](auto masked_tag, const uint32_t* v, float bh, int key0) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                if constexpr (!MASKED && ST_SW_PAIRS >= 0) {
                    // Packed fp32 arithmetic (two keys per FFMA2 / FADD2 issue slot), and exp2 of ST_SW_PAIRS of the eight key
                    // pairs computed on the FMA pipe instead of the MUFU: the statistics pass is bound by the 16 exp2 / clk / SM
                    // of the MUFU (1 024 cycles per 128 x 128 tile), while the FMA pipe has issue slots to spare.
                    //   2^x = 2^n 2^r, n = round(x) by the magic-number add, r = x - n in [-1/2, 1/2], 2^r by a cubic (relative
                    //   error < 7.5e-5, only the row sums l are formed from these values), 2^n by an integer add to the exponent.
                    f32x2 xp[ST_PW / 2];
                    float cmax = -1e30f;
                    const f32x2 c1c1 = f2_pack(a.c1, a.c1);
#pragma unroll
                    for (int j = 0; j < ST_PW / 2; ++j) {
                        xp[j] = f2_fma(f2_pack(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), c1c1,
                                       f2_pack(bwl[2 * j], bwl[2 * j + 1]));
                        float x0, x1;
                        f2_unpack(xp[j], x0, x1);
                        cmax = fmaxf(cmax, fmaxf(x0, x1));
                    }
                    if (cmax + bh > m2 + 8.f) {
                        l *= ex2_approx(m2 - (cmax + bh));
                        m2 = cmax + bh;
                    }
                    const float shift = bh - m2;
                    const f32x2 shift2 = f2_pack(shift, shift);
                    f32x2 sum2 = f2_pack(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < ST_PW / 2; ++j) {
                        const f32x2 xs = f2_add(xp[j], shift2);
                        float x0, x1, e0, e1;
                        f2_unpack(xs, x0, x1);
                        if (j < ST_SW_PAIRS) {
                            const f32x2 xc = f2_pack(fmaxf(x0, -125.f), fmaxf(x1, -125.f));  // keep 2^n a normal number
                            const f32x2 magic = f2_pack(12582912.f, 12582912.f);             // 1.5 * 2^23: x + magic rounds x to an integer
                            const f32x2 t = f2_add(xc, magic);
                            const f32x2 nf = f2_add(t, f2_pack(-12582912.f, -12582912.f));
                            const f32x2 r = f2_fma(nf, f2_pack(-1.f, -1.f), xc);
                            f32x2 p = f2_fma(f2_pack(0.055171654f, 0.055171654f), r, f2_pack(0.24261113f, 0.24261113f));
                            p = f2_fma(p, r, f2_pack(0.69326097f, 0.69326097f));
                            p = f2_fma(p, r, f2_pack(0.99992806f, 0.99992806f));
                            float t0, t1, p0, p1;
                            f2_unpack(t, t0, t1);
                            f2_unpack(p, p0, p1);
                            e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
                            e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
                        } else {
                            e0 = ex2_approx(x0);
                            e1 = ex2_approx(x1);
                        }
                        sum2 = f2_add(sum2, f2_pack(e0, e1));
                    }
                    float s0, s1;
                    f2_unpack(sum2, s0, s1);
                    l += s0 + s1;
                    return;
                }
                float x[ST_PW];
                float cmax = -1e30f;
#pragma unroll
                for (int i = 0; i < ST_PW; ++i) {
                    x[i] = fmaf(__uint_as_float(v[i]), a.c1, bwl[i]);
                    if (MASKED && key0 + i >= a.N) x[i] = -1e30f;
                    cmax = fmaxf(cmax, x[i]);
                }
                // m2 is a reference exponent, not necessarily the exact row max: it only moves when exceeded by more
                // than 2^8, so exp2(x - m2) <= 256 stays finite and (m2, l) remain a consistent pair
                if (cmax + bh > m2 + 8.f) {
                    l *= ex2_approx(m2 - (cmax + bh));
                    m2 = cmax + bh;
                }
                const float shift = bh - m2;
                float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
                for (int i = 0; i < ST_PW; i += 2) {
                    const float e0 = ex2_approx(x[i] + shift), e1 = ex2_approx(x[i + 1] + shift);
                    // masked columns add exactly nothing (x + shift cancels to a finite number while m2 is still -1e30)
                    sum0 += (MASKED && key0 + i >= a.N) ? 0.f : e0;
                    sum1 += (MASKED && key0 + i + 1 >= a.N) ? 0.f : e1;
                }
                l += sum0 + sum1;
            };
            const int key0 = t * ST_KEYS + part * ST_PW;  // key of v0[0]; v1[0] is 64 keys later
            if (RAGGED && (t + 1) * ST_KEYS > a.N) {  // warp-uniform
                half_tile(std::true_type{}, v0, bh2[0], key0);
                half_tile(std::true_type{}, v1, bh2[1], key0 + 64);
            } else {
                half_tile(std::false_type{}, v0, bh2[0], key0);
                half_tile(std::false_type{}, v1, bh2[1], key0 + 64);
            }
            PF(2);
        }
        PF(3);
        if (lane == 0 && warp == 2) PF_FLUSH(6);
        // merge the column parts of each row
        if (part > 0) xchg[(part - 1) * QROWS + row] = make_float2(m2, l);
        asm volatile("bar.sync 1, %0;" ::"n"(ST_PARTS * 128) : "memory");
        if (part == 0) {
            float mm = m2;
#pragma unroll
            for (int o = 0; o < ST_PARTS - 1; ++o) mm = fmaxf(mm, xchg[o * QROWS + row].x);
            l *= ex2_approx(m2 - mm);
#pragma unroll
            for (int o = 0; o < ST_PARTS - 1; ++o) {
                const float2 ov = xchg[o * QROWS + row];
                l += ov.y * ex2_approx(ov.x - mm);
            }
            if (!RAGGED || q0 + row < a.N) {  // rows past N of a ragged query block belong to nobody
                a.stats[grow * 2] = mm;
                a.stats[grow * 2 + 1] = l;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, ST_SBUFS * ST_KEYS);
    }
}

// ============================================================================================= phase B
// State-mover warps of tc_apply.  8 (default): warps 2-3 and 12-17, 640 threads.  16 (-DET_APPLY_MOVERS=16): warps 12-27,
// 1 024 threads, register file re-divided per warpgroup by setmaxnreg (softmax 104, movers 48, the rest 40).
#ifndef ET_APPLY_MOVERS
#define ET_APPLY_MOVERS 8
#endif
constexpr int kMoverWarps = ET_APPLY_MOVERS;
constexpr bool kWideMovers = kMoverWarps == 16;
static_assert(kMoverWarps == 8 || kMoverWarps == 16, "8 or 16 state-mover warps");
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// warp 0 TMA, 1 PV-MMA issuer, 4-11 softmax, movers, and one more warp that issues the S' MMAs: a tcgen05.mma occupies its
// issuing thread ~85 cycles, and with one issuer the 12 S' MMAs of a tile pair sat between the two PV sets of the pair on the
// critical path (softmax publish -> PV -> pv_done -> next publish); S' depends only on K' and a free S buffer
constexpr int kSIssueWarp = kWideMovers ? 28 : 10 + kMoverWarps;
// ... and a third issuer for the -p (v_n - dV) MMAs (DELTA mode): they need only the old state tile and V, not the softmax,
// so they run ahead of the publish -> a_n v_n -> pv_done chain into a second accumulator that the epilogue adds
constexpr int kPIssueWarp = kSIssueWarp + 1;
constexpr int kApThreads = kWideMovers ? 1024 : (kPIssueWarp + 1) * 32;
constexpr int MV_CPT = 1024 / (kMoverWarps * 32);    // 16-byte chunks per mover thread and tile
constexpr int MV_CSTEP = kMoverWarps * 2;            // columns covered by one pass of the mover threads
constexpr int AP_KEYS = 64;
constexpr int AP_BLK = AP_KEYS * 64 * 2;          // 8 KB: one 64-key x 64-column operand block
constexpr int AP_STAGE = 5 * AP_BLK;              // K, onehot-y, onehot-x, V1, V2
constexpr int AP_PT = AP_KEYS * QROWS * 2;        // 16 KB: a_state tile [key][row]
constexpr int AP_P = QROWS * AP_KEYS * 2;         // 16 KB: a_n tile (A operand, MN-major) = write-back staging
constexpr int AP_OFF_ST = 3 * QROWS * 128;        // Q' = three 16 KB blocks: q, 8 bias_h, 8 bias_w
constexpr int AP_OFF_PT = AP_OFF_ST + 2 * AP_STAGE;
// a_n tile buffers: 1 = softmax(t+1) publishes only after PV(t) has retired (ping-pong through one buffer); 2 = it may run one
// tile ahead (the old-state ring gives up one of its four stages to make room) -- build-time experiment switch
#ifndef ET_APPLY_AN_BUFS
#define ET_APPLY_AN_BUFS 1
#endif
constexpr int AN_BUFS = ET_APPLY_AN_BUFS;
// Two buffers were measured in round 2 (no gain: profiles/r2_experiments.md) with the single-issuer kernel.  With the softmax a
// tile further ahead, pv_done[u] can complete twice before a mover has observed the first completion (parity aliasing), so
// the barrier protocol below is only valid for one buffer.
static_assert(AN_BUFS == 1, "tc_apply's pv_done protocol assumes a single a_n buffer");
#ifndef ET_APPLY_PT_STAGES
#define ET_APPLY_PT_STAGES (ET_APPLY_AN_BUFS == 1 ? 4 : 3)
#endif
constexpr int AP_PT_STAGES = ET_APPLY_PT_STAGES;  // a_state tiles are prefetched this many tiles ahead by the mover warps
constexpr int AP_OFF_P = AP_OFF_PT + AP_PT_STAGES * AP_PT;
constexpr int AP_OFF_MISC = AP_OFF_P + AN_BUFS * AP_P;
constexpr int AP_OFF_IDX = AP_OFF_MISC + 1024;    // int32 copy of this batch entry's selected-key index (DELTA mode, k <= AP_IDX_MAX)
#ifndef ET_APPLY_IDX_MAX
#define ET_APPLY_IDX_MAX 3584
#endif
constexpr int AP_IDX_MAX = ET_APPLY_IDX_MAX;      // 14 KB: what is left of the 227 KB
constexpr int AP_SMEM = AP_OFF_IDX + AP_IDX_MAX * 4 + 1024;
static_assert(AP_SMEM <= 227 * 1024, "tc_apply shared memory");

template <bool BF16, int MODE, bool RAGGED>
__global__ void __launch_bounds__(kApThreads, 1)
tc_apply_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bw,
                const __grid_constant__ CUtensorMap tm_oh, const TcArgs a) {
    et_pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    auto Qb = [&](int i) { return smem + i * QROWS * 128; };
    // K' blocks (i = 0 k, 1 onehot-y, 2 onehot-x) of the two 64-key tiles u = 0, 1 of a PAIR sit next to each other, so that one
    // N = 128 MMA covers both tiles: a tcgen05.mma costs ~90 cycles whether N is 64 or 128 (profiles/r1_mma_rate_microbench.txt),
    // and S' was 12 of the 20 MMAs per tile.  V blocks keep a two-deep ring per 64-key tile.
    auto Kb = [&](int u, int i) { return smem + AP_OFF_ST + (2 * i + u) * AP_BLK; };
    auto V1 = [&](int u) { return smem + AP_OFF_ST + (6 + 2 * u) * AP_BLK; };
    auto V2 = [&](int u) { return smem + AP_OFF_ST + (7 + 2 * u) * AP_BLK; };
    // MN-major A tiles [64 keys][128 rows]: two 8 KB blocks of 64 rows; key kk = one 128-byte line per block,
    // its 16-byte chunk c (8 rows) stored at chunk position c ^ (kk & 7)
    auto Pt = [&](int u) { return smem + AP_OFF_PT + u * AP_PT; };
    auto An = [&](int t) { return smem + AP_OFF_P + (t % AN_BUFS) * AP_P; };  // a_n tile of key tile t
    auto a_chunk = [](int key, int seg) { return (seg >> 3) * 8192 + key * 128 + (((seg & 7) ^ (key & 7)) << 4); };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AP_OFF_MISC + 512);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;    // [0]: K' operand blocks of a tile pair landed
    uint64_t* ps_full = bars + 3;   // [4]: a_state tile landed (128 cp.async arrivals, one per mover thread)
    uint64_t* s_full = bars + 7;
    uint64_t* s_empty = bars + 9;
    uint64_t* p_ready = bars + 20;  // [AN_BUFS]: a_n tile written (8 softmax warps arrive), one phase per use of the buffer
    uint64_t* pv_done = bars + 12;  // [2]: PV MMAs of tile t commit to slot t & 1 (at most one phase outstanding each)
    uint64_t* o_full = bars + 14;
    uint64_t* an_free = bars + 22;  // [AN_BUFS]: a_n tile copied out by the mover warps
    uint64_t* v_full = bars + 16;   // [2]: V blocks of a stage landed (released by pv_done)
    uint64_t* k_empty = bars + 18;  // [0]: S' MMAs of the pair are complete -> the K' blocks may be reloaded
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * QROWS, h = blockIdx.y, b = blockIdx.z;
    const int nkeys = (MODE == ET_ATTN_DELTA) ? a.k : a.N;
    const int T = (nkeys + AP_KEYS - 1) / AP_KEYS;
    const int nkb = a.has_bias ? 3 : 1;  // 64-column blocks of the augmented reduction dimension
    uint16_t* a_state = static_cast<uint16_t*>(a.a_state);
    const size_t a_head = ((size_t)b * a.H + h) * (size_t)a.N * a.NP;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_kv) : "memory");
        mbar_init(smem_u32(q_full), 1);
        mbar_init(smem_u32(o_full), MODE == ET_ATTN_DELTA ? 2 : 1);  // both PV issuers commit in DELTA mode
        for (int u = 0; u < AN_BUFS; ++u) {
            mbar_init(smem_u32(&p_ready[u]), 8);
            mbar_init(smem_u32(&an_free[u]), kMoverWarps);
        }
        for (int u = 0; u < 2; ++u) {
            mbar_init(smem_u32(&k_full[u]), 1);
            mbar_init(smem_u32(&v_full[u]), 1);
            mbar_init(smem_u32(&k_empty[u]), 1);
            mbar_init(smem_u32(&s_full[u]), 1);
            mbar_init(smem_u32(&s_empty[u]), 16);  // a pair buffer is drained by 8 softmax warps x 2 tiles
            mbar_init(smem_u32(&pv_done[u]), MODE == ET_ATTN_DELTA ? 2 : 1);
        }
        for (int u = 0; u < AP_PT_STAGES; ++u) mbar_init(smem_u32(&ps_full[u]), kMoverWarps * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);  // S' pair buffers 2 x 128 columns, O 64 columns
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 4 * AP_KEYS;
    auto state_movers = [&]() {
        // ------------------------------------------------------------------ A-gate state movers
        // A selected column x this CTA's 128 rows is 256 contiguous bytes of the column-major state = 16 threads x 16 B;
        // thread mt moves segment (mt & 15) of columns (mt >> 4) + MV_CSTEP i, i < MV_CPT.  Per tile: prefetch the old state tile
        // AP_PT_STAGES tiles ahead (cp.async straight into the MN-major A-operand layout) and write the new a_n tile back
        // (modules.py:200).  Scattered 256-byte HBM segments stall the issuing warps (LSU back-pressure), so this traffic
        // has its own warps and never holds up the exp / MMA pipeline.
        if (MODE != ET_ATTN_DENSE) {
            const int mt = (kWideMovers ? warp - 12 : (warp < 4 ? warp - 2 : warp - 10)) * 32 + lane;
            const int segi = mt & 15, seg = segi * 8, col0 = mt >> 4;
            const bool row_chunk_ok = !RAGGED || q0 + seg < a.NP;  // ragged last query block: 8-row chunks past the column's NP rows do not exist
            // the index of this batch entry goes to shared memory once: every tile needs it twice per thread (prefetch and
            // write-back), and an L2 round trip per tile was 15 % of the movers' time (profiles/r1_tc_apply_roles.txt)
            int* s_idx = reinterpret_cast<int*>(smem + AP_OFF_IDX);
            const bool idx_in_smem = MODE == ET_ATTN_DELTA && nkeys <= AP_IDX_MAX;
            if (idx_in_smem) {
                for (int j = mt; j < nkeys; j += kMoverWarps * 32) s_idx[j] = (int)a.idx[(size_t)b * a.k + j];
                asm volatile("bar.sync 2, %0;" ::"n"(kMoverWarps * 32) : "memory");
            }
            auto tile_tok = [&](int tt, int i) -> int {  // token of column col0 + 8 i of tile tt (or -1)
                const int j = tt * AP_KEYS + col0 + MV_CSTEP * i;
                if (j >= nkeys) return -1;
                if (MODE != ET_ATTN_DELTA) return j;
                return idx_in_smem ? s_idx[j] : (int)a.idx[(size_t)b * a.k + j];
            };
            auto load_state = [&](int tt, const int (&tok)[MV_CPT]) {  // a_state[:, idx of tile tt] -> Pt ring (A-operand layout)
                uint8_t* dst = Pt(tt % AP_PT_STAGES);
#pragma unroll
                for (int i = 0; i < MV_CPT; ++i) {
                    uint8_t* d = dst + a_chunk(col0 + MV_CSTEP * i, segi);
                    if (tok[i] >= 0 && row_chunk_ok) cp_async_16(smem_u32(d), a_state + a_head + (size_t)tok[i] * a.NP + q0 + seg);
                    else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);  // ragged tile: p = 0, never garbage
                }
                cp_async_arrive_noinc(smem_u32(&ps_full[tt % AP_PT_STAGES]));
            };
            PF_DECL
            int tok_next[MV_CPT];
            if (MODE == ET_ATTN_DELTA) {
                for (int tt = 0; tt < AP_PT_STAGES && tt < T; ++tt) {
#pragma unroll
                    for (int i = 0; i < MV_CPT; ++i) tok_next[i] = tile_tok(tt, i);
                    load_state(tt, tok_next);
                }
#pragma unroll
                for (int i = 0; i < MV_CPT; ++i) tok_next[i] = AP_PT_STAGES < T ? tile_tok(AP_PT_STAGES, i) : -1;
            }
            for (int t = 0; t < T; ++t) {
                int tok_wb[MV_CPT];
#pragma unroll
                for (int i = 0; i < MV_CPT; ++i) tok_wb[i] = tile_tok(t, i);
                PF(0);
                if (MODE == ET_ATTN_DELTA && t >= 1 && t - 1 + AP_PT_STAGES < T) {
                    // ring slot of tile t-1 is free once its PV MMAs are done
                    mbar_wait(smem_u32(&pv_done[(t - 1) & 1]), ((t - 1) >> 1) & 1);
                    PF(1);
                    load_state(t - 1 + AP_PT_STAGES, tok_next);
                    if (t + AP_PT_STAGES < T) {
#pragma unroll
                        for (int i = 0; i < MV_CPT; ++i) tok_next[i] = tile_tok(t + AP_PT_STAGES, i);
                    }
                    PF(2);
                }
                mbar_wait(smem_u32(&p_ready[t % AN_BUFS]), (t / AN_BUFS) & 1);  // every row of the a_n tile is written
                PF(3);
                uint4 wb[MV_CPT];
#pragma unroll
                for (int i = 0; i < MV_CPT; ++i) wb[i] = *reinterpret_cast<const uint4*>(An(t) + a_chunk(col0 + MV_CSTEP * i, segi));
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&an_free[t % AN_BUFS]));  // all chunks read: the a_n tile may be rewritten
                PF(4);
#pragma unroll
                for (int i = 0; i < MV_CPT; ++i)
                    if (tok_wb[i] >= 0 && row_chunk_ok)  // evict-first: the 400 MB of state columns stream through L2 once per frame
                        asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a_state + a_head + (size_t)tok_wb[i] * a.NP + q0 + seg),
                                     "r"(wb[i].x), "r"(wb[i].y), "r"(wb[i].z), "r"(wb[i].w) : "memory");
                PF(5);
            }
            if (warp == 2 && lane == 0) PF_FLUSH(3);
        }
    };

    auto role_producer = [&]() {
        // ------------------------------------------------------------------ producer + state write-back
        if (lane == 0) {
            const int qrow = b * a.N + q0;
            mbar_expect_tx(smem_u32(q_full), nkb * QROWS * 128);
            tma_load_2d(smem_u32(Qb(0)), &tm_q, smem_u32(q_full), h * 64, qrow);
            if (a.has_bias) {
                const int brow = (b * a.H + h) * a.N + q0;
                tma_load_2d(smem_u32(Qb(1)), &tm_bh, smem_u32(q_full), 0, brow);
                tma_load_2d(smem_u32(Qb(2)), &tm_bw, smem_u32(q_full), 0, brow);
            }
        }
        // K' blocks and V blocks of a stage are separate transactions: K'(t) is reloadable as soon as S'(t-2) has been
        // computed (a full tile before the PV MMAs of t-2 finish), so the S' MMAs - which run one tile ahead of the PV
        // MMAs - never wait for the TMA round trip.  Issue order: K'(0) K'(1) V(0) K'(2) V(1) ...
        PF_DECL
        if (lane == 0) {
            for (int i = 0; i <= T; ++i) {
                if (i < T && (i & 1) == 0) {  // K' of the pair (i, i + 1); a missing second tile loads rows nobody reads
                    const int pr = i >> 1;
                    PF(1);
                    if (pr >= 1) mbar_wait(smem_u32(&k_empty[0]), (pr & 1) ^ 1);
                    PF(0);
                    const uint32_t fb = smem_u32(&k_full[0]);
                    mbar_expect_tx(fb, 2 * nkb * AP_BLK);
                    for (int u = 0; u < 2; ++u) {
                        const int key0 = (i + u) * AP_KEYS;
                        if (MODE == ET_ATTN_DELTA) tma_load_2d(smem_u32(Kb(u, 0)), &tm_kv, fb, h * 64, b * a.k + key0);
                        else tma_load_2d(smem_u32(Kb(u, 0)), &tm_kv, fb, a.D + h * 64, b * a.N + key0);
                        if (a.has_bias) {
                            const int orow = (MODE == ET_ATTN_DELTA) ? b * a.k + key0 : key0;
                            tma_load_2d(smem_u32(Kb(u, 1)), &tm_oh, fb, 0, orow);
                            tma_load_2d(smem_u32(Kb(u, 2)), &tm_oh, fb, 64, orow);
                        }
                    }
                }
                if (i >= 1) {
                    const int j = i - 1, u = j & 1, key0 = j * AP_KEYS;
                    PF(1);
                    if (j >= 2) mbar_wait(smem_u32(&pv_done[u]), ((j >> 1) & 1) ^ 1);  // tile j-2: the PV MMAs are done with slot u
                    PF(2);
                    const uint32_t fb = smem_u32(&v_full[u]);
                    mbar_expect_tx(fb, (MODE == ET_ATTN_DELTA ? 2 : 1) * AP_BLK);
                    if (MODE == ET_ATTN_DELTA) {
                        const int r0 = b * a.k + key0;
                        tma_load_2d(smem_u32(V1(u)), &tm_kv, fb, h * 64, a.sel_rows + r0);
                        tma_load_2d(smem_u32(V2(u)), &tm_kv, fb, h * 64, 2 * a.sel_rows + r0);
                    } else {
                        tma_load_2d(smem_u32(V1(u)), &tm_kv, fb, 2 * a.D + h * 64, b * a.N + key0);
                    }
                }
            }
        }
        PF(1);
        if (lane == 0) PF_FLUSH(0);
    };
    auto role_an_issuer = [&]() {
        // ------------------------------------------------------------------ a_n . v_n MMA issuer
        if (lane == 0) {
            // PV products: A = [key][row] tiles (MN-major), B = V tiles (MN-major); the p . Vd product is subtracted
            const uint32_t idesc_o = umma_idesc_ex(128, 64, a.is_bf16, 1) | (1u << 15);
            PF_DECL
            mbar_wait(smem_u32(q_full), 0);
            PF(7);
            auto issue_pv = [&](int t) {
                const int u = t & 1;
                mbar_wait(smem_u32(&p_ready[t % AN_BUFS]), (t / AN_BUFS) & 1);
                PF(3);
                mbar_wait(smem_u32(&v_full[u]), (t >> 1) & 1);
                PF(8);
                tcgen05_fence_after();
                const uint64_t dan = umma_smem_desc_mn_a(smem_u32(An(t)));
                const uint64_t dv1 = umma_smem_desc_mn(smem_u32(V1(u)));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)  // 16 keys per step = 16 lines of 128 B (2048 B) in the A and V tiles
                    tcgen05_mma_f16(tmem_o, dan + (uint64_t)(128 * kk), dv1 + (uint64_t)(128 * kk), idesc_o, (t > 0 || kk > 0));
                PF(4);
                tcgen05_commit(smem_u32(&pv_done[u]));
                if (t == T - 1) tcgen05_commit(smem_u32(o_full));
                PF(6);
            };
            for (int t = 0; t < T; ++t) issue_pv(t);
            PF_FLUSH(1);
        }
    };
    auto role_softmax = [&]() {
        // ------------------------------------------------------------------ softmax / epilogue
        const int quarter = warp & 3;
        const int half = (warp - 4) >> 2;  // key columns [32 half, +32) of every 64-key tile
        const int row = quarter * 32 + lane;
        const size_t grow = ((size_t)b * a.H + h) * a.N + q0 + row;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // The S' tile is read in mma-fragment shape (tcgen05.ld 16x256b): this thread owns, for a < 4, query row
        // quarter * 32 + 8 a + lane / 4 and key columns 8 n + 2 (lane % 4), + 1 (n < 4) of its 32-key half.  After
        // exp / normalise / pack the registers are 8x8 b16 fragments, and stmatrix.trans writes each as eight 16-byte
        // chunks (one key x 8 rows): exactly the MN-major, 128B-swizzled A-operand tile, 4 stores per thread and tile.
        const int lr = lane >> 2, lc = (lane & 3) * 2;
        float m2r[4], linvr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int qr = q0 + quarter * 32 + 8 * i + lr;
            const size_t g = ((size_t)b * a.H + h) * a.N + (RAGGED ? min(qr, a.N - 1) : qr);
            m2r[i] = a.stats[g * 2];
            linvr[i] = 1.f / a.stats[g * 2 + 1];
        }
        // stmatrix: lane 8 g + j addresses key (32 half + 8 n + j), rows quarter * 32 + 8 g ... + 7
        const int sm_r = quarter * 32 + 8 * (lane >> 3), sm_j = lane & 7;
        const uint32_t st_addr0 = smem_u32(smem + AP_OFF_P) + (sm_r >> 6) * 8192 + (half * 32 + sm_j) * 128 + ((((sm_r >> 3) & 7) ^ sm_j) << 4);
        // previous accumulator values of this thread's output slice: loaded now, consumed in the epilogue
        uint16_t* acc = static_cast<uint16_t*>(a.acc);
        uint16_t* out = static_cast<uint16_t*>(a.out);
        const size_t off = ((size_t)b * a.N + q0 + row) * a.D + h * 64 + half * 32;
        const bool row_ok = !RAGGED || q0 + row < a.N;  // ragged last query block
        uint4 prev[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            prev[c] = (MODE == ET_ATTN_DELTA && row_ok) ? *reinterpret_cast<const uint4*>(acc + off + c * 8) : make_uint4(0, 0, 0, 0);
        PF_DECL
        for (int t = 0; t < T; ++t) {
            const int u = (t >> 1) & 1;                      // pair buffer
            const uint32_t ph = (t >> 2) & 1;
            const uint32_t scol = (uint32_t)(u * 2 * AP_KEYS + (t & 1) * AP_KEYS + half * 32);
            PF(11);
            mbar_wait(smem_u32(&s_full[u]), ph);
            PF(0);
            tcgen05_fence_after();
            uint32_t v0[16], v1[16];  // lanes 0-15 / 16-31 of this warp's TMEM quarter
            tmem_load_16x256b_x4(taddr + scol, v0);
            tmem_load_16x256b_x4(taddr + (16u << 16) + scol, v1);
            tmem_wait_ld();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_empty[u]));
            PF(1);
            const bool full_tile = (t + 1) * AP_KEYS <= nkeys;
            // normalised attention values of the selected columns, rounded to dtype exactly as stored in the state
            uint32_t fr[4][4];  // [key group n][row block i]: keys 8 n + lc (low half), + 1 (high half)
#pragma unroll
            for (int n = 0; n < 4; ++n) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t* src = i < 2 ? v0 : v1;
                    const int e = 4 * n + 2 * (i & 1);
                    const float p0 = ex2_approx(fmaf(__uint_as_float(src[e]), a.c1, -m2r[i])) * linvr[i];
                    const float p1 = ex2_approx(fmaf(__uint_as_float(src[e + 1]), a.c1, -m2r[i])) * linvr[i];
                    fr[n][i] = pack2_elem<BF16>(p0, p1);
                }
            }
            if (!full_tile) {  // ragged last tile: keys beyond k contribute nothing
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const int j = t * AP_KEYS + half * 32 + 8 * n + lc;
                    const uint32_t mask = (j >= nkeys ? 0u : 0x0000ffffu) | (j + 1 >= nkeys ? 0u : 0xffff0000u);
#pragma unroll
                    for (int i = 0; i < 4; ++i) fr[n][i] &= mask;
                }
            }
            PF(2);
            if (t >= AN_BUFS) {
                // the previous tile in this a_n buffer has been consumed by the PV MMAs and copied out by the write-back warps
                const int tp = t - AN_BUFS;
                mbar_wait(smem_u32(&pv_done[tp & 1]), (tp >> 1) & 1);
                if (MODE != ET_ATTN_DENSE) mbar_wait(smem_u32(&an_free[tp % AN_BUFS]), (tp / AN_BUFS) & 1);
            }
            PF(3);
#pragma unroll
            for (int n = 0; n < 4; ++n)
                stmatrix_x4_trans(st_addr0 + (t % AN_BUFS) * AP_P + n * 1024, fr[n][0], fr[n][1], fr[n][2], fr[n][3]);
            PF(4);
            fence_proxy_async();  // generic-proxy smem writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&p_ready[t % AN_BUFS]));
            PF(5);
        }
        PF(11);
        // ---- epilogue: acc += O, out = acc ; each thread writes its row's 32 of the head's 64 output columns
        uint32_t o[32];
        if (T > 0) {
            mbar_wait(smem_u32(o_full), 0);
            PF(12);
            tcgen05_fence_after();
            tmem_load_32x32(taddr + (uint32_t)(4 * AP_KEYS + half * 32), o);
            if (MODE == ET_ATTN_DELTA) {  // second accumulator: - p (v_n - dV)
                uint32_t o2[32];
                tmem_load_32x32(taddr + (uint32_t)(4 * AP_KEYS + 64 + half * 32), o2);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) + __uint_as_float(o2[i]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
        PF(13);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
            const uint32_t pw[4] = {prev[c].x, prev[c].y, prev[c].z, prev[c].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float lo = __uint_as_float(o[c * 8 + 2 * i]), hi = __uint_as_float(o[c * 8 + 2 * i + 1]);
                if (MODE == ET_ATTN_DELTA) {
                    lo += elem_to_float<BF16>((uint16_t)(pw[i] & 0xffffu));
                    hi += elem_to_float<BF16>((uint16_t)(pw[i] >> 16));
                }
                w[i] = pack2_elem<BF16>(lo, hi);
            }
            const uint4 pk = make_uint4(w[0], w[1], w[2], w[3]);
            if (row_ok) {
                if (MODE != ET_ATTN_DENSE) *reinterpret_cast<uint4*>(acc + off + c * 8) = pk;
                *reinterpret_cast<uint4*>(out + off + c * 8) = pk;
            }
        }
        PF(14);
        if (lane == 0 && warp == 4) PF_FLUSH(2);
    };
    auto role_p_issuer = [&]() {
        // ------------------------------------------------------------------ -p (v_n - dV) MMA issuer (DELTA mode)
        if (MODE == ET_ATTN_DELTA && lane == 0) {
            const uint32_t idesc_neg = umma_idesc_ex(128, 64, a.is_bf16, 1) | (1u << 15) | (1u << 13);  // MN-major A and B, a_negate
            PF_DECL
            for (int t = 0; t < T; ++t) {
                const int u = t & 1;
                PF(2);
                mbar_wait(smem_u32(&v_full[u]), (t >> 1) & 1);
                PF(0);
                mbar_wait(smem_u32(&ps_full[t % AP_PT_STAGES]), (t / AP_PT_STAGES) & 1);  // old state tile landed
                PF(1);
                tcgen05_fence_after();
                fence_proxy_async();  // cp.async (generic proxy) writes -> visible to the tensor core
                const uint64_t dp = umma_smem_desc_mn_a(smem_u32(Pt(t % AP_PT_STAGES)));
                const uint64_t dv2 = umma_smem_desc_mn(smem_u32(V2(u)));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tcgen05_mma_f16(tmem_o + 64, dp + (uint64_t)(128 * kk), dv2 + (uint64_t)(128 * kk), idesc_neg, (t > 0 || kk > 0));
                tcgen05_commit(smem_u32(&pv_done[u]));
                if (t == T - 1) tcgen05_commit(smem_u32(o_full));
            }
            PF(2);
            PF_FLUSH(9);
        }
    };
    auto role_s_issuer = [&]() {
        // ------------------------------------------------------------------ S' MMA issuer
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_ex(128, 2 * AP_KEYS, a.is_bf16, 0);
            PF_DECL
            mbar_wait(smem_u32(q_full), 0);
            auto issue_s = [&](int pr) {  // S' of the tile pair pr (128 keys) into pair buffer pr & 1
                const int w = pr & 1;
                mbar_wait(smem_u32(&k_full[0]), pr & 1);
                PF(0);
                mbar_wait(smem_u32(&s_empty[w]), ((pr >> 1) & 1) ^ 1);
                PF(1);
                tcgen05_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint64_t dq = umma_smem_desc(smem_u32(Qb(kb)));
                    const uint64_t dk = umma_smem_desc(smem_u32(Kb(0, kb)));  // 128 key rows: tiles u = 0, 1 back to back
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tcgen05_mma_f16(tmem_base + w * 2 * AP_KEYS, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc_s,
                                        (kb > 0 || kk > 0));
                }
                tcgen05_commit(smem_u32(&s_full[w]));
                tcgen05_commit(smem_u32(&k_empty[0]));
                PF(2);
            };
            for (int pr = 0; 2 * pr < T; ++pr) issue_s(pr);
            PF_FLUSH(8);
        }
    };

    if constexpr (!kWideMovers) {
        // warp 0 TMA, 1 a_n v_n issuer, 2-3 movers, 4-11 softmax, 12-17 movers, 18 S' issuer, 19 p Vd issuer
        if (warp == 0) role_producer();
        else if (warp == 1) role_an_issuer();
        else if (warp < 4) state_movers();
        else if (warp < 12) role_softmax();
        else if (warp == kPIssueWarp) role_p_issuer();
        else if (warp == kSIssueWarp) role_s_issuer();
        else state_movers();
    } else {
        // by warpgroup, so that each setmaxnreg is executed by the four warps of a group together and dominates the code it
        // budgets: 0 = TMA / a_n issuer, 1-2 = softmax, 3-6 = movers, 7 = S' / p issuers
        const int wg = warp >> 2;
        if (wg == 0) {
            setmaxnreg_dec<40>();
            if (warp == 0) role_producer();
            else if (warp == 1) role_an_issuer();
        } else if (wg <= 2) {
            setmaxnreg_inc<104>();
            role_softmax();
        } else if (wg <= 6) {
            setmaxnreg_dec<48>();
            state_movers();
        } else {
            setmaxnreg_dec<40>();
            if (warp == kSIssueWarp) role_s_issuer();
            else if (warp == kPIssueWarp) role_p_issuer();
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// One-hot key coordinates for the augmented K operand: row j = [onehot(ky_j) (64) | onehot(kx_j) (64)].
template <bool BF16>
__global__ void __launch_bounds__(256) onehot_kernel(const long long* idx, uint16_t* oh, int rows, int gw) {
    et_pdl_prologue();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk (8 columns) per thread
    if (g >= rows * 16) return;
    const int j = g >> 4, chunk = g & 15;
    const int tok = idx != nullptr ? (int)idx[j] : j;
    const int ky = tok / gw, kx = tok - ky * gw;
    const int hot = chunk < 8 ? ky : 64 + kx;
    const uint32_t one = float_to_elem<BF16>(1.f);
    uint32_t w[4] = {0, 0, 0, 0};
    const int base = chunk * 8;
    if (hot >= base && hot < base + 8) w[(hot - base) >> 1] = one << (((hot - base) & 1) * 16);
    *reinterpret_cast<uint4*>(oh + (size_t)j * 128 + base) = make_uint4(w[0], w[1], w[2], w[3]);
}


template <bool BF16, bool RAGGED>
int launch_tc(const void* qkv, const void* sel, void* onehot, void* onehot_all, const TcArgs& a, int mode, cudaStream_t s) {
    int rc;
    const bool gen = a.has_bias && a.gw != 64;  // general grid: the statistics pass takes the bias from one-hot MMAs too
    if ((rc = et_raise_smem(tc_stats_kernel<BF16, false, RAGGED>, StSmem<false>::TOTAL))) return rc;
    if ((rc = et_raise_smem(tc_stats_kernel<BF16, true, RAGGED>, StSmem<true>::TOTAL))) return rc;
    if ((rc = et_raise_smem(tc_apply_kernel<BF16, ET_ATTN_DENSE, RAGGED>, AP_SMEM))) return rc;
    if ((rc = et_raise_smem(tc_apply_kernel<BF16, ET_ATTN_FIRST, RAGGED>, AP_SMEM))) return rc;
    if ((rc = et_raise_smem(tc_apply_kernel<BF16, ET_ATTN_DELTA, RAGGED>, AP_SMEM))) return rc;
    CUtensorMap tm128, tm64, tmsel, tmbh, tmbw, tmoh, tmoh_all;
    if ((rc = make_tmap_2d(&tm128, qkv, (long long)a.B * a.N, 3LL * a.D, 128, a.is_bf16))) return rc;
    if ((rc = make_tmap_2d(&tm64, qkv, (long long)a.B * a.N, 3LL * a.D, 64, a.is_bf16))) return rc;
    tmbh = tmbw = tmoh = tmoh_all = tm128;  // placeholders when there is no rel-pos bias (never dereferenced)
    const int oh_rows = mode == ET_ATTN_DELTA ? a.B * a.k : a.N;
    if (a.has_bias) {
        const long long brows = (long long)a.B * a.H * a.N;
        if ((rc = make_tmap_2d(&tmbh, a.bias_h, brows, 64, 128, a.is_bf16))) return rc;
        if ((rc = make_tmap_2d(&tmbw, a.bias_w, brows, 64, 128, a.is_bf16))) return rc;
        if ((rc = make_tmap_2d(&tmoh, onehot, oh_rows, 128, 64, a.is_bf16))) return rc;
        et_launch(onehot_kernel<BF16>, dim3((oh_rows * 16 + 255) / 256), dim3(256), 0, s, mode == ET_ATTN_DELTA ? a.idx : nullptr,
                                                                        static_cast<uint16_t*>(onehot), oh_rows, a.gw);
        ET_COUNT_LAUNCH(1);
        if (gen) {  // one-hot coordinates of ALL keys for the statistics pass (the apply pass of DELTA mode has the selected ones)
            void* all = mode == ET_ATTN_DELTA ? onehot_all : onehot;
            if (mode == ET_ATTN_DELTA) {
                et_launch(onehot_kernel<BF16>, dim3((a.N * 16 + 255) / 256), dim3(256), 0, s, (const long long*)nullptr,
                          static_cast<uint16_t*>(all), a.N, a.gw);
                ET_COUNT_LAUNCH(1);
            }
            if ((rc = make_tmap_2d(&tmoh_all, all, a.N, 128, 128, a.is_bf16))) return rc;
        }
    }
    const dim3 grid((a.N + QROWS - 1) / QROWS, a.H, a.B);
    if (gen) et_launch(tc_stats_kernel<BF16, true, RAGGED>, dim3(grid), dim3(kStThreads), StSmem<true>::TOTAL, s, tm128, tmbh, tmbw, tmoh_all, a);
    else et_launch(tc_stats_kernel<BF16, false, RAGGED>, dim3(grid), dim3(kStThreads), StSmem<false>::TOTAL, s, tm128, tmbh, tmbw, tmoh_all, a);
    ET_COUNT_LAUNCH(1);
    if (g_tc_time_apply) {
        if (g_ev0 == nullptr) {
            cudaEventCreate(&g_ev0);
            cudaEventCreate(&g_ev1);
        }
        cudaEventRecord(g_ev0, s);
    }
    if (mode == ET_ATTN_DELTA) {
        if ((rc = make_tmap_2d(&tmsel, sel, 3LL * a.sel_rows, (long long)a.D, 64, a.is_bf16))) return rc;
        const int cl = (g_tc_apply_cluster > 1 && grid.x % g_tc_apply_cluster == 0) ? g_tc_apply_cluster : 1;
        et_launch_cluster(tc_apply_kernel<BF16, ET_ATTN_DELTA, RAGGED>, dim3(grid), dim3(kApThreads), AP_SMEM, s, cl, tm128, tmsel, tmbh, tmbw, tmoh, a);
    } else if (mode == ET_ATTN_FIRST) {
        et_launch(tc_apply_kernel<BF16, ET_ATTN_FIRST, RAGGED>, dim3(grid), dim3(kApThreads), AP_SMEM, s, tm128, tm64, tmbh, tmbw, tmoh, a);
    } else {
        et_launch(tc_apply_kernel<BF16, ET_ATTN_DENSE, RAGGED>, dim3(grid), dim3(kApThreads), AP_SMEM, s, tm128, tm64, tmbh, tmbw, tmoh, a);
    }
    ET_COUNT_LAUNCH(1);
    if (g_tc_time_apply) cudaEventRecord(g_ev1, s);
    return ET_OK;
}

}  // namespace

extern "C" float et_debug_elapsed_ms(void) {
    float ms = -1.f;
    if (g_ev1 != nullptr && cudaEventSynchronize(g_ev1) == cudaSuccess) cudaEventElapsedTime(&ms, g_ev0, g_ev1);
    return ms;
}

// Entry used by et_global_attention (et_attn.cu) when the shape qualifies for the tensor-core path.
// `sel` = workspace rows [K_sel | v_n | v_n - dV], each (B * k, D); `onehot` = (max(B * k, N), 128) scratch, `onehot_all` =
// (N, 128) scratch (DELTA mode on grids that are not 64 wide); bias tables in the tc layout: (B, H, N, 64) each, holding
// 8 x bias, zero padded.  Any N >= 128 (ragged query blocks / key tiles are masked), grids up to 64 x 64.
int et_tc_global_attention(const void* qkv, const void* sel, void* onehot, void* onehot_all, const void* bias_h, const void* bias_w, int mode,
                           const long long* idx, int k, void* a_state, void* acc, void* out, float* stats, int B, int N,
                           int NP, int H, int gh, int gw, int is_bf16, cudaStream_t stream) {
    TcArgs a;
    a.bias_h = bias_h; a.bias_w = bias_w; a.stats = stats; a.idx = idx; a.a_state = a_state; a.acc = acc; a.out = out;
    a.B = B; a.N = N; a.NP = NP; a.H = H; a.D = H * 64; a.gh = gh; a.gw = gw; a.k = k; a.is_bf16 = is_bf16;
    a.sel_rows = B * k;
    a.has_bias = bias_h != nullptr;
    a.c1 = 0.125f * kLog2e;
    a.prof = g_tc_prof;
    const bool ragged = N % QROWS != 0;
    if (is_bf16) return ragged ? launch_tc<true, true>(qkv, sel, onehot, onehot_all, a, mode, stream)
                               : launch_tc<true, false>(qkv, sel, onehot, onehot_all, a, mode, stream);
    return ragged ? launch_tc<false, true>(qkv, sel, onehot, onehot_all, a, mode, stream)
                  : launch_tc<false, false>(qkv, sel, onehot, onehot_all, a, mode, stream);
}
