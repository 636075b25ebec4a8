// Dense windowed attention on tcgen05, second generation: one CTA per (window, head, HALF of the window's query rows),
// two CTAs resident per SM, the decomposed rel-pos bias computed inside the CTA (no bias tables in HBM, no helper launches).
// Replaces Block._forward_attention with window partition / bias-token padding / rel-pos / recombine
// (reference blocks.py:205-240,257-301,346-376; eventful_transformer/utils.py:139-171).
//
// Why (profiles/r1_ncu_full_summary_8streams.csv): the first-generation kernel (et_attn_win_tc.cu) ran one 139 KB CTA per
// SM -- load, S, softmax, P V and epilogue of a window strictly one after the other, 13 us per CTA, tensor pipe 11 % active --
// and needed two helper launches per call that round-tripped the bias rows through HBM.  Here:
//   * a CTA owns hq = ceil(wh / 2) window rows of queries (a rectangular 4-D TMA box, <= 128 tokens) and all Wn keys;
//     110 KB of shared memory and 256 TMEM columns, so two CTAs share an SM and one's softmax overlaps the other's loads / MMAs;
//   * windowed blocks size their rel-pos embedding to the window (blocks.py:86-91), so the gathered table is Toeplitz:
//     rel_y[qy][ky] = E_y[qy - ky + wh - 1].  One extra MMA  U = Q . [E_y ; E_x]^T  (128 x 64) gives every query row its
//     products with all 2w - 1 relative offsets; the row's 14 + 14 bias values are a static slice of U shifted by (qy, qx).
//     They go, times 8 and rounded to the operand dtype, into 32 extra K columns of the query operand; the key operand carries
//     one-hot(ky) | one-hot(kx) there, so  S' = [q | 8 bias_y | 8 bias_x] . [k | onehot | onehot]^T  has the bias folded in;
//   * softmax: two threads per query row (one per half of the key columns), max and exp2 / sum straight from TMEM, P written
//     as a K-major operand in 16-byte chunks (8 keys per store) over the dead Q / K operand memory;
//   * O = P V with V as MN-major B operand; the epilogue scales by 1 / l and writes only in-grid rows (crop + recombine).
// Warp roles (320 threads): warp 0 TMA, warp 1 TMEM alloc + MMA issue, warps 2-9 softmax (lane quarter = warp & 3,
// key-column half = (warp - 2) / 4).
#include "et_tcgen05.cuh"

using namespace et_tc;

// Profiling build (make prof): per-phase cycle buckets of the softmax role, summed over all CTAs, into row 7 of the buffer
// registered with et_debug_set(4, ...) (profiles/window_phases.py).
extern unsigned long long* g_tc_prof;
#ifdef ET_TC_PROFILE
#define WPF_DECL long long pf_[16]; for (int i_ = 0; i_ < 16; ++i_) pf_[i_] = 0; long long pf_t_ = clock64();
#define WPF(i) do { const long long n_ = clock64(); pf_[i] += n_ - pf_t_; pf_t_ = n_; } while (0)
#define WPF_FLUSH() do { if (a.prof) for (int i_ = 0; i_ < 16; ++i_) atomicAdd(a.prof + 7 * 16 + i_, (unsigned long long)pf_[i_]); } while (0)
#else
#define WPF_DECL
#define WPF(i)
#define WPF_FLUSH()
#endif

namespace {

constexpr int kThreads = 320;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int MAXK = 208;                 // keys per window, padded to a multiple of 16
constexpr int OFF_Q = 0;                  // 128 query rows x 64 channels            16 KB
constexpr int OFF_QB = 16384;             // their 32 bias columns (128-byte rows)    16 KB
constexpr int OFF_K = 32768;              // 208 keys x 64 channels                   26 KB
constexpr int OFF_KOH = OFF_K + MAXK * 128;   // one-hot key coordinates              26 KB
constexpr int OFF_V = OFF_KOH + MAXK * 128;   // values (the E table lives here first) 26 KB
constexpr int OFF_MISC = OFF_V + MAXK * 128;  // barriers, TMEM slot, row statistics exchange
constexpr int SMEM_BYTES = OFF_MISC + 1280 + 1024;  // barriers + exchange buffer, + alignment slack
constexpr int P_ATOM = 16384;             // one 64-key block of P: 128 rows x 128 bytes
static_assert(4 * P_ATOM <= OFF_V, "P must fit inside the dead Q / bias / K / one-hot region");
static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs per SM");

struct Win2Args {
    const void* pad_token;  // qkv bias (3D elements)
    const void* rel_y;      // (wh, wh, 64) Toeplitz gather of the y embedding, or null
    const void* rel_x;      // (ww, ww, 64)
    void* out;
    int B, N, gh, gw, wh, ww, nwx, nwy, H, D, Wn, NK, hq, has_bias;
    unsigned long long* prof;
    int inv_ww;  // (1 << 20) / ww + 1: x / ww == (x * inv_ww) >> 20 for x < 4096, ww <= 16 (no integer division in the kernel)
    float c1;
};

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    if constexpr (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool BF16>
__device__ __forceinline__ uint32_t one_elem() {
    return BF16 ? 0x3f80u : 0x3c00u;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_load_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// v[i] <- v[i + s] for the low entries (a static shift applied when `on`): building block of the barrel shifter
template <int S>
__device__ __forceinline__ void shift_down(float (&v)[32], bool on) {
#pragma unroll
    for (int i = 0; i + S < 32; ++i) v[i] = on ? v[i + S] : v[i];
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 2)
tc_window2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const Win2Args a) {
    et_pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Qq = smem + OFF_Q;
    uint8_t* Qb = smem + OFF_QB;
    uint8_t* Kk = smem + OFF_K;
    uint8_t* Koh = smem + OFF_KOH;
    uint8_t* Vs = smem + OFF_V;   // E table until U is computed, then the values
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
    uint64_t* qk_full = bars;         // TMA: Q and K landed
    uint64_t* v_full = bars + 1;      // TMA: V landed
    uint64_t* ops_ready = bars + 2;   // softmax warps: E table / one-hot built, pad rows patched (8 arrivals)
    uint64_t* u_full = bars + 3;      // MMA: U in TMEM (E consumed)
    uint64_t* qb_ready = bars + 4;    // softmax warps: bias columns written (8 arrivals)
    uint64_t* s_full = bars + 5;      // MMA: S' in TMEM
    uint64_t* p_ready = bars + 6;     // softmax warps: P written, V patched (8 arrivals)
    uint64_t* o_full = bars + 7;      // MMA: O in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* xchg = reinterpret_cast<float*>(smem + OFF_MISC + 128);  // [2 column halves][128 rows] max, then sum (2 x 1 KB)

    WPF_DECL
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwin = a.nwx * a.nwy;
    auto div_ww = [&](int x) { return (x * a.inv_ww) >> 20; };
    const int half = blockIdx.x & 1, bw = blockIdx.x >> 1, h = blockIdx.y;
    const int b = bw / nwin, win = bw - b * nwin;
    const int wy = win / a.nwx, wx = win - wy * a.nwx;
    const int y0 = half * a.hq;                                  // first window row of this CTA's queries
    const int nq = min(a.hq, a.wh - y0) * a.ww;                  // valid query rows (<= 128)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_kv) : "memory");
        mbar_init(smem_u32(qk_full), 1);
        mbar_init(smem_u32(v_full), 1);
        mbar_init(smem_u32(ops_ready), 8);
        mbar_init(smem_u32(u_full), 1);
        mbar_init(smem_u32(qb_ready), 8);
        mbar_init(smem_u32(s_full), 1);
        mbar_init(smem_u32(p_ready), 8);
        mbar_init(smem_u32(o_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    WPF(0);  // prologue: barrier init, TMEM allocation, CTA sync

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t fb = smem_u32(qk_full);
            mbar_expect_tx(fb, (a.hq * a.ww + a.Wn) * 128);  // a box always transfers its full size (out-of-grid rows as zeros)
            tma_load_4d(smem_u32(Qq), &tm_q, fb, h * 64, wx * a.ww, wy * a.wh + y0, b);
            tma_load_4d(smem_u32(Kk), &tm_kv, fb, a.D + h * 64, wx * a.ww, wy * a.wh, b);
            if (a.has_bias) mbar_wait(smem_u32(u_full), 0);  // V overlays the E table
            const uint32_t vb = smem_u32(v_full);
            mbar_expect_tx(vb, a.Wn * 128);
            tma_load_4d(smem_u32(Vs), &tm_kv, vb, 2 * a.D + h * 64, wx * a.ww, wy * a.wh, b);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_u = umma_idesc_ex(128, 64, BF16, 0);
            const uint32_t idesc_s = umma_idesc_ex(128, a.NK, BF16, 0);
            const uint32_t idesc_o = umma_idesc_ex(128, 64, BF16, 1);  // A = P K-major, B = V MN-major
            const uint64_t dq = umma_smem_desc(smem_u32(Qq));
            mbar_wait(smem_u32(ops_ready), 0);
            tcgen05_fence_after();
            if (a.has_bias) {
                const uint64_t de = umma_smem_desc(smem_u32(Vs));
                for (int kk = 0; kk < 4; ++kk) tcgen05_mma_f16(tmem_base, dq + (uint64_t)(2 * kk), de + (uint64_t)(2 * kk), idesc_u, kk > 0);
                tcgen05_commit(smem_u32(u_full));
                mbar_wait(smem_u32(qb_ready), 0);
                tcgen05_fence_after();
            }
            const uint64_t dk = umma_smem_desc(smem_u32(Kk));
            for (int kk = 0; kk < 4; ++kk) tcgen05_mma_f16(tmem_base, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc_s, kk > 0);
            if (a.has_bias) {
                const uint64_t dqb = umma_smem_desc(smem_u32(Qb)), dko = umma_smem_desc(smem_u32(Koh));
                for (int kk = 0; kk < 2; ++kk) tcgen05_mma_f16(tmem_base, dqb + (uint64_t)(2 * kk), dko + (uint64_t)(2 * kk), idesc_s, 1);
            }
            tcgen05_commit(smem_u32(s_full));
            mbar_wait(smem_u32(p_ready), 0);
            tcgen05_fence_after();
            const uint64_t dv = umma_smem_desc_mn(smem_u32(Vs));
            for (int kk = 0; kk < a.NK / 16; ++kk) {
                const uint64_t dp = umma_smem_desc(smem_u32(smem + (kk >> 2) * P_ATOM)) + (uint64_t)(2 * (kk & 3));
                tcgen05_mma_f16(tmem_base, dp, dv + (uint64_t)(128 * kk), idesc_o, kk > 0);
            }
            tcgen05_commit(smem_u32(o_full));
        }
    } else {
        const int st = threadIdx.x - 64;       // 0..255
        const int quarter = warp & 3;          // TMEM lane quarter of this warp
        const int ch = (warp - 2) >> 2;        // key-column half (and, for the bias step, y / x part)
        const int row = quarter * 32 + lane;   // query row inside the CTA tile
        const int ry_ = div_ww(row);
        const int ly = y0 + ry_, lx = row - ry_ * a.ww;  // window-local coordinates of the query
        const uint16_t* pad = static_cast<const uint16_t*>(a.pad_token);
        // padding tokens equal the qkv bias: the patch loops below touch chunk (st & 7) of a row only, so every thread needs one
        // 16-byte piece of the head's q / k / v bias -- fetched here, once, instead of inside the loops
        const bool edge_win = (wx + 1) * a.ww > a.gw || (wy + 1) * a.wh > a.gh;
        uint4 pad_q = make_uint4(0, 0, 0, 0), pad_k = pad_q, pad_v = pad_q;
        if (edge_win) {
            pad_q = *reinterpret_cast<const uint4*>(pad + h * 64 + (st & 7) * 8);
            pad_k = *reinterpret_cast<const uint4*>(pad + a.D + h * 64 + (st & 7) * 8);
            pad_v = *reinterpret_cast<const uint4*>(pad + 2 * a.D + h * 64 + (st & 7) * 8);
        }

        // ---- operands that no TMA writes: the E table (in the V region), the one-hot key block, zero pad rows of V later
        if (a.has_bias) {
            // E rows j = 0..2wh-2: y embedding of relative offset 2wh-2-j (REVERSED, so that a query at qy finds its bias
            // for key row ky at column (wh-1-qy) + ky); rows 32..32+2ww-2: the same for x; other rows zero
            const uint16_t* ry = static_cast<const uint16_t*>(a.rel_y);
            const uint16_t* rx = static_cast<const uint16_t*>(a.rel_x);
            for (int c = st; c < 64 * 8; c += 256) {
                const int i = c >> 3, chunk = c & 7;
                const int part = i >> 5, w = part ? a.ww : a.wh, off = 2 * w - 2 - (i & 31);
                uint8_t* dst = Vs + i * 128 + ((chunk ^ (i & 7)) << 4);
                if (off >= 0) {
                    // offset = q - k + w - 1: realised by (q, k) = (off - (w - 1), 0) or (0, w - 1 - off)
                    const int qc = off >= w - 1 ? off - (w - 1) : 0, kc = off >= w - 1 ? 0 : w - 1 - off;
                    cp_async_16(smem_u32(dst), (part ? rx : ry) + ((size_t)qc * w + kc) * 64 + chunk * 8);
                } else {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                }
            }
            // one-hot block: key j -> 1 at column (j / ww) and at column 16 + (j % ww); columns 0..31 = chunks 0..3.
            // One thread per key: chunk c holds columns 8c .. 8c + 7, element e of a chunk sits in word e / 2, half e % 2.
            if (st < a.NK) {
                const int j = st;
                const int ky = div_ww(j), kx = j - ky * a.ww;
                const uint32_t one = one_elem<BF16>();
#pragma unroll
                for (int chunk = 0; chunk < 4; ++chunk) {
                    const int d = (chunk < 2 ? ky : 16 + kx) - chunk * 8;
                    const bool on = j < a.Wn && d >= 0 && d < 8;
                    const uint32_t wv = on ? one << ((d & 1) * 16) : 0u;
                    const int wi = on ? d >> 1 : -1;
                    *reinterpret_cast<uint4*>(Koh + j * 128 + ((chunk ^ (j & 7)) << 4)) =
                        make_uint4(wi == 0 ? wv : 0u, wi == 1 ? wv : 0u, wi == 2 ? wv : 0u, wi == 3 ? wv : 0u);
                }
            }
        }
        // rows [Wn, NK) of K must be finite?  No: their S' columns are never read.  (V pad rows are zeroed below.)
        WPF(1);  // E table + one-hot block
        mbar_wait(smem_u32(qk_full), 0);
        WPF(2);  // wait: Q and K landed
        // ---- padding tokens (outside the grid) equal the qkv bias: patch the zero-filled rows (swizzled chunks)
        const bool edge = (wx + 1) * a.ww > a.gw || (wy + 1) * a.wh > a.gh;
        if (edge) {
            for (int c = st; c < a.Wn * 8; c += 256) {  // keys
                const int r = c >> 3, chunk = c & 7;
                const int ky = div_ww(r), kx = r - ky * a.ww;
                if (wy * a.wh + ky >= a.gh || wx * a.ww + kx >= a.gw)
                    *reinterpret_cast<uint4*>(Kk + r * 128 + ((chunk ^ (r & 7)) << 4)) = pad_k;
            }
            for (int c = st; c < nq * 8; c += 256) {    // queries of this half
                const int r = c >> 3, chunk = c & 7;
                const int qr = div_ww(r);
                const int qy = y0 + qr, qx = r - qr * a.ww;
                if (wy * a.wh + qy >= a.gh || wx * a.ww + qx >= a.gw)
                    *reinterpret_cast<uint4*>(Qq + r * 128 + ((chunk ^ (r & 7)) << 4)) = pad_q;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(ops_ready));
        WPF(3);  // pad patch + cp.async wait + fence + arrive

        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // ---- rel-pos bias of this row: column half 0 takes the y part, half 1 the x part
        if (a.has_bias) {
            mbar_wait(smem_u32(u_full), 0);
            WPF(4);  // wait: U = Q E^T
            tcgen05_fence_after();
            uint32_t raw[32];
            tmem_load_32x32(trow + (uint32_t)(ch * 32), raw);
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
            // bias[kc] = q . E[qc - kc + w - 1] = U[(w - 1 - qc) + kc] with the reversed table: a barrel shift by
            // w - 1 - qc (static register indices, no divergence), then the first w entries
            const int w = ch ? a.ww : a.wh;
            const int sh = max(0, w - 1 - (ch ? lx : ly));
            shift_down<1>(v, sh & 1);
            shift_down<2>(v, sh & 2);
            shift_down<4>(v, sh & 4);
            shift_down<8>(v, sh & 8);
            float bias[16];
#pragma unroll
            for (int kc = 0; kc < 16; ++kc) bias[kc] = kc < w ? 8.f * v[kc] : 0.f;
            uint32_t w8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w8[i] = pack2<BF16>(bias[2 * i], bias[2 * i + 1]);
            uint8_t* qrow = Qb + row * 128;
            *reinterpret_cast<uint4*>(qrow + (((2 * ch) ^ (row & 7)) << 4)) = make_uint4(w8[0], w8[1], w8[2], w8[3]);
            *reinterpret_cast<uint4*>(qrow + (((2 * ch + 1) ^ (row & 7)) << 4)) = make_uint4(w8[4], w8[5], w8[6], w8[7]);
            tcgen05_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(qb_ready));
            WPF(5);  // bias columns
        }

        // ---- V: zero the pad rows [Wn, NK) (P is zero there, but 0 x garbage must stay 0) and patch out-of-grid tokens.  Done
        // here, while the S' MMAs run: V has been in flight since U was read, and after the softmax only the fence remains
        mbar_wait(smem_u32(v_full), 0);
        WPF(9);  // wait: V landed
        for (int c = st; c < (a.NK - a.Wn) * 8; c += 256)
            *reinterpret_cast<uint4*>(Vs + (a.Wn + c / 8) * 128 + (c & 7) * 16) = make_uint4(0, 0, 0, 0);
        if (edge) {
            for (int c = st; c < a.Wn * 8; c += 256) {
                const int r = c >> 3, chunk = c & 7;
                const int ky = div_ww(r), kx = r - ky * a.ww;
                if (wy * a.wh + ky >= a.gh || wx * a.ww + kx >= a.gw)
                    *reinterpret_cast<uint4*>(Vs + r * 128 + ((chunk ^ (r & 7)) << 4)) = pad_v;
            }
        }
        WPF(10);  // V patch
        // ---- softmax over this thread's half of the keys: columns [c_lo, c_hi)
        const int half_cols = a.NK / 2 / 8 * 8;            // multiple of 8 so that P chunks stay whole (NK = 208 -> 104)
        const int c_lo = ch ? half_cols : 0, c_hi = ch ? a.NK : half_cols;
        mbar_wait(smem_u32(s_full), 0);
        WPF(6);  // wait: S'
        tcgen05_fence_after();
        float mx = -1e30f;
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            if (c0 + 32 <= min(c_hi, a.Wn)) {  // whole block of real keys: no per-element masks (a warp-uniform branch)
                uint32_t t[32];
                tmem_load_32x32(trow + (uint32_t)c0, t);
                float m4[4] = {mx, -1e30f, -1e30f, -1e30f};
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    m4[0] = fmaxf(m4[0], fmaxf(__uint_as_float(t[i]), __uint_as_float(t[i + 1])));
                    m4[1] = fmaxf(m4[1], fmaxf(__uint_as_float(t[i + 2]), __uint_as_float(t[i + 3])));
                    m4[2] = fmaxf(m4[2], fmaxf(__uint_as_float(t[i + 4]), __uint_as_float(t[i + 5])));
                    m4[3] = fmaxf(m4[3], fmaxf(__uint_as_float(t[i + 6]), __uint_as_float(t[i + 7])));
                }
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else {
                for (int c1 = c0; c1 < min(c0 + 32, c_hi); c1 += 8) {
                    uint32_t t[8];
                    tmem_load_32x8(trow + (uint32_t)c1, t);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (c1 + i < a.Wn) mx = fmaxf(mx, __uint_as_float(t[i]));
                }
            }
        }
        xchg[ch * 128 + row] = mx;
        pair_sync(1 + quarter);
        mx = fmaxf(mx, xchg[(ch ^ 1) * 128 + row]);
        WPF(7);  // row max
        const float m2 = mx * a.c1;
        const f32x2 c1c1 = f2_pack(a.c1, a.c1), nm2 = f2_pack(-m2, -m2);
        f32x2 sum2 = f2_pack(0.f, 0.f);
        float sum = 0.f;
        auto p_chunk = [&](int key0) {  // the 16-byte chunk of keys key0 .. key0 + 7 in this row of the K-major P operand
            const int kc = key0 >> 3;
            return reinterpret_cast<uint4*>(smem + (kc >> 3) * P_ATOM + row * 128 + (((kc & 7) ^ (row & 7)) << 4));
        };
        auto emit8_full = [&](int key0, const uint32_t* t) {  // packed fp32 math: two keys per FFMA2 / FADD2 issue slot
            uint32_t w4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float x0, x1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(t[2 * i]), __uint_as_float(t[2 * i + 1])), c1c1, nm2), x0, x1);
                const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                sum2 = f2_add(sum2, f2_pack(p0, p1));
                w4[i] = pack2<BF16>(p0, p1);
            }
            *p_chunk(key0) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        };
        auto emit8 = [&](int key0, const uint32_t* t) {  // masked form for the block that holds the padding keys
            float p[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                p[i] = (key0 + i < a.Wn) ? ex2_approx(fmaf(__uint_as_float(t[i]), a.c1, -m2)) : 0.f;
                sum += p[i];
            }
            *p_chunk(key0) = make_uint4(pack2<BF16>(p[0], p[1]), pack2<BF16>(p[2], p[3]), pack2<BF16>(p[4], p[5]), pack2<BF16>(p[6], p[7]));
        };
        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            if (c0 + 32 <= min(c_hi, a.Wn)) {
                uint32_t t[32];
                tmem_load_32x32(trow + (uint32_t)c0, t);
#pragma unroll
                for (int i = 0; i < 32; i += 8) emit8_full(c0 + i, t + i);
            } else {
                for (int c1 = c0; c1 < min(c0 + 32, c_hi); c1 += 8) {
                    uint32_t t[8];
                    tmem_load_32x8(trow + (uint32_t)c1, t);
                    if (c1 + 8 <= a.Wn) emit8_full(c1, t);
                    else emit8(c1, t);
                }
            }
        }
        {
            float s0, s1;
            f2_unpack(sum2, s0, s1);
            sum += s0 + s1;
        }
        WPF(8);  // exp2 + P stores
        pair_sync(1 + quarter);  // both maxima have been read
        xchg[ch * 128 + row] = sum;
        tcgen05_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(p_ready));
        pair_sync(1 + quarter);
        sum += xchg[(ch ^ 1) * 128 + row];

        // ---- epilogue: O / l -> out[b, token, h * 64 + 32 ch ...] for in-grid rows (window recombine + crop)
        mbar_wait(smem_u32(o_full), 0);
        WPF(11);  // wait: O = P V
        tcgen05_fence_after();
        uint32_t o[32];
        tmem_load_32x32(trow + (uint32_t)(ch * 32), o);
        const int gy = wy * a.wh + ly, gx = wx * a.ww + lx;
        if (row < nq && gy < a.gh && gx < a.gw) {
            const float inv = 1.f / sum;
            uint16_t* dst = static_cast<uint16_t*>(a.out) + ((size_t)b * a.N + (size_t)gy * a.gw + gx) * a.D + h * 64 + ch * 32;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t w4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    w4[i] = pack2<BF16>(__uint_as_float(o[c * 8 + 2 * i]) * inv, __uint_as_float(o[c * 8 + 2 * i + 1]) * inv);
                *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
        }
    }
    if (warp == 2) WPF(12);  // epilogue stores
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
    if (warp == 2) {
        WPF(13);  // final sync + TMEM release
        if (lane == 0) WPF_FLUSH();
    }
}

int make_tmap_grid4d(CUtensorMap* map, const void* base, int B, int gh, int gw, int C, int box_h, int box_w, int is_bf16) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return et_fail(ET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)gw * C * 2, (cuuint64_t)gh * gw * C * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return et_fail(ET_ERR_CUDA, "cuTensorMapEncodeTiled(4-D window map) failed with CUresult %d", (int)r);
    return ET_OK;
}

}  // namespace

// Whether the second-generation kernel applies: 64-wide heads, windows of at most 16 x 16 tokens and 208 keys, and half a
// window's queries fitting one 128-row tile.
bool et_tc_window2_applies(int wh, int ww, int dh) {
    const int hq = (wh + 1) / 2;
    return dh == 64 && wh <= 16 && ww <= 16 && wh * ww <= MAXK && hq * ww <= 128;
}

// rel_y / rel_x: the (wh, wh, 64) / (ww, ww, 64) gathered tables of a window-sized embedding (Toeplitz), or null.
int et_tc_window2_attention(const void* qkv, const void* pad_token, const void* rel_y, const void* rel_x, void* out, int B, int N,
                            int gh, int gw, int wh, int ww, int H, int is_bf16, cudaStream_t s) {
    Win2Args a;
    a.pad_token = pad_token; a.rel_y = rel_y; a.rel_x = rel_x; a.out = out; a.B = B; a.N = N; a.gh = gh; a.gw = gw; a.wh = wh; a.ww = ww;
    a.nwy = (gh + wh - 1) / wh; a.nwx = (gw + ww - 1) / ww; a.H = H; a.D = H * 64; a.Wn = wh * ww;
    a.NK = (a.Wn + 15) / 16 * 16; a.hq = (wh + 1) / 2; a.has_bias = rel_y != nullptr ? 1 : 0; a.c1 = 0.125f * kLog2e;
    a.inv_ww = (1 << 20) / ww + 1;
    a.prof = g_tc_prof;
    const int nwin = a.nwx * a.nwy;
    int rc;
    if ((rc = et_raise_smem(tc_window2_kernel<true>, SMEM_BYTES))) return rc;
    if ((rc = et_raise_smem(tc_window2_kernel<false>, SMEM_BYTES))) return rc;
    CUtensorMap tq, tkv;
    if ((rc = make_tmap_grid4d(&tq, qkv, B, gh, gw, 3 * a.D, a.hq, ww, is_bf16))) return rc;
    if ((rc = make_tmap_grid4d(&tkv, qkv, B, gh, gw, 3 * a.D, wh, ww, is_bf16))) return rc;
    const dim3 grid(B * nwin * 2, H);
    if (is_bf16) et_launch(tc_window2_kernel<true>, dim3(grid), dim3(kThreads), SMEM_BYTES, s, tq, tkv, a);
    else et_launch(tc_window2_kernel<false>, dim3(grid), dim3(kThreads), SMEM_BYTES, s, tq, tkv, a);
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}
