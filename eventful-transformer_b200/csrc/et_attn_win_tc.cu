// Dense windowed attention of EventfulTokenwiseBlock / Block on tcgen05 (sm_100a): one CTA per (window, head).
// Replaces Block._forward_attention with window partition / bias-token padding / rel-pos / recombine
// (reference blocks.py:205-240,257-301,346-376; eventful_transformer/utils.py:139-171).
//
//   * a 4-D TMA box (64 channels x ww x wh x 1 image) fetches the window's q / k / v rows straight from the
//     (B, gh, gw, 3D) token grid in window-local order; out-of-grid rows arrive as zeros and are patched to the
//     qkv bias vector (the image of a zero token, blocks.py:275-287);
//   * S' = Q' K'^T with Q' = [q | 8 bias_h | 8 bias_w | 0], K' = [k | onehot(ly) | onehot(lx) | 0]: the decomposed
//     rel-pos bias comes out of the MMA; M = 2 x 128 query rows (w^2 <= 256), N = w^2 rounded up to 16 keys.  The
//     bias part is 32 columns wide (wh + ww <= 32): 64 bytes per query row, fetched with cp.async into the swizzled
//     operand rows, and two K = 16 MMAs;
//   * two softmax warpgroups (one per 128-row half, one row per thread) take max / exp2 / sum straight from TMEM and
//     write P as an MN-major A tile that overlays the (now dead) Q' / K' operand memory;
//   * O = P V with V as MN-major B operand from TMA; epilogue scales by 1 / l and writes only in-grid rows.
// Warp roles (320 threads): warp 0 TMA, warp 1 TMEM alloc + MMA issue, warps 2-5 rows 0-127, warps 6-9 rows 128-255.
#include "et_tcgen05.cuh"

using namespace et_tc;

namespace {

constexpr int kThreads = 320;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int W_QBLK = 256 * 128;    // 32 KB: 256 query rows x 64 channels
constexpr int W_KBLK = 208 * 128;    // 26 KB: up to 208 keys x 64 channels (multiple of 1024)
constexpr int W_OFF_QB = W_QBLK;     // Q' bias block
constexpr int W_OFF_K = 2 * W_QBLK;  // K block
constexpr int W_OFF_KOH = W_OFF_K + W_KBLK;
constexpr int W_OFF_V = W_OFF_KOH + W_KBLK;
constexpr int W_OFF_MISC = W_OFF_V + W_KBLK;
constexpr int W_SMEM = W_OFF_MISC + 256 + 1024;
constexpr int W_PBLK = W_KBLK;       // one 64-row block of a P tile: [key][64 rows] = keys x 128 B
static_assert(4 * W_PBLK <= W_OFF_V, "P tiles must fit inside the Q'/K' operand region they overlay");

struct WinArgs {
    const void* pad_token;  // qkv bias (3D)
    const void* bias;       // (B * nwin * H, 256, 32): [8 bias_h | 8 bias_w | 0] per window-local query row
    void* out;
    int B, N, gh, gw, wh, ww, nwx, nwy, H, D, Wn, NK, has_bias, is_bf16;
    float c1;
};

template <bool BF16>
__device__ __forceinline__ uint32_t pack2_elem(float lo, float hi) {  // one F2FP; `lo` in bits 0-15
    uint32_t r;
    if constexpr (BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool BF16>
__device__ __forceinline__ uint32_t to_elem(float v) {
    if constexpr (BF16) return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    else return (uint32_t)__half_as_ushort(__float2half_rn(v));
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :
        : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
tc_window_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_bias,
                 const __grid_constant__ CUtensorMap tm_oh, const WinArgs a) {
    et_pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Qq = smem;
    uint8_t* Qb = smem + W_OFF_QB;
    uint8_t* Kk = smem + W_OFF_K;
    uint8_t* Koh = smem + W_OFF_KOH;
    uint8_t* Vs = smem + W_OFF_V;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W_OFF_MISC);
    uint64_t* ld_full = bars;        // TMA landed
    uint64_t* ops_ready = bars + 1;  // operands patched (8 softmax warps)
    uint64_t* s_full = bars + 2;     // [2] S of the 128-row half is in TMEM
    uint64_t* p_ready = bars + 4;    // [2] P of the half written (4 warps)
    uint64_t* o_full = bars + 6;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwin = a.nwx * a.nwy;
    const int bw = blockIdx.x, h = blockIdx.y;
    const int b = bw / nwin, win = bw - b * nwin;
    const int wy = win / a.nwx, wx = win - wy * a.nwx;
    const int nkb = a.has_bias ? 2 : 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_qkv) : "memory");
        mbar_init(smem_u32(ld_full), 1);
        mbar_init(smem_u32(ops_ready), 8);
        for (int m = 0; m < 2; ++m) {
            mbar_init(smem_u32(&s_full[m]), 1);
            mbar_init(smem_u32(&p_ready[m]), 4);
            mbar_init(smem_u32(&o_full[m]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t fb = smem_u32(ld_full);
            const int rows = a.Wn * 128;
            mbar_expect_tx(fb, 3 * rows + (a.has_bias ? a.NK * 128 : 0));
            tma_load_4d(smem_u32(Qq), &tm_qkv, fb, h * 64, wx * a.ww, wy * a.wh, b);
            tma_load_4d(smem_u32(Kk), &tm_qkv, fb, a.D + h * 64, wx * a.ww, wy * a.wh, b);
            tma_load_4d(smem_u32(Vs), &tm_qkv, fb, 2 * a.D + h * 64, wx * a.ww, wy * a.wh, b);
            if (a.has_bias) {
                tma_load_2d(smem_u32(Koh), &tm_oh, fb, 0, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_ex(128, a.NK, a.is_bf16, 0);
            const uint32_t idesc_o = umma_idesc_ex(128, 64, a.is_bf16, 1) | (1u << 15);  // A and B MN-major
            mbar_wait(smem_u32(ops_ready), 0);
            tcgen05_fence_after();
            for (int m = 0; m < 2; ++m) {
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint64_t dq = umma_smem_desc(smem_u32((kb ? Qb : Qq) + m * 128 * 128));
                    const uint64_t dk = umma_smem_desc(smem_u32(kb ? Koh : Kk));
                    const int nkk = kb ? 2 : 4;  // the bias / one-hot block carries 32 meaningful columns
                    for (int kk = 0; kk < nkk; ++kk)
                        tcgen05_mma_f16(tmem_base + m * 256, dq + (uint64_t)(2 * kk), dk + (uint64_t)(2 * kk), idesc_s,
                                        (kb > 0 || kk > 0));
                }
                tcgen05_commit(smem_u32(&s_full[m]));
            }
            for (int m = 0; m < 2; ++m) {
                mbar_wait(smem_u32(&p_ready[m]), 0);
                tcgen05_fence_after();
                // P_m: MN-major A tile [key][row], two 64-row blocks W_PBLK bytes apart
                uint64_t dp = umma_smem_desc(smem_u32(smem + m * 2 * W_PBLK));
                dp &= ~((uint64_t)0x3fff << 16);
                dp |= (uint64_t)(W_PBLK >> 4) << 16;
                const uint64_t dv = umma_smem_desc_mn(smem_u32(Vs));
                for (int kk = 0; kk < a.NK / 16; ++kk)
                    tcgen05_mma_f16(tmem_base + m * 256, dp + (uint64_t)(128 * kk), dv + (uint64_t)(128 * kk), idesc_o, kk > 0);
                tcgen05_commit(smem_u32(&o_full[m]));
            }
        }
    } else {
        const int st = threadIdx.x - 64;          // 0..255
        const int m = (warp - 2) >> 2;            // 128-row half handled by this warpgroup
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;      // row inside the half
        const int wrow = m * 128 + row;           // window-local token
        const uint16_t* pad = static_cast<const uint16_t*>(a.pad_token);
        // ---- this thread's bias row: 64 bytes (8 bias_h | 8 bias_w | 0) -> logical chunks 0-3 of the swizzled operand row
        if (a.has_bias) {
            const uint16_t* brow = static_cast<const uint16_t*>(a.bias) + ((size_t)(bw * a.H + h) * 256 + st) * 32;
            uint8_t* qrow = Qb + st * 128;  // rows 128-255 continue in the second 16 KB block
#pragma unroll
            for (int c = 0; c < 4; ++c) cp_async_16(smem_u32(qrow + ((c ^ (st & 7)) << 4)), brow + c * 8);
        }
        // ---- rows that TMA does not write must be finite: zero rows [Wn, 256) of Q and [Wn, NK) of K / V
        for (int c = st; c < (256 - a.Wn) * 8; c += 256)
            *reinterpret_cast<uint4*>(Qq + (a.Wn + c / 8) * 128 + (c & 7) * 16) = make_uint4(0, 0, 0, 0);
        for (int c = st; c < (a.NK - a.Wn) * 8; c += 256) {
            *reinterpret_cast<uint4*>(Kk + (a.Wn + c / 8) * 128 + (c & 7) * 16) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(Vs + (a.Wn + c / 8) * 128 + (c & 7) * 16) = make_uint4(0, 0, 0, 0);
        }
        mbar_wait(smem_u32(ld_full), 0);
        // ---- padding tokens (outside the grid) equal the qkv bias: patch the zero-filled rows (swizzled chunks)
        const bool edge = (wx + 1) * a.ww > a.gw || (wy + 1) * a.wh > a.gh;
        if (edge) {
            for (int c = st; c < a.Wn * 24; c += 256) {
                const int r = c / 24, rem = c - r * 24, part = rem >> 3, ch = rem & 7;
                const int ly = r / a.ww, lx = r - ly * a.ww;
                if (wy * a.wh + ly >= a.gh || wx * a.ww + lx >= a.gw) {
                    uint8_t* tile = part == 0 ? Qq : (part == 1 ? Kk : Vs);
                    *reinterpret_cast<uint4*>(tile + r * 128 + ((ch ^ (r & 7)) << 4)) =
                        *reinterpret_cast<const uint4*>(pad + part * a.D + h * 64 + ch * 8);
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(ops_ready));

        // ---- softmax of this thread's row over the Wn valid keys, single pass pair (max, then exp / sum / P)
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(m * 256);
        mbar_wait(smem_u32(&s_full[m]), 0);
        tcgen05_fence_after();
        float mx = -1e30f;
        for (int c0 = 0; c0 < a.NK; c0 += 32) {
            if (c0 + 32 <= a.NK) {
                uint32_t v[32];
                tmem_load_32x32(taddr + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < a.Wn) mx = fmaxf(mx, __uint_as_float(v[i]));
            } else {
                uint32_t v[16];
                tmem_load_32x16(taddr + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < a.Wn) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
        }
        const float m2 = mx * a.c1;
        if (m == 0) mbar_wait(smem_u32(&s_full[1]), 0);  // P overlays Q' / K': every S MMA must have retired
        uint8_t* pt = smem + m * 2 * W_PBLK + (row >> 6) * W_PBLK + (row & 7) * 2;
        const int rchunk = (row >> 3) & 7;
        float sum = 0.f;
        auto emit2 = [&](int key, float x0, float x1) {  // keys key, key + 1 (key even)
            float p0 = 0.f, p1 = 0.f;
            if (key < a.Wn) {
                p0 = ex2_approx(fmaf(x0, a.c1, -m2));
                sum += p0;
            }
            if (key + 1 < a.Wn) {
                p1 = ex2_approx(fmaf(x1, a.c1, -m2));
                sum += p1;
            }
            const uint32_t pk = pack2_elem<BF16>(p0, p1);
            *reinterpret_cast<uint16_t*>(pt + key * 128 + ((rchunk ^ (key & 7)) << 4)) = (uint16_t)(pk & 0xffffu);
            *reinterpret_cast<uint16_t*>(pt + (key + 1) * 128 + ((rchunk ^ ((key + 1) & 7)) << 4)) = (uint16_t)(pk >> 16);
        };
        for (int c0 = 0; c0 < a.NK; c0 += 32) {
            if (c0 + 32 <= a.NK) {
                uint32_t v[32];
                tmem_load_32x32(taddr + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; i += 2) emit2(c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            } else {
                uint32_t v[16];
                tmem_load_32x16(taddr + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; i += 2) emit2(c0 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            }
        }
        tcgen05_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&p_ready[m]));

        // ---- epilogue: O / l -> out[b, token, h * 64 ...] for in-grid rows (window recombine + crop)
        mbar_wait(smem_u32(&o_full[m]), 0);
        tcgen05_fence_after();
        uint32_t lo[32], hi[32];
        tmem_load_32x32(taddr, lo);
        tmem_load_32x32(taddr + 32u, hi);
        const int ly = wrow / a.ww, lx = wrow - ly * a.ww;
        const int gy = wy * a.wh + ly, gx = wx * a.ww + lx;
        if (wrow < a.Wn && gy < a.gh && gx < a.gw) {
            const float inv = 1.f / sum;
            uint16_t* dst = static_cast<uint16_t*>(a.out) + ((size_t)b * a.N + (size_t)gy * a.gw + gx) * a.D + h * 64;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t* src = c < 4 ? lo : hi;
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = (c & 3) * 8 + 2 * i;
                    w[i] = pack2_elem<BF16>(__uint_as_float(src[e]) * inv, __uint_as_float(src[e + 1]) * inv);
                }
                *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// K' one-hot block of a (wh x ww) window: row j -> column (j / ww) and column wh + (j % ww); 256 rows x 64 columns.
template <bool BF16>
__global__ void __launch_bounds__(256) window_onehot_kernel(uint16_t* oh, int wh, int ww) {
    et_pdl_prologue();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk per thread: 256 rows x 8 chunks
    if (g >= 256 * 8) return;
    const int j = g >> 3, chunk = g & 7;
    uint32_t w[4] = {0, 0, 0, 0};
    if (j < wh * ww) {
        const uint32_t one = to_elem<BF16>(1.f);
        const int hots[2] = {j / ww, wh + j % ww};
        for (int t = 0; t < 2; ++t) {
            const int d = hots[t] - chunk * 8;
            if (d >= 0 && d < 8) w[d >> 1] |= one << ((d & 1) * 16);
        }
    }
    *reinterpret_cast<uint4*>(oh + (size_t)j * 64 + chunk * 8) = make_uint4(w[0], w[1], w[2], w[3]);
}

int make_tmap_grid4d(CUtensorMap* map, const void* base, int B, int gh, int gw, int C, int wh, int ww, int is_bf16) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return et_fail(ET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)gw * C * 2, (cuuint64_t)gh * gw * C * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)ww, (cuuint32_t)wh, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return et_fail(ET_ERR_CUDA, "cuTensorMapEncodeTiled(4-D window map) failed with CUresult %d", (int)r);
    return ET_OK;
}

}  // namespace

// Elements of bias scratch: (B * nwin * H, 256, 32) combined [8 bias_h | 8 bias_w | 0] rows + the (256, 64) one-hot block.
long long et_tc_window_scratch_elems(int B, int nwin, int H) { return (long long)B * nwin * H * 256 * 32 + 256 * 64; }

// bias_comb: scratch as above, already filled by relpos_bias_kernel in the combined window layout (or nullptr).
int et_tc_window_attention(const void* qkv, const void* pad_token, void* bias_comb, void* out, int B, int N, int gh, int gw,
                           int wh, int ww, int H, int has_bias, int is_bf16, cudaStream_t s) {
    WinArgs a;
    a.pad_token = pad_token; a.bias = bias_comb; a.out = out; a.B = B; a.N = N; a.gh = gh; a.gw = gw; a.wh = wh; a.ww = ww;
    a.nwy = (gh + wh - 1) / wh; a.nwx = (gw + ww - 1) / ww; a.H = H; a.D = H * 64; a.Wn = wh * ww;
    a.NK = (a.Wn + 15) / 16 * 16; a.has_bias = has_bias; a.is_bf16 = is_bf16; a.c1 = 0.125f * kLog2e;
    const int nwin = a.nwx * a.nwy;
    int rc;
    if ((rc = et_raise_smem(tc_window_kernel<true>, W_SMEM))) return rc;
    if ((rc = et_raise_smem(tc_window_kernel<false>, W_SMEM))) return rc;
    CUtensorMap tq, tb, toh;
    if ((rc = make_tmap_grid4d(&tq, qkv, B, gh, gw, 3 * a.D, wh, ww, is_bf16))) return rc;
    tb = toh = tq;
    if (has_bias) {
        uint16_t* oh = static_cast<uint16_t*>(bias_comb) + (size_t)B * nwin * H * 256 * 32;
        if ((rc = make_tmap_2d(&toh, oh, 256, 64, a.NK, is_bf16))) return rc;
        if (is_bf16) et_launch(window_onehot_kernel<true>, dim3(8), dim3(256), 0, s, oh, wh, ww);
        else et_launch(window_onehot_kernel<false>, dim3(8), dim3(256), 0, s, oh, wh, ww);
        ET_COUNT_LAUNCH(1);
    }
    const dim3 grid(B * nwin, H);
    if (is_bf16) et_launch(tc_window_kernel<true>, dim3(grid), dim3(kThreads), W_SMEM, s, tq, tb, toh, a);
    else et_launch(tc_window_kernel<false>, dim3(grid), dim3(kThreads), W_SMEM, s, tq, tb, toh, a);
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}
