// Gathered-row linear with scatter epilogue on 5th-gen tensor cores (sm_100a).
//
//   y = act(A @ W^T + bias);   out[scatter_row(m), :] = y[m, :]
//
// Warp-specialised, one 128 x BLOCK_N output tile per CTA:
//   warp 0    : TMA producer   (cp.async.bulk.tensor 2-D, 128-byte swizzle, K slices of 64)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, fp32 accumulate in TMEM)
//   warps 2-9 : epilogue       two warps per TMEM lane quarter, half of the tile's columns each:
//                               tcgen05.ld -> bias / GELU -> packed 16-byte chunks into a swizzled staging tile (the
//                               dead smem ring), then row-contiguous 16-byte stores (a warp writes whole output rows),
//                               rows redirected through the gate's index = TokenBuffer scatter
// smem ring of STAGES {A 128x64, W BLOCK_Nx64} tiles guarded by full/empty mbarriers; the MMA warp
// releases a stage with tcgen05.commit and signals the epilogue through a third barrier.
// TMA out-of-bounds zero fill handles the M / n_feat / K tails, so any M, K % 8 == 0 and
// n_feat % 8 == 0 are accepted.
#include "et_tcgen05.cuh"

int et_generic_linear(const void* A, int64_t M, int64_t K, const void* W, const void* bias, int64_t n_feat, int act, void* out,
                      int64_t ld_out, const int64_t* idx, const int32_t* count, int64_t k, int64_t n_out_rows,
                      cudaStream_t stream);

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 320;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;

struct LinearArgs {
    const void* bias;
    void* out;
    const long long* idx;    // scatter: output row of A row m is (m / k) * n_out_rows + idx[m]
    const long long* a_idx;  // gather:  A row m is row (m / k) * a_rows + a_idx[m] of the source tensor (TMA tile::gather4)
    const void* a_src;       // the gather source (R * a_rows, K) and, optionally, the gate state of the same shape that the
    void* state;             // CTAs of the first column of tiles advance: state[row] = a_src[row] for every gathered row
    const int* count;
    long long ld_out;
    int M, K, n_feat, act, k, n_out_rows, a_rows, is_bf16;
    unsigned long long* prof;
};

using namespace et_tc;

// Four rows of a 2-D tensor (tensor map with a one-row box), each 64 elements wide from column c0, land as four consecutive
// 128-byte rows of the destination in the 128B-swizzled K-major operand layout (profiles/microbench/gather4_probe.cu).
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1, int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}
// Source rows of the four A rows m .. m + 3 that lane `lane` of the producer warp gathers (rows past M or past the
// device-side count repeat a valid row: their results are never stored).
__device__ __forceinline__ void gather_rows(const LinearArgs& args, int m_first, int (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = min(m_first + i, args.M - 1);
        const int b = m / args.k, j = m - b * args.k;
        const bool valid = args.count == nullptr || j < args.count[b];
        r[i] = b * args.a_rows + (valid ? (int)args.a_idx[m] : 0);
    }
}

// Gate-state advance p[idx] = c[idx] (reference modules.py:151) for the 16 tile rows m_first .. m_first + 15, by one warp:
// plain 16-byte copies of the gathered source rows, issued by epilogue warps while the mainloop runs.
__device__ __forceinline__ void advance_state_rows(const LinearArgs& args, int m_first, int lane) {
    const uint4* src = static_cast<const uint4*>(args.a_src);
    uint4* dst = static_cast<uint4*>(args.state);
    const int cpr = args.K >> 3;  // 16-byte chunks per row
#pragma unroll 1
    for (int r = 0; r < 16; r += 4) {
        long long row[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m_first + r + i;
            row[i] = -1;
            if (m < args.M) {
                const int b = m / args.k, j = m - b * args.k;
                if (args.count == nullptr || j < args.count[b]) row[i] = ((long long)b * args.a_rows + args.a_idx[m]) * cpr;
            }
        }
        for (int c = lane; c < cpr; c += 32) {
            uint4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (row[i] >= 0) v[i] = src[row[i] + c];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (row[i] >= 0) dst[row[i] + c] = v[i];
        }
    }
}

// et_debug_set(7, device pointer to 3 x 16 u64): per-role cycle buckets of linear_tcgen05_kernel, summed over CTAs
// (profiling build only: make prof).
unsigned long long* g_gemm_prof = nullptr;
#ifdef ET_TC_PROFILE
#define GPF_DECL long long pf_[16]; for (int i_ = 0; i_ < 16; ++i_) pf_[i_] = 0; long long pf_t_ = clock64();
#define GPF(i) do { const long long n_ = clock64(); pf_[i] += n_ - pf_t_; pf_t_ = n_; } while (0)
#define GPF_FLUSH(role) do { if (args.prof) for (int i_ = 0; i_ < 16; ++i_) atomicAdd(args.prof + (role) * 16 + i_, (unsigned long long)pf_[i_]); } while (0)
#else
#define GPF_DECL
#define GPF(i)
#define GPF_FLUSH(role)
#endif

// GELU(x) = x Phi(x) with the exact-erf definition (torch.nn.GELU default), two elements at a time.
// Phi(x) = 1/2 + sign(x) (1/2 - r / 2), r = erfc(z), z = |x| / sqrt 2, by Abramowitz-Stegun 7.1.28:
//   erfc(z) = (1 + a1 z + ... + a6 z^6)^-16, |error| < 3e-7 (|error| of GELU < 1e-6 in fp32 arithmetic, far below the 16-bit
//   output rounding).  Packed FFMA2 / FMUL2 arithmetic and ONE MUFU (rcp) per element: the previous form (A&S 7.1.26: rcp + ex2)
//   needed two, and the GELU epilogue of mlp_1 was MUFU-bound (16 results/clk/SM) and longer than its mainloop could hide.
__device__ __forceinline__ void gelu_erf2(float& a, float& b) {
    const float kr = 0.70710678118654752440f;
    const f32x2 z = f2_mul(f2_pack(fabsf(a), fabsf(b)), f2_pack(kr, kr));
    f32x2 p = f2_fma(f2_pack(0.0000430638f, 0.0000430638f), z, f2_pack(0.0002765672f, 0.0002765672f));
    p = f2_fma(p, z, f2_pack(0.0001520143f, 0.0001520143f));
    p = f2_fma(p, z, f2_pack(0.0092705272f, 0.0092705272f));
    p = f2_fma(p, z, f2_pack(0.0422820123f, 0.0422820123f));
    p = f2_fma(p, z, f2_pack(0.0705230784f, 0.0705230784f));
    p = f2_fma(p, z, f2_pack(1.f, 1.f));
    p = f2_mul(p, p);
    p = f2_mul(p, p);
    p = f2_mul(p, p);
    p = f2_mul(p, p);  // (1 + ...)^16; overflows to +inf for |x| > ~40, whose reciprocal is the correct 0
    float pa, pb;
    f2_unpack(p, pa, pb);
    const f32x2 r = f2_pack(rcp_approx(pa), rcp_approx(pb));
    float sa, sb;  // 1/2 - r/2, carrying the sign of x
    f2_unpack(f2_fma(r, f2_pack(-0.5f, -0.5f), f2_pack(0.5f, 0.5f)), sa, sb);
    const f32x2 phi = f2_add(f2_pack(copysignf(sa, a), copysignf(sb, b)), f2_pack(0.5f, 0.5f));
    f2_unpack(f2_mul(f2_pack(a, b), phi), a, b);
}

template <int BLOCK_N, int STAGES, int MH>  // MH = 128-row halves per CTA tile (1 or 2): the halves share the W tile
struct GemmSmem {
    static constexpr int A_BYTES = MH * A_TILE_BYTES;
    static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_TILE_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int ROW_OFFSET = BAR_OFFSET + 256;     // int[128]: scatter target row of every row of a half
    static constexpr int BIAS_OFFSET = ROW_OFFSET + 512;    // the tile's BLOCK_N bias values (16-bit)
    static constexpr int TOTAL = BIAS_OFFSET + 512 + 1024;  // barriers + row table + bias + alignment slack
    static constexpr int OUT_STRIDE = (BLOCK_N * 2 + 127) / 128 * 128;  // staging tile row pitch (whole swizzle groups)
    static_assert(BLOCK_M * OUT_STRIDE <= STAGES * STAGE_BYTES, "epilogue staging tile must fit in the smem ring");
};

template <int BLOCK_N, int STAGES, int MH>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                      const LinearArgs args) {
    // Programmatic dependent launch: everything that does not depend on the previous kernel (barrier / TMEM setup and
    // the weight tiles of the first pipeline stages) runs before griddepcontrol.wait.
    et_pdl_trigger();
    using L = GemmSmem<BLOCK_N, STAGES, MH>;
    constexpr int ACC_COLS = MH * BLOCK_N;  // accumulator of half h: columns [h * BLOCK_N, +BLOCK_N)
    constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full_bar = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    GPF_DECL
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BLOCK_N;
    const int m0 = blockIdx.y * (BLOCK_M * MH);
    const int num_k_blocks = (args.K + BLOCK_K - 1) / BLOCK_K;
    if (args.count != nullptr) {
        // device-side row count (threshold policy): a tile made only of rows beyond it has nothing to compute or store
        et_pdl_wait();
        const int mlast = min(args.M, m0 + BLOCK_M * MH) - 1;
        const int b0 = m0 / args.k;
        if (mlast / args.k == b0 && m0 - b0 * args.k >= args.count[b0]) return;
    }

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(tmem_full_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // whole warp: allocate the accumulator columns, then let other CTAs allocate
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    GPF(0);  // prologue: barrier init, TMEM allocation, CTA sync

    if (warp == 0 && args.a_idx != nullptr) {
        // ---- gathered A operand: the whole warp produces; lane l fetches rows 4 l .. 4 l + 3 of each 128-row half with one
        // TMA gather4 per k-block straight from the gate state (the rows selected by the gate index), lane 0 fetches W
        const int pre = num_k_blocks < STAGES ? num_k_blocks : STAGES;
        if (lane == 0) {
            for (int kb = 0; kb < pre; ++kb) {
                const uint32_t fb = smem_u32(&full_bar[kb]);
                mbar_expect_tx(fb, L::STAGE_BYTES);
                tma_load_2d(smem_u32(smem + kb * L::STAGE_BYTES) + L::A_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
            }
        }
        et_pdl_wait();
        int rows[MH][4];
#pragma unroll
        for (int hh = 0; hh < MH; ++hh) gather_rows(args, m0 + hh * BLOCK_M + 4 * lane, rows[hh]);
        __syncwarp();
        for (int kb = 0; kb < num_k_blocks; ++kb) {
            const int s = kb % STAGES;
            const uint32_t a_dst = smem_u32(smem + s * L::STAGE_BYTES);
            const uint32_t fb = smem_u32(&full_bar[s]);
            if (kb >= pre) {
                mbar_wait(smem_u32(&empty_bar[s]), ((kb / STAGES) & 1) ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(fb, L::STAGE_BYTES);
                    tma_load_2d(a_dst + L::A_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
                }
                __syncwarp();
            }
#pragma unroll
            for (int hh = 0; hh < MH; ++hh)
                tma_gather4(a_dst + hh * A_TILE_BYTES + lane * 512, &tmap_a, fb, kb * BLOCK_K, rows[hh][0], rows[hh][1], rows[hh][2],
                            rows[hh][3]);
        }
    } else if (warp == 0) {
        if (lane == 0) {
            const int pre = num_k_blocks < STAGES ? num_k_blocks : STAGES;
            for (int kb = 0; kb < pre; ++kb) {  // weights are never written by a kernel: fetch them ahead of the wait
                const uint32_t fb = smem_u32(&full_bar[kb]);
                mbar_expect_tx(fb, L::STAGE_BYTES);
                tma_load_2d(smem_u32(smem + kb * L::STAGE_BYTES) + L::A_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
            }
            et_pdl_wait();
            for (int kb = 0; kb < pre; ++kb)
                tma_load_2d(smem_u32(smem + kb * L::STAGE_BYTES), &tmap_a, smem_u32(&full_bar[kb]), kb * BLOCK_K, m0);
            for (int kb = pre; kb < num_k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t parity = (kb / STAGES) & 1;
                GPF(2);
                mbar_wait(smem_u32(&empty_bar[s]), parity ^ 1);
                GPF(1);
                const uint32_t a_dst = smem_u32(smem + s * L::STAGE_BYTES);
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(fb, L::STAGE_BYTES);
                tma_load_2d(a_dst, &tmap_a, fb, kb * BLOCK_K, m0);
                tma_load_2d(a_dst + L::A_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
            }
            GPF(2);
            GPF_FLUSH(0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(BLOCK_N, args.is_bf16);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % STAGES;
                const uint32_t parity = (kb / STAGES) & 1;
                GPF(2);
                mbar_wait(smem_u32(&full_bar[s]), parity);
                if (kb == 0) GPF(3); else GPF(1);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
                const uint64_t db = umma_smem_desc(a_addr + L::A_BYTES);
#pragma unroll
                for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
#pragma unroll
                    for (int hh = 0; hh < MH; ++hh) {  // the 128-row halves alternate: two independent accumulators
                        const uint64_t da = umma_smem_desc(a_addr + hh * A_TILE_BYTES);
                        // advance 16 elements = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
                        tcgen05_mma_f16(tmem_base + hh * BLOCK_N, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc,
                                        (kb > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                tcgen05_commit(smem_u32(&empty_bar[s]));  // frees the smem stage once these MMAs retire
            }
            tcgen05_commit(smem_u32(tmem_full_bar));  // accumulator complete
            GPF(2);
            GPF_FLUSH(1);
        }
    } else {
        // ---- epilogue: warp w may only touch TMEM lanes [32 * (w % 4), +32)
        et_pdl_wait();  // the index, the output rows and (through TMA) the activations belong to earlier kernels
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int chalf = ew >> 2;  // which half of the tile's columns this warp converts
        const int row = quarter * 32 + lane;
        int* s_row = reinterpret_cast<int*>(smem + L::ROW_OFFSET);
        uint16_t* s_bias = reinterpret_cast<uint16_t*>(smem + L::BIAS_OFFSET);
        uint16_t* out = static_cast<uint16_t*>(args.out);
        // the tile's bias slice goes to shared memory once (a global load per 8 columns in the conversion loop exposed its
        // latency 12-16 times per thread)
        if (args.bias != nullptr && (int)threadIdx.x - 64 < BLOCK_N / 8) {
            const int t8 = ((int)threadIdx.x - 64) * 8;
            *reinterpret_cast<uint4*>(s_bias + t8) =
                n0 + t8 < args.n_feat ? *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(args.bias) + n0 + t8)
                                      : make_uint4(0, 0, 0, 0);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (args.state != nullptr && blockIdx.x == 0) {
#pragma unroll 1
            for (int hh = 0; hh < MH; ++hh) advance_state_rows(args, m0 + hh * BLOCK_M + ew * 16, lane);
        }
#pragma unroll 1
        for (int hh = 0; hh < MH; ++hh) {
        const int m = m0 + hh * BLOCK_M + row;
        if (hh > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // staging tile and row table of the previous half are read
        if (chalf == 0) {
            bool valid = m < args.M;
            long long out_row = m;
            if (valid && args.idx != nullptr) {
                const int b = m / args.k, j = m - b * args.k;
                if (args.count != nullptr && j >= args.count[b]) valid = false;
                if (valid) out_row = (long long)b * args.n_out_rows + args.idx[m];
            }
            s_row[row] = valid ? (int)out_row : -1;
        }
        __syncwarp();  // tcgen05.ld is warp-collective (.sync.aligned): reconverge after the guarded lookups
        GPF(4);
        if (hh == 0) mbar_wait(smem_u32(tmem_full_bar), 0);
        GPF(1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(hh * BLOCK_N);
        // phase 1: accumulator -> bias / activation -> 16-byte chunks in the staging tile.  Every MMA has retired, so
        // the smem ring is dead and is reused; chunk c of row r sits at position (c & ~7) | ((c ^ r) & 7).
        uint8_t* stage_row = smem + row * L::OUT_STRIDE;
        auto convert = [&](const uint32_t* acc, int c0, int width) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g * 8 < width) {
                    const int ng = n0 + c0 + g * 8;
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(acc[g * 8 + i]);
                    if (ng < args.n_feat) {
                        if (args.bias != nullptr) {
                            float bv[8];
                            const uint4 braw = *reinterpret_cast<const uint4*>(s_bias + c0 + g * 8);
                            if (args.is_bf16) unpack16<__nv_bfloat16>(braw, bv);
                            else unpack16<__half>(braw, bv);
#pragma unroll
                            for (int i = 0; i < 8; ++i) y[i] += bv[i];
                        }
                        if (args.act == ET_ACT_GELU) {
#pragma unroll
                            for (int i = 0; i < 8; i += 2) gelu_erf2(y[i], y[i + 1]);
                        }
                    }
                    const int c = (c0 >> 3) + g;
                    st16(stage_row + (((c & ~7) | ((c ^ row) & 7)) << 4), args.is_bf16 ? pack16<__nv_bfloat16>(y) : pack16<__half>(y));
                }
            }
        };
        constexpr int HALF_N = BLOCK_N / 2;
        const int cbeg = chalf * HALF_N;
#pragma unroll 1
        for (int c0 = cbeg; c0 + 32 <= cbeg + HALF_N; c0 += 32) {
            uint32_t acc[32];
            tmem_load_32x32(taddr + (uint32_t)c0, acc);
            GPF(2);
            convert(acc, c0, 32);
            GPF(3);
        }
        if constexpr (HALF_N % 32 != 0) {
            uint32_t acc[16];
            tmem_load_32x16(taddr + (uint32_t)(cbeg + HALF_N - 16), acc);
            GPF(2);
            convert(acc, cbeg + HALF_N - 16, 16);
            GPF(3);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // staging tile and row table complete (epilogue warps only)
        GPF(5);
        // phase 2: a warp writes whole output rows: LPR lanes x 16 bytes are one row's BLOCK_N columns
        constexpr int LPR = BLOCK_N / 8;
        constexpr int RPI = 32 / LPR > 0 ? 32 / LPR : 1;  // rows per warp-wide store
        const int sub = lane / LPR, c = lane % LPR;
        const int n = n0 + c * 8;
#pragma unroll 4
        for (int it = 0; it < 16 / RPI; ++it) {
            const int r = ew * 16 + it * RPI + sub;
            if (sub < RPI && n < args.n_feat) {
                const int orow = s_row[r];
                if (orow >= 0)
                    st16(out + (size_t)orow * args.ld_out + n,
                         ld16(smem + r * L::OUT_STRIDE + (((c & ~7) | ((c ^ r) & 7)) << 4)));
            }
        }
        GPF(6);
        }  // halves
        if (warp == 2 && lane == 0) GPF_FLUSH(2);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant for multi-wave problems (multi-stream batches): one CTA per SM walks over 128 x BLOCK_N output
// tiles (n fastest, so the CTAs of a wave share A rows and keep W resident in L2).  The smem ring runs across tile
// boundaries (the producer is already loading tile i+1 while tile i is multiplied), the accumulator is double buffered
// in TMEM (2 x BLOCK_N columns) and the epilogue of tile i overlaps the mainloop of tile i+1 inside the same CTA;
// barrier / TMEM / tensor-map setup is paid once per CTA instead of once per tile.
// ------------------------------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES>
struct PersistSmem {
    static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int OUT_STRIDE = (BLOCK_N * 2 + 127) / 128 * 128;
    static constexpr int STAGE_OFFSET = STAGES * STAGE_BYTES;          // epilogue staging tile (the ring stays live)
    static constexpr int BAR_OFFSET = STAGE_OFFSET + BLOCK_M * OUT_STRIDE;
    static constexpr int ROW_OFFSET = BAR_OFFSET + 256;
    static constexpr int BIAS_OFFSET = ROW_OFFSET + 512;
    static constexpr int TOTAL = BIAS_OFFSET + 512 + 1024;
    static_assert(TOTAL <= 227 * 1024, "persistent GEMM tile does not fit in shared memory");
};

// Epilogue warps of the persistent kernel: four per TMEM lane quarter (a quarter of the tile's columns each).  With two per
// quarter the epilogue (TMEM load -> bias / GELU -> pack -> staging -> row stores) kept its warps busy 70-90 % of a tile's time
// and the MMA issuer waited for accumulator buffers (profiles/r1_gemm_roles_persistent.txt); four per scheduler hide the
// TMEM-load, MUFU and store latencies of each other.
constexpr int kPersistEpiWarps = 16;
constexpr int kPersistThreads = 64 + kPersistEpiWarps * 32;
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kPersistThreads, 1)
linear_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         const LinearArgs args) {
    et_pdl_trigger();
    using L = PersistSmem<BLOCK_N, STAGES>;
    constexpr int TMEM_COLS = 2 * BLOCK_N <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;       // [2] accumulator buffer complete (MMA -> epilogue)
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2] accumulator buffer drained (8 epilogue warps -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_k_blocks = (args.K + BLOCK_K - 1) / BLOCK_K;
    const int n_tiles = (args.n_feat + BLOCK_N - 1) / BLOCK_N;
    const int m_tiles = (args.M + BLOCK_M - 1) / BLOCK_M;
    const int num_tiles = n_tiles * m_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int u = 0; u < 2; ++u) {
            mbar_init(smem_u32(&tmem_full[u]), 1);
            mbar_init(smem_u32(&tmem_empty[u]), kPersistEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && args.a_idx != nullptr) {
        // gathered A operand (see linear_tcgen05_kernel): the whole warp produces, lane l owns rows 4 l .. 4 l + 3 of the tile
        et_pdl_wait();
        int kbg = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
            int rows[4];
            gather_rows(args, m0 + 4 * lane, rows);
            for (int kb = 0; kb < num_k_blocks; ++kb, ++kbg) {
                const int s = kbg % STAGES;
                mbar_wait(smem_u32(&empty_bar[s]), ((kbg / STAGES) & 1) ^ 1);
                const uint32_t dst = smem_u32(smem + s * L::STAGE_BYTES);
                const uint32_t fb = smem_u32(&full_bar[s]);
                if (lane == 0) {
                    mbar_expect_tx(fb, L::STAGE_BYTES);
                    tma_load_2d(dst + A_TILE_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
                }
                __syncwarp();
                tma_gather4(dst + lane * 512, &tmap_a, fb, kb * BLOCK_K, rows[0], rows[1], rows[2], rows[3]);
            }
        }
    } else if (warp == 0) {
        if (lane == 0) {
            et_pdl_wait();
            GPF_DECL
            int kbg = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
                for (int kb = 0; kb < num_k_blocks; ++kb, ++kbg) {
                    const int s = kbg % STAGES;
                    GPF(2);
                    mbar_wait(smem_u32(&empty_bar[s]), ((kbg / STAGES) & 1) ^ 1);
                    GPF(1);
                    const uint32_t dst = smem_u32(smem + s * L::STAGE_BYTES);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, L::STAGE_BYTES);
                    tma_load_2d(dst, &tmap_a, fb, kb * BLOCK_K, m0);
                    tma_load_2d(dst + A_TILE_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
                }
            }
            GPF(2);
            GPF_FLUSH(0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(BLOCK_N, args.is_bf16);
            int kbg = 0, it = 0;
            GPF_DECL
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                GPF(2);
                mbar_wait(smem_u32(&tmem_empty[buf]), ((it >> 1) & 1) ^ 1);  // epilogue of tile it - 2 has drained it
                GPF(3);
                tcgen05_fence_after();
                for (int kb = 0; kb < num_k_blocks; ++kb, ++kbg) {
                    const int s = kbg % STAGES;
                    GPF(2);
                    mbar_wait(smem_u32(&full_bar[s]), (kbg / STAGES) & 1);
                    GPF(1);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
                    const uint64_t da = umma_smem_desc(a_addr);
                    const uint64_t db = umma_smem_desc(a_addr + A_TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                        tcgen05_mma_f16(tmem_base + buf * BLOCK_N, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc,
                                        (kb > 0 || kk > 0) ? 1u : 0u);
                    tcgen05_commit(smem_u32(&empty_bar[s]));
                }
                tcgen05_commit(smem_u32(&tmem_full[buf]));
            }
            GPF(2);
            GPF_FLUSH(1);
        }
    } else {
        et_pdl_wait();
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int chalf = ew >> 2;  // which part of the tile's columns this warp converts
        constexpr int CPARTS = kPersistEpiWarps / 4;
        const int row = quarter * 32 + lane;
        int* s_row = reinterpret_cast<int*>(smem + L::ROW_OFFSET);
        uint8_t* stage = smem + L::STAGE_OFFSET;
        uint8_t* stage_row = stage + row * L::OUT_STRIDE;
        uint16_t* out = static_cast<uint16_t*>(args.out);
        uint16_t* s_bias = reinterpret_cast<uint16_t*>(smem + L::BIAS_OFFSET);
        auto load_bias = [&](int tile) {  // the tile's bias slice -> shared memory (read back as broadcast LDS.128)
            const int t8 = ((int)threadIdx.x - 64) * 8;
            if (args.bias != nullptr && tile < num_tiles && t8 < BLOCK_N) {
                const int ng = (tile % n_tiles) * BLOCK_N + t8;
                *reinterpret_cast<uint4*>(s_bias + t8) =
                    ng < args.n_feat ? *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(args.bias) + ng)
                                     : make_uint4(0, 0, 0, 0);
            }
        };
        load_bias(blockIdx.x);
        asm volatile("bar.sync 1, %0;" ::"n"(kPersistEpiWarps * 32) : "memory");
        int it = 0;
        GPF_DECL
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
            const int m = m0 + row;
            GPF(6);
            if (chalf == 0) {
                bool valid = m < args.M;
                long long out_row = m;
                if (valid && args.idx != nullptr) {
                    const int b = m / args.k, j = m - b * args.k;
                    if (args.count != nullptr && j >= args.count[b]) valid = false;
                    if (valid) out_row = (long long)b * args.n_out_rows + args.idx[m];
                }
                s_row[row] = valid ? (int)out_row : -1;
            }
            if (args.state != nullptr && n0 == 0) advance_state_rows(args, m0 + ew * 16, lane);
            __syncwarp();
            GPF(4);
            mbar_wait(smem_u32(&tmem_full[buf]), (it >> 1) & 1);
            GPF(1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BLOCK_N);
            auto convert = [&](const uint32_t* acc, int c0, int width) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g * 8 < width) {
                        const int ng = n0 + c0 + g * 8;
                        float y[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(acc[g * 8 + i]);
                        if (ng < args.n_feat) {
                            if (args.bias != nullptr) {
                                float bv[8];
                                const uint4 braw = *reinterpret_cast<const uint4*>(s_bias + c0 + g * 8);
                                if (args.is_bf16) unpack16<__nv_bfloat16>(braw, bv);
                                else unpack16<__half>(braw, bv);
#pragma unroll
                                for (int i = 0; i < 8; ++i) y[i] += bv[i];
                            }
                            if (args.act == ET_ACT_GELU) {
#pragma unroll
                                for (int i = 0; i < 8; i += 2) gelu_erf2(y[i], y[i + 1]);
                            }
                        }
                        const int c = (c0 >> 3) + g;
                        st16(stage_row + (((c & ~7) | ((c ^ row) & 7)) << 4),
                             args.is_bf16 ? pack16<__nv_bfloat16>(y) : pack16<__half>(y));
                    }
                }
            };
            constexpr int HALF_N = BLOCK_N / CPARTS;
            const int cbeg = chalf * HALF_N;
#pragma unroll 1
            for (int c0 = cbeg; c0 + 32 <= cbeg + HALF_N; c0 += 32) {
                uint32_t acc[32];
                tmem_load_32x32(taddr + (uint32_t)c0, acc);
                if (c0 + 64 > cbeg + HALF_N && HALF_N % 32 == 0) {  // last read of this buffer: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[buf]));
                }
                convert(acc, c0, 32);
            }
            if constexpr (HALF_N % 32 != 0) {
                uint32_t acc[16];
                tmem_load_32x16(taddr + (uint32_t)(cbeg + HALF_N - 16), acc);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[buf]));
                convert(acc, cbeg + HALF_N - 16, 16);
            }
            GPF(3);
            asm volatile("bar.sync 1, %0;" ::"n"(kPersistEpiWarps * 32) : "memory");  // staging tile and row table complete, bias slice consumed
            GPF(5);
            load_bias(tile + (int)gridDim.x);  // next tile's slice; visible after the trailing barrier
            constexpr int LPR = BLOCK_N / 8;
            constexpr int RPI = 32 / LPR > 0 ? 32 / LPR : 1;
            constexpr int ROWS_PER_WARP = BLOCK_M / kPersistEpiWarps;
            constexpr int NIT = ROWS_PER_WARP / RPI;
            const int sub = lane / LPR, c = lane % LPR;
            const int n = n0 + c * 8;
            // all shared-memory reads first, then the stores: one latency instead of NIT
            uint4 rowv[NIT];
            int orow[NIT];
#pragma unroll
            for (int r_it = 0; r_it < NIT; ++r_it) {
                const int r = ew * ROWS_PER_WARP + r_it * RPI + (sub < RPI ? sub : 0);
                orow[r_it] = (sub < RPI && n < args.n_feat) ? s_row[r] : -1;
                rowv[r_it] = ld16(stage + r * L::OUT_STRIDE + (((c & ~7) | ((c ^ r) & 7)) << 4));
            }
#pragma unroll
            for (int r_it = 0; r_it < NIT; ++r_it)
                if (orow[r_it] >= 0) st16(out + (size_t)orow[r_it] * args.ld_out + n, rowv[r_it]);
            asm volatile("bar.sync 1, %0;" ::"n"(kPersistEpiWarps * 32) : "memory");  // staging tile and row table may be rewritten
        }
        GPF(6);
        if (warp == 2 && lane == 0) GPF_FLUSH(2);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent kernel (tcgen05 cta_group::2): two CTAs of a cluster -- two SMs of one TPC -- compute
// one 256 x 256 output tile.  Each CTA stages its own 128 rows of A and HALF of the W tile (128 of the 256 output features);
// the leader CTA's single MMA thread issues M = 256 MMAs that read A from both CTAs' shared memory and the two W halves,
// and each CTA's tensor memory receives its 128 rows x 256 columns of the accumulator.  Per k-block a CTA therefore loads
// 32 KB instead of 48 KB for the same 128 x 256 outputs: 64 B/clk instead of 96 B/clk of L2 -> SM ingress, which is what kept
// the single-CTA kernel's MMA issuer waiting 35 % of the time (profiles/r1_gemm_roles_persistent.txt).
//   full[s]       leader's barrier: TMA bytes of BOTH CTAs land on it (cp.async.bulk.tensor ... cta_group::2, barrier address
//                 with the peer bit cleared); one arrive.expect_tx by the leader's producer
//   empty[s]      per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tmem_full[u]  per CTA, same multicast commit;  tmem_empty[u]  leader's barrier, 8 + 8 epilogue warps of both CTAs arrive
// ------------------------------------------------------------------------------------------------------------------
constexpr int PAIR_N = 256;
template <int STAGES>
struct PairSmem {
    static constexpr int W_HALF_BYTES = (PAIR_N / 2) * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + W_HALF_BYTES;   // per CTA
    static constexpr int OUT_STRIDE = PAIR_N * 2;
    static constexpr int STAGE_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = STAGE_OFFSET + BLOCK_M * OUT_STRIDE;
    static constexpr int ROW_OFFSET = BAR_OFFSET + 256;
    static constexpr int BIAS_OFFSET = ROW_OFFSET + 512;
    static constexpr int TOTAL = BIAS_OFFSET + 512 + 1024;
    static_assert(TOTAL <= 227 * 1024, "pair GEMM tile does not fit in shared memory");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load whose completion bytes are credited to the LEADER CTA's mbarrier (address with the peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at this shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(kPersistThreads, 1)
linear_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const LinearArgs args) {
    et_pdl_trigger();
    using L = PairSmem<STAGES>;
    constexpr int TMEM_COLS = 512;  // two 256-column accumulator buffers
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_k_blocks = (args.K + BLOCK_K - 1) / BLOCK_K;
    const int n_tiles = (args.n_feat + PAIR_N - 1) / PAIR_N;
    const int m_tiles = (args.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const int num_tiles = n_tiles * m_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int u = 0; u < 2; ++u) {
            mbar_init(smem_u32(&tmem_full[u]), 1);
            mbar_init(smem_u32(&tmem_empty[u]), 2 * kPersistEpiWarps);  // the epilogue warps of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();  // both CTAs' barriers are initialised and the pair's tensor memory is allocated
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            et_pdl_wait();
            int kbg = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m0 = (tile / n_tiles) * 2 * BLOCK_M + (int)rank * BLOCK_M;
                const int n0 = (tile % n_tiles) * PAIR_N + (int)rank * (PAIR_N / 2);
                for (int kb = 0; kb < num_k_blocks; ++kb, ++kbg) {
                    const int s = kbg % STAGES;
                    mbar_wait(smem_u32(&empty_bar[s]), ((kbg / STAGES) & 1) ^ 1);
                    const uint32_t dst = smem_u32(smem + s * L::STAGE_BYTES);
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    if (rank == 0) mbar_expect_tx(fb, 2 * L::STAGE_BYTES);  // both CTAs' bytes
                    tma_load_2d_pair(dst, &tmap_a, fb, kb * BLOCK_K, m0);
                    tma_load_2d_pair(dst + A_TILE_BYTES, &tmap_w, fb, kb * BLOCK_K, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            uint32_t idesc = umma_idesc(PAIR_N, args.is_bf16);
            idesc = (idesc & ~(0x1fu << 24)) | ((uint32_t)(256 >> 4) << 24);  // M = 256 across the pair
            int kbg = 0, it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int buf = it & 1;
                mbar_wait(smem_u32(&tmem_empty[buf]), ((it >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this buffer
                tcgen05_fence_after();
                for (int kb = 0; kb < num_k_blocks; ++kb, ++kbg) {
                    const int s = kbg % STAGES;
                    mbar_wait(smem_u32(&full_bar[s]), (kbg / STAGES) & 1);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
                    const uint64_t da = umma_smem_desc(a_addr);
                    const uint64_t db = umma_smem_desc(a_addr + A_TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                        tcgen05_mma_f16_pair(tmem_base + buf * PAIR_N, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc,
                                             (kb > 0 || kk > 0) ? 1u : 0u);
                    tcgen05_commit_pair(smem_u32(&empty_bar[s]));
                }
                tcgen05_commit_pair(smem_u32(&tmem_full[buf]));
            }
        }
    } else {
        et_pdl_wait();
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int chalf = ew >> 2;  // which quarter of the tile's columns this warp converts (four warps per TMEM lane quarter)
        constexpr int CPARTS = kPersistEpiWarps / 4;
        const int row = quarter * 32 + lane;
        int* s_row = reinterpret_cast<int*>(smem + L::ROW_OFFSET);
        uint8_t* stage = smem + L::STAGE_OFFSET;
        uint8_t* stage_row = stage + row * L::OUT_STRIDE;
        uint16_t* out = static_cast<uint16_t*>(args.out);
        uint16_t* s_bias = reinterpret_cast<uint16_t*>(smem + L::BIAS_OFFSET);
        auto load_bias = [&](int tile) {
            const int t8 = ((int)threadIdx.x - 64) * 8;
            if (args.bias != nullptr && tile < num_tiles && t8 < PAIR_N) {
                const int ng = (tile % n_tiles) * PAIR_N + t8;
                *reinterpret_cast<uint4*>(s_bias + t8) =
                    ng < args.n_feat ? *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(args.bias) + ng)
                                     : make_uint4(0, 0, 0, 0);
            }
        };
        load_bias(pair);
        asm volatile("bar.sync 1, %0;" ::"n"(kPersistEpiWarps * 32) : "memory");
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            const int buf = it & 1;
            const int m0 = (tile / n_tiles) * 2 * BLOCK_M + (int)rank * BLOCK_M, n0 = (tile % n_tiles) * PAIR_N;
            const int m = m0 + row;
            if (chalf == 0) {
                bool valid = m < args.M;
                long long out_row = m;
                if (valid && args.idx != nullptr) {
                    const int b = m / args.k, j = m - b * args.k;
                    out_row = (long long)b * args.n_out_rows + args.idx[m];
                }
                s_row[row] = valid ? (int)out_row : -1;
            }
            __syncwarp();
            mbar_wait(smem_u32(&tmem_full[buf]), (it >> 1) & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * PAIR_N);
            constexpr int HALF_N = PAIR_N / CPARTS;
            const int cbeg = chalf * HALF_N;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + HALF_N; c0 += 32) {
                uint32_t acc[32];
                tmem_load_32x32(taddr + (uint32_t)c0, acc);
                if (c0 + 32 == cbeg + HALF_N) {  // last read of this buffer: hand it back to the leader's MMA thread
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(smem_u32(&tmem_empty[buf]), 0);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int ng = n0 + c0 + g * 8;
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(acc[g * 8 + i]);
                    if (ng < args.n_feat) {
                        if (args.bias != nullptr) {
                            float bv[8];
                            const uint4 braw = *reinterpret_cast<const uint4*>(s_bias + c0 + g * 8);
                            if (args.is_bf16) unpack16<__nv_bfloat16>(braw, bv);
                            else unpack16<__half>(braw, bv);
#pragma unroll
                            for (int i = 0; i < 8; ++i) y[i] += bv[i];
                        }
                        if (args.act == ET_ACT_GELU) {
#pragma unroll
                            for (int i = 0; i < 8; i += 2) gelu_erf2(y[i], y[i + 1]);
                        }
                    }
                    const int c = (c0 >> 3) + g;
                    st16(stage_row + (((c & ~7) | ((c ^ row) & 7)) << 4), args.is_bf16 ? pack16<__nv_bfloat16>(y) : pack16<__half>(y));
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kPersistEpiWarps * 32) : "memory");  // staging tile and row table complete, bias slice consumed
            load_bias(tile + num_pairs);
            constexpr int NIT = BLOCK_M / kPersistEpiWarps;  // a warp-wide store writes one row's 256 columns
            const int n = n0 + lane * 8;
            uint4 rowv[NIT];
            int orow[NIT];
#pragma unroll
            for (int r_it = 0; r_it < NIT; ++r_it) {
                const int r = ew * NIT + r_it;
                orow[r_it] = n < args.n_feat ? s_row[r] : -1;
                rowv[r_it] = ld16(stage + r * L::OUT_STRIDE + (((lane & ~7) | ((lane ^ r) & 7)) << 4));
            }
#pragma unroll
            for (int r_it = 0; r_it < NIT; ++r_it)
                if (orow[r_it] >= 0) st16(out + (size_t)orow[r_it] * args.ld_out + n, rowv[r_it]);
            asm volatile("bar.sync 1, %0;" ::"n"(kPersistEpiWarps * 32) : "memory");
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // both CTAs are done with the pair's tensor memory and with each other's barriers
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side
template <int BLOCK_N, int STAGES, int MH = 1>
int launch_linear(const void* A, const void* W, const LinearArgs& args, cudaStream_t stream) {
    using L = GemmSmem<BLOCK_N, STAGES, MH>;
    int rc = et_raise_smem(linear_tcgen05_kernel<BLOCK_N, STAGES, MH>, L::TOTAL);
    if (rc) return rc;
    CUtensorMap ta, tw;
    rc = args.a_idx ? make_tmap_2d(&ta, A, (long long)(args.M / args.k) * args.a_rows, args.K, 1, args.is_bf16)  // one-row box: gather4
                    : make_tmap_2d(&ta, A, args.M, args.K, BLOCK_M * MH, args.is_bf16);
    if (rc) return rc;
    rc = make_tmap_2d(&tw, W, args.n_feat, args.K, BLOCK_N, args.is_bf16);
    if (rc) return rc;
    dim3 grid((args.n_feat + BLOCK_N - 1) / BLOCK_N, (args.M + BLOCK_M * MH - 1) / (BLOCK_M * MH));
    et_launch(linear_tcgen05_kernel<BLOCK_N, STAGES, MH>, dim3(grid), dim3(kGemmThreads), L::TOTAL, stream, ta, tw, args);
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}

template <int BLOCK_N, int STAGES>
int launch_persistent(const void* A, const void* W, const LinearArgs& args, cudaStream_t stream) {
    using L = PersistSmem<BLOCK_N, STAGES>;
    int rc = et_raise_smem(linear_persistent_kernel<BLOCK_N, STAGES>, L::TOTAL);
    if (rc) return rc;
    const int sms = et_sm_count();
    CUtensorMap ta, tw;
    rc = args.a_idx ? make_tmap_2d(&ta, A, (long long)(args.M / args.k) * args.a_rows, args.K, 1, args.is_bf16)
                    : make_tmap_2d(&ta, A, args.M, args.K, BLOCK_M, args.is_bf16);
    if (rc) return rc;
    rc = make_tmap_2d(&tw, W, args.n_feat, args.K, BLOCK_N, args.is_bf16);
    if (rc) return rc;
    const long long tiles = (long long)((args.n_feat + BLOCK_N - 1) / BLOCK_N) * ((args.M + BLOCK_M - 1) / BLOCK_M);
    const int grid = (int)(tiles < sms ? tiles : sms);
    et_launch(linear_persistent_kernel<BLOCK_N, STAGES>, dim3(grid), dim3(kPersistThreads), L::TOTAL, stream, ta, tw, args);
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}


template <int STAGES>
int launch_pair(const void* A, const void* W, const LinearArgs& args, cudaStream_t stream) {
    using L = PairSmem<STAGES>;
    int rc = et_raise_smem(linear_pair_kernel<STAGES>, L::TOTAL);
    if (rc) return rc;
    const int sms = et_sm_count();
    CUtensorMap ta, tw;
    rc = make_tmap_2d(&ta, A, args.M, args.K, BLOCK_M, args.is_bf16);
    if (rc) return rc;
    rc = make_tmap_2d(&tw, W, args.n_feat, args.K, PAIR_N / 2, args.is_bf16);
    if (rc) return rc;
    const long long tiles = (long long)((args.n_feat + PAIR_N - 1) / PAIR_N) * ((args.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
    const int pairs = (int)(tiles < sms / 2 ? tiles : sms / 2);
    et_launch_cluster(linear_pair_kernel<STAGES>, dim3(2 * pairs), dim3(kPersistThreads), L::TOTAL, stream, 2, ta, tw, args);
    ET_COUNT_LAUNCH(1);
    return ET_OK;
}

int g_force_pair = 0;  // et_debug_set(13, 1 = CTA-pair kernel whenever it applies, 2 = never, 0 = auto)
int g_force_persist = 0;  // et_debug_set(9, 1 = always persistent, 2 = never, 0 = auto)
int g_force_mh = 0;  // et_debug_set(8, 1 | 2): rows per CTA tile = 128 x value (0 = auto)
int g_force_block_n = 0;
int g_force_depth = 0;  // test / tuning hook: et_debug_set(5, 1 = deep pipelines, 2 = shallow (two CTAs per SM), 0 = auto)  // test hook: et_debug_set(1, BLOCK_N)

}  // namespace

extern int g_attn_tc;
extern int g_attn_win_gen;
extern int g_tc_apply_cluster;
extern unsigned long long* g_gate_dbg;
extern int g_tc_time_apply;
extern unsigned long long* g_tc_prof;
extern int g_gate_cta_waves;

extern "C" {

int et_debug_set(int key, long long value) {
    if (key == 1) {
        g_force_block_n = (int)value;
        return ET_OK;
    }
    if (key == 6) {
        g_tc_time_apply = (int)value;
        return ET_OK;
    }
    if (key == 5) {
        g_force_depth = (int)value;
        return ET_OK;
    }
    if (key == 8) {
        g_force_mh = (int)value;
        return ET_OK;
    }
    if (key == 9) {
        g_force_persist = (int)value;
        return ET_OK;
    }
    if (key == 10) {
        g_gate_cta_waves = value > 0 ? (int)value : 2;
        return ET_OK;
    }
    if (key == 2) {
        g_attn_tc = value != 0;
        return ET_OK;
    }
    if (key == 12) {
        g_tc_apply_cluster = value >= 1 && value <= 8 ? (int)value : 1;
        return ET_OK;
    }
    if (key == 11) {
        g_attn_win_gen = value == 1 ? 1 : 2;
        return ET_OK;
    }
    if (key == 7) {
        g_gemm_prof = reinterpret_cast<unsigned long long*>(value);
        return ET_OK;
    }
    if (key == 13) {
        g_force_pair = (int)value;
        return ET_OK;
    }
    if (key == 4) {
        g_tc_prof = reinterpret_cast<unsigned long long*>(value);
        return ET_OK;
    }
    if (key == 3) {
        g_gate_dbg = reinterpret_cast<unsigned long long*>(value);
        return ET_OK;
    }
    return et_fail(ET_ERR_ARG, "et_debug_set: unknown key %d", key);
}

// Shared body of et_linear / et_linear_gather.  a_idx != nullptr: A is the gather source (M / k * a_rows, K) and the A operand
// rows are A[(m / k) * a_rows + a_idx[m]] (TMA gather4 producer); `state` (same shape as the source) is advanced at those rows.
static int linear_impl(const void* A, const int64_t* a_idx, int64_t a_rows, void* state, int64_t M, int64_t K, const void* W,
                       const void* bias, int64_t n_feat, int act, void* out, int64_t ld_out, const int64_t* idx,
                       const int32_t* count, int64_t k, int64_t n_out_rows, int dtype, void* stream) {
    ET_CHECK_ARG(A && W && out, "et_linear: null pointer");
    if (dtype == ET_F32) {  // fp32 models: CUDA-core SGEMM with the same epilogue (et_generic.cu)
        if (a_idx != nullptr) return et_fail(ET_ERR_UNSUPPORTED, "et_linear_gather: 16-bit dtypes only (fp32 models gather with et_gate_gather)");
        ET_CHECK_ARG(M >= 0 && K > 0 && n_feat > 0 && M < (1LL << 31), "et_linear: bad shape");
        ET_CHECK_ARG(et_aligned16(A) && et_aligned16(W), "et_linear: pointers must be 16-byte aligned");
        ET_CHECK_ARG(act == ET_ACT_NONE || act == ET_ACT_GELU, "et_linear: unknown activation %d", act);
        if (idx != nullptr) ET_CHECK_ARG(k > 0 && M % k == 0 && n_out_rows > 0, "et_linear: scatter needs k | M and n_out_rows");
        ET_CHECK_ARG(count == nullptr || idx != nullptr, "et_linear: count needs idx");
        if (M == 0) return ET_OK;
        int rc32 = et_generic_linear(A, M, K, W, bias, n_feat, act, out, ld_out, idx, count, k, n_out_rows, et_stream(stream));
        if (rc32) return rc32;
        ET_CHECK_LAUNCH("et_linear");
        return ET_OK;
    }
    ET_CHECK_ARG(dtype == ET_BF16 || dtype == ET_F16, "et_linear: unknown dtype %d", dtype);
    ET_CHECK_ARG(M >= 0 && K > 0 && n_feat > 0 && M < (1LL << 31) && K % 8 == 0 && n_feat % 8 == 0 && ld_out % 8 == 0,
                 "et_linear: need K, n_feat, ld_out multiples of 8 (M=%lld K=%lld n_feat=%lld ld_out=%lld)",
                 (long long)M, (long long)K, (long long)n_feat, (long long)ld_out);
    ET_CHECK_ARG(et_aligned16(A) && et_aligned16(W) && et_aligned16(out) && et_aligned16(bias),
                 "et_linear: pointers must be 16-byte aligned");
    ET_CHECK_ARG(act == ET_ACT_NONE || act == ET_ACT_GELU, "et_linear: unknown activation %d", act);
    if (idx != nullptr) ET_CHECK_ARG(k > 0 && M % k == 0 && n_out_rows > 0, "et_linear: scatter needs k | M and n_out_rows");
    if (a_idx != nullptr)
        ET_CHECK_ARG(k > 0 && M % k == 0 && a_rows > 0 && (M / k) * a_rows < (1LL << 31) && et_aligned16(state),
                     "et_linear_gather: gather needs k | M, a_rows and a source of fewer than 2^31 rows");
    ET_CHECK_ARG(count == nullptr || idx != nullptr || a_idx != nullptr, "et_linear: count needs idx");
    if (M == 0) return ET_OK;
    LinearArgs a;
    a.bias = bias; a.out = out; a.idx = reinterpret_cast<const long long*>(idx); a.count = count; a.ld_out = ld_out;
    a.a_idx = reinterpret_cast<const long long*>(a_idx); a.a_rows = (int)a_rows; a.a_src = A; a.state = state;
    a.M = (int)M; a.K = (int)K; a.n_feat = (int)n_feat; a.act = act; a.k = (int)((idx || a_idx) ? k : 1);
    a.n_out_rows = (int)n_out_rows; a.is_bf16 = dtype == ET_BF16;
    a.prof = g_gemm_prof;

    // Tile width: minimise waves(tiles over the SMs of this device) x per-tile cost (~ BLOCK_N + fixed overhead).
    const long long sms = et_sm_count();
    const int candidates[5] = {256, 192, 128, 96, 64};
    int best = 64;
    double best_cost = 1e30;
    const long long mt = (M + BLOCK_M - 1) / BLOCK_M;
    for (int c = 0; c < 5; ++c) {
        const int bn = candidates[c];
        if (bn > 64 && bn >= 2 * n_feat) continue;
        const long long tiles = mt * ((n_feat + bn - 1) / bn);
        const long long slots = tiles > sms ? 2 * sms : sms;  // two CTAs per SM with the shallow pipelines
        const double cost = (double)((tiles + slots - 1) / slots) * (bn + 32) * (tiles > sms ? 1.25 : 1.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    if (g_force_block_n) best = g_force_block_n;
    int rc;
    cudaStream_t s = et_stream(stream);
    // Two pipeline depths per tile width: "deep" (one CTA per SM, longest TMA prefetch) and "shallow" (<= 110 KB of
    // smem so two CTAs share an SM: the epilogue of one overlaps the mainloop of the other and a launch of up to 296
    // tiles is a single wave).  Shallow is used when the tile count exceeds one CTA-per-SM wave.
    const long long tiles_best = mt * ((n_feat + best - 1) / best);
    const bool shallow = g_force_depth ? g_force_depth == 2 : tiles_best > sms;
    // 256-row CTA tiles (two 128-row accumulators sharing one W tile, one CTA per SM): a third less operand traffic from
    // L2, but no second CTA whose mainloop hides the epilogue.  Measured (profiles/r1_gemm_sweep.txt): only long-K
    // layers of multi-stream batches gain (mlp_2 at M = 16384: 84 -> 80 us); everything else keeps 128-row tiles.
    // CTA-pair kernel (cta_group::2, 256 x 256 tiles): multi-wave problems without device-side counts / gathered operands
    {
        const long long pair_tiles = ((M + 255) / 256) * ((n_feat + PAIR_N - 1) / PAIR_N);
        const bool eligible = count == nullptr && a_idx == nullptr && n_feat >= PAIR_N;
        const bool pair = eligible && (g_force_pair ? g_force_pair == 1 : (M >= 12288 && act == ET_ACT_NONE && pair_tiles >= sms && g_force_mh == 0 && g_force_persist == 0 && g_force_block_n == 0));
        if (pair) {
            rc = launch_pair<4>(A, W, a, s);
            if (rc) return rc;
            ET_CHECK_LAUNCH("et_linear");
            return ET_OK;
        }
    }
    // persistent kernel: multi-wave problems (at least two waves of 128 x 256 tiles); measured in profiles/r1_gemm_sweep.txt
    {
        const int pbn = g_force_block_n == 128 || g_force_block_n == 192 ? g_force_block_n : 256;
        const long long ptiles = mt * ((n_feat + pbn - 1) / pbn);
        const bool long_k = K >= 2048;  // long-K layers do better with 256-row tiles (below)
        const bool persist = g_force_persist ? g_force_persist == 1 : (ptiles >= 2 * sms && g_force_mh == 0 && !long_k && count == nullptr);
        if (persist) {
            switch (pbn) {
                case 128: rc = launch_persistent<128, 5>(A, W, a, s); break;
                case 192: rc = launch_persistent<192, 4>(A, W, a, s); break;
                default: rc = launch_persistent<256, 3>(A, W, a, s); break;
            }
            if (rc) return rc;
            ET_CHECK_LAUNCH("et_linear");
            return ET_OK;
        }
    }
    int mh = g_force_mh;
    int bn2 = best >= 192 ? best : 192;
    if (g_force_block_n) bn2 = g_force_block_n;
    if (mh == 0) mh = (K >= 2048 && ((M + 255) / 256) * ((n_feat + bn2 - 1) / bn2) >= 128) ? 2 : 1;
    if (mh == 2 && (bn2 == 128 || bn2 == 192 || bn2 == 256)) {
        switch (bn2) {
            case 128: rc = launch_linear<128, 2, 2>(A, W, a, s); break;
            case 192: rc = launch_linear<192, 3, 2>(A, W, a, s); break;
            default: rc = launch_linear<256, 3, 2>(A, W, a, s); break;
        }
    } else
    switch (best) {
        case 256: rc = shallow ? launch_linear<256, 2>(A, W, a, s) : launch_linear<256, 4>(A, W, a, s); break;
        case 192: rc = shallow ? launch_linear<192, 2>(A, W, a, s) : launch_linear<192, 5>(A, W, a, s); break;
        case 128: rc = shallow ? launch_linear<128, 3>(A, W, a, s) : launch_linear<128, 6>(A, W, a, s); break;
        case 96: rc = shallow ? launch_linear<96, 3>(A, W, a, s) : launch_linear<96, 7>(A, W, a, s); break;
        case 64: rc = shallow ? launch_linear<64, 4>(A, W, a, s) : launch_linear<64, 8>(A, W, a, s); break;
        default: return et_fail(ET_ERR_ARG, "et_linear: unsupported BLOCK_N %d", best);
    }
    if (rc) return rc;
    ET_CHECK_LAUNCH("et_linear");
    return ET_OK;
}

int et_linear(const void* A, int64_t M, int64_t K, const void* W, const void* bias, int64_t n_feat, int act, void* out,
              int64_t ld_out, const int64_t* idx, const int32_t* count, int64_t k, int64_t n_out_rows, int dtype,
              void* stream) {
    return linear_impl(A, nullptr, 0, nullptr, M, K, W, bias, n_feat, act, out, ld_out, idx, count, k, n_out_rows, dtype, stream);
}

int et_linear_gather(const void* a_src, int64_t a_rows, const int64_t* a_idx, void* state, int64_t M, int64_t K, const void* W,
                     const void* bias, int64_t n_feat, int act, void* out, int64_t ld_out, const int64_t* idx,
                     const int32_t* count, int64_t k, int64_t n_out_rows, int dtype, void* stream) {
    ET_CHECK_ARG(a_src && a_idx, "et_linear_gather: null pointer");
    return linear_impl(a_src, a_idx, a_rows, state, M, K, W, bias, n_feat, act, out, ld_out, idx, count, k, n_out_rows, dtype,
                       stream);
}

}  // extern "C"
