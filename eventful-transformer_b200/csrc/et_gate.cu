// Gate / buffer kernels: HBM-bound token streaming with 128-bit loads and warp-shuffle reductions.
//
//   gate_select_kernel  : [add] -> [LayerNorm] -> delta vs. reference state -> per-token L2 norm
//                         -> block-wide radix top-k (or threshold compaction), one launch.  The
//                         norms are produced by a grid of CTAs; the last CTA of each row to
//                         finish (atomic ticket) runs the selection from shared memory.
//   gate_gather_kernel  : c~ = c[idx] (LayerNorm recomputed in registers), e~, p[idx] = c~
//   buffer_scatter      : TokenBuffer row / column scatter
//   add                 : residual add
#include <cstdlib>
#include <mutex>
#include <vector>

#include "et_common.cuh"

thread_local char g_et_error[512] = "";
long long g_et_launches = 0;

namespace {
constexpr int kMaxDevices = 64;
std::mutex g_dev_mutex;
int g_sm_count[kMaxDevices];
struct SmemGrant { const void* fn; int device; int bytes; };
std::vector<SmemGrant> g_smem_grants;
}  // namespace

int et_sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (g_sm_count[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        g_sm_count[dev] = sms;
    }
    return g_sm_count[dev];
}

int et_raise_smem_impl(const void* kernel, int bytes) {
    if (bytes > 227 * 1024) return et_fail(ET_ERR_UNSUPPORTED, "kernel needs %d bytes of shared memory (> 227 KB)", bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    for (SmemGrant& g : g_smem_grants)
        if (g.fn == kernel && g.device == dev) {
            if (g.bytes >= bytes) return ET_OK;
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
            if (e != cudaSuccess) return et_fail(ET_ERR_CUDA, "cudaFuncSetAttribute(%d bytes): %s", bytes, cudaGetErrorString(e));
            g.bytes = bytes;
            return ET_OK;
        }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return et_fail(ET_ERR_CUDA, "cudaFuncSetAttribute(%d bytes): %s", bytes, cudaGetErrorString(e));
    g_smem_grants.push_back({kernel, dev, bytes});
    return ET_OK;
}
int g_gate_cta_waves = 2;  // CTAs of the gate kernels per SM over the whole launch (et_debug_set key 10)
int g_et_pdl = []() { const char* e = getenv("EVENTFUL_B200_PDL"); return (e && e[0] == '1') ? 1 : 0; }();

unsigned long long* g_gate_dbg = nullptr;  // et_debug_set(3, device pointer to 8 x u64) enables phase timestamps

namespace {

constexpr int kGateThreads = 256;
constexpr int kSmemKeys = 8192;

struct GateArgs {
    const void* xa;
    const void* xb;
    void* xsum;
    void* c_out;  // optional: the gate input c = LN(x) of every token (gather source of et_linear_gather)
    const void* ln_w;
    const void* ln_b;
    float eps;
    const void* p;
    int N, D, mode, k;
    float thr;
    float* norm;
    long long* idx;
    int* count;
    int* ticket;
    int tokens_per_cta;
    int low_bit;  // lowest significant bit of the fp32 key for this dtype
    unsigned long long* dbg;  // optional phase timestamps (et_debug_set key 3)
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Order-preserving key of a float (torch's radix-select convention: NaN sorts largest).
__device__ __forceinline__ uint32_t order_key(float v) {
    uint32_t x = __float_as_uint(v);
    if (v != v) return 0xffffffffu;
    return (x & 0x80000000u) ? ~x : (x | 0x80000000u);
}

// Loads one token row (CPL 16-byte chunks per lane), optionally adds a second row, and
// optionally LayerNorms it.  `v` receives the row as stored in dtype T (already rounded).
template <typename T, int LPT, int CPL>
__device__ __forceinline__ void load_row(const T* xa, const T* xb, T* xsum, size_t off, int nchunks, int lane,
                                         bool valid, float (&v)[CPL * ElemTraits<T>::VEC]) {
    constexpr int VEC = ElemTraits<T>::VEC;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int ch = lane + c * LPT;
        const bool on = valid && ch < nchunks;
        if (on) {
            uint4 u = ld_stream16(xa + off + (size_t)ch * VEC);
            unpack16<T>(u, &v[c * VEC]);
            if (xb != nullptr) {
                float w[VEC];
                uint4 ub = ld_stream16(xb + off + (size_t)ch * VEC);
                unpack16<T>(ub, w);
#pragma unroll
                for (int i = 0; i < VEC; i += 2) {
                    v[c * VEC + i] += w[i];
                    v[c * VEC + i + 1] += w[i + 1];
                    round2_to<T>(v[c * VEC + i], v[c * VEC + i + 1]);
                }
                if (xsum != nullptr) st16(xsum + off + (size_t)ch * VEC, pack16<T>(&v[c * VEC]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[c * VEC + i] = 0.f;
        }
    }
}

// In-register LayerNorm over one row spread across LPT lanes (two-pass, fp32), result rounded to T.
// w / b may point to shared memory (gate_select stages them once per CTA) or global memory (gate_gather).
template <typename T, int LPT, int CPL>
__device__ __forceinline__ void layer_norm_row(float (&v)[CPL * ElemTraits<T>::VEC], const T* w, const T* b,
                                               int nchunks, int lane, int D, float eps) {
    constexpr int VEC = ElemTraits<T>::VEC;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPL * VEC; ++i) s += v[i];
    const float mean = group_sum<LPT>(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        if (lane + c * LPT < nchunks) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float d = v[c * VEC + i] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(group_sum<LPT>(q) / (float)D + eps);
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int ch = lane + c * LPT;
        if (ch < nchunks) {
            float g[VEC], o[VEC];
            unpack16<T>(ld16(w + (size_t)ch * VEC), g);
            unpack16<T>(ld16(b + (size_t)ch * VEC), o);
#pragma unroll
            for (int i = 0; i < VEC; i += 2) {
                v[c * VEC + i] = (v[c * VEC + i] - mean) * rstd * g[i] + o[i];
                v[c * VEC + i + 1] = (v[c * VEC + i + 1] - mean) * rstd * g[i + 1] + o[i + 1];
                round2_to<T>(v[c * VEC + i], v[c * VEC + i + 1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Selection stage, executed by one CTA per row after all norms of the row are visible.
// ------------------------------------------------------------------------------------------
template <int KPT>  // register-resident key slots per thread (N <= KPT * 256)
__device__ __forceinline__ void select_row(const GateArgs& a, int r, int* s_hist, uint32_t* s_keys) {
    const int tid = threadIdx.x;
    const int N = a.N;
    const float* norm = a.norm + (size_t)r * N;
    long long* out = a.idx + (size_t)r * (a.mode == ET_SELECT_TOPK ? a.k : N);
    const int per = (N + kGateThreads - 1) / kGateThreads;
    const int lo = min(N, tid * per), hi = min(N, lo + per);
    const int lane = tid & 31, warp = tid >> 5;

    if (a.mode == ET_SELECT_THRESHOLD) {
        // strict `>` against the threshold rounded to the tensor dtype (policies.py:28), ascending
        int c = 0;
        for (int i = lo; i < hi; ++i) c += (__ldcg(norm + i) > a.thr) ? 1 : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_hist[warp] = inc;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; ++w) base += s_hist[w];
        int pos = base + inc - c;
        for (int i = lo; i < hi; ++i)
            if (__ldcg(norm + i) > a.thr) out[pos++] = i;
        if (tid == kGateThreads - 1) a.count[r] = pos;
        return;
    }

    if (a.k == 0) return;
    // ---- k-th largest key by radix search on 2-bit digits, no atomics: thread t keeps the keys of tokens
    // [t * per, (t + 1) * per) in registers (coalesced 128-bit loads staged through shared memory); per step every
    // thread tallies its candidate keys into one 64-bit word of four 16-bit counters, two REDUX sum a warp and one
    // barrier sums the CTA.  Steps = ceil(significant key bits / 2): 8 for bf16, 10 for fp16, 16 for fp32 norms.
    uint32_t keys[KPT];
    const bool in_regs = per <= KPT;
    if (in_regs) {
        if ((N & 3) == 0) {
            const float4* n4 = reinterpret_cast<const float4*>(norm);
            for (int i = tid; i < N / 4; i += kGateThreads) {
                const float4 f = __ldcg(n4 + i);
                *reinterpret_cast<uint4*>(s_keys + 4 * i) = make_uint4(order_key(f.x), order_key(f.y), order_key(f.z), order_key(f.w));
            }
        } else {
            for (int i = tid; i < N; i += kGateThreads) s_keys[i] = order_key(__ldcg(norm + i));
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KPT; ++j) keys[j] = (lo + j < hi) ? s_keys[lo + j] : 0u;
    }
    if (a.dbg != nullptr && tid == 0) a.dbg[3] = gtime();
    uint32_t kth = 0, known = 0;  // `known` = mask of the key bits decided so far
    int remaining = a.k;
    int slot = 0;
    for (int shift = 30; shift + 1 >= a.low_bit; shift -= 2, slot ^= 1) {
        unsigned long long cnt = 0;  // four 16-bit counters: digit d in bits [16 d, 16 d + 16)
        if (in_regs) {
#pragma unroll
            for (int j = 0; j < KPT; ++j)
                if (lo + j < hi && (keys[j] & known) == kth) cnt += 1ull << (((keys[j] >> shift) & 3u) * 16);
        } else {
            for (int i = lo; i < hi; ++i) {
                const uint32_t key = order_key(__ldcg(norm + i));
                if ((key & known) == kth) cnt += 1ull << (((key >> shift) & 3u) * 16);
            }
        }
        const uint32_t lo32 = __reduce_add_sync(0xffffffffu, (uint32_t)cnt);
        const uint32_t hi32 = __reduce_add_sync(0xffffffffu, (uint32_t)(cnt >> 32));
        if (lane == 0) {
            s_hist[slot * 16 + warp * 2] = (int)lo32;
            s_hist[slot * 16 + warp * 2 + 1] = (int)hi32;
        }
        __syncthreads();
        uint32_t t01 = 0, t23 = 0;
#pragma unroll
        for (int q = 0; q < kGateThreads / 32; ++q) {
            t01 += (uint32_t)s_hist[slot * 16 + q * 2];
            t23 += (uint32_t)s_hist[slot * 16 + q * 2 + 1];
        }
        const int c3 = (int)(t23 >> 16), c2 = (int)(t23 & 0xffffu), c1 = (int)(t01 >> 16);
        int digit, above;  // the digit whose suffix count (from the top) reaches `remaining`
        if (c3 >= remaining) { digit = 3; above = 0; }
        else if (c3 + c2 >= remaining) { digit = 2; above = c3; }
        else if (c3 + c2 + c1 >= remaining) { digit = 1; above = c3 + c2; }
        else { digit = 0; above = c3 + c2 + c1; }
        kth |= (uint32_t)digit << shift;
        known |= 3u << shift;
        remaining -= above;
    }
    kth &= (a.low_bit == 0 ? 0xffffffffu : ~((1u << a.low_bit) - 1u));
    if (a.dbg != nullptr && tid == 0) a.dbg[4] = gtime();
    // kth = k-th largest key (low insignificant bits zero). Strictly greater keys first (ascending index), then
    // keys equal to it (ascending index) until k are written: torch's CUDA radix-select order.
    const uint32_t mask = a.low_bit == 0 ? 0xffffffffu : ~((1u << a.low_bit) - 1u);
    int cg = 0, ce = 0;
    if (in_regs) {
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const bool on = lo + j < hi;
            cg += on && (keys[j] & mask) > kth;
            ce += on && (keys[j] & mask) == kth;
        }
    } else {
        for (int i = lo; i < hi; ++i) {
            const uint32_t key = order_key(__ldcg(norm + i)) & mask;
            cg += key > kth;
            ce += key == kth;
        }
    }
    int ig = cg, ie = ce;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int tg = __shfl_up_sync(0xffffffffu, ig, o);
        int te = __shfl_up_sync(0xffffffffu, ie, o);
        if (lane >= o) {
            ig += tg;
            ie += te;
        }
    }
    if (a.dbg != nullptr && tid == 0) a.dbg[6] = gtime();
    __syncthreads();
    if (lane == 31) {
        s_hist[64 + warp] = ig;
        s_hist[96 + warp] = ie;
    }
    __syncthreads();
    if (a.dbg != nullptr && tid == 0) a.dbg[7] = gtime();
    int bg = 0, be = 0, n_greater = 0;
#pragma unroll
    for (int w = 0; w < kGateThreads / 32; ++w) {
        if (w < warp) {
            bg += s_hist[64 + w];
            be += s_hist[96 + w];
        }
        n_greater += s_hist[64 + w];
    }
    remaining = a.k - n_greater;  // ties taken (equals the search's residual count)
    int pg = bg + ig - cg, pe = be + ie - ce;
    // branch-free: one predicated store per key (divergent if / else chains cost ~3 us here)
    auto emit = [&](int i, uint32_t key, bool on) {
        const bool g = on && key > kth, e = on && key == kth;
        const int slot_i = g ? pg : n_greater + pe;
        const bool take = g || (e && pe < remaining);
        pg += g ? 1 : 0;
        pe += e ? 1 : 0;
        if (take) out[slot_i] = i;
    };
    if (in_regs) {  // static indices only: the key array must stay in registers
#pragma unroll
        for (int j = 0; j < KPT; ++j) emit(lo + j, keys[j] & mask, lo + j < hi);
    } else {
        for (int i = lo; i < hi; ++i) emit(i, order_key(__ldcg(norm + i)) & mask, true);
    }
}

template <typename T, int LPT, int CPL>
__global__ void __launch_bounds__(kGateThreads) gate_select_kernel(const GateArgs a) {
    et_pdl_prologue();
    constexpr int VEC = ElemTraits<T>::VEC;
    constexpr int GROUPS = kGateThreads / LPT;
    __shared__ int s_hist[128];
    __shared__ uint32_t s_keys[kSmemKeys];
    __shared__ int s_misc[4];

    if (a.dbg != nullptr && threadIdx.x == 0) atomicMin(a.dbg + 0, gtime());
    const int r = blockIdx.y;
    const int lane = threadIdx.x % LPT;
    const int group = threadIdx.x / LPT;
    const int nchunks = a.D / VEC;
    const int t0 = blockIdx.x * a.tokens_per_cta;
    const int t1 = min(a.N, t0 + a.tokens_per_cta);
    const T* xa = static_cast<const T*>(a.xa);
    const T* xb = static_cast<const T*>(a.xb);
    T* xsum = static_cast<T*>(a.xsum);
    const T* p = static_cast<const T*>(a.p);
    const int iters = (a.tokens_per_cta + GROUPS - 1) / GROUPS;
    const T* ln_w = static_cast<const T*>(a.ln_w);
    const T* ln_b = static_cast<const T*>(a.ln_b);
    const bool stage_ln = ln_w != nullptr && (size_t)a.D * sizeof(T) * 2 <= sizeof(s_keys);

    // Raw 16-byte chunks of one token (input, residual, reference state).  Two tokens per warp are kept in flight:
    // all their loads (and the LayerNorm-parameter staging) are issued before the first reduction starts, so the
    // phase is one DRAM round trip instead of a chain of three.
    struct Raw {
        uint4 xa[CPL], xb[CPL], p[CPL];
    };
    auto issue = [&](int it, Raw& rw) {
        const int tok = t0 + it * GROUPS + group;
        const bool valid = it < iters && tok < t1;
        const size_t off = ((size_t)r * a.N + (valid ? tok : 0)) * a.D;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int ch = lane + c * LPT;
            const bool on = valid && ch < nchunks;
            rw.xa[c] = on ? ld_stream16(xa + off + (size_t)ch * VEC) : make_uint4(0, 0, 0, 0);
            rw.xb[c] = (on && xb != nullptr) ? ld_stream16(xb + off + (size_t)ch * VEC) : make_uint4(0, 0, 0, 0);
            rw.p[c] = (on && p != nullptr) ? ld_stream16(p + off + (size_t)ch * VEC) : make_uint4(0, 0, 0, 0);
        }
    };
    auto process = [&](int it, const Raw& rw) {
        const int tok = t0 + it * GROUPS + group;
        const bool valid = it < iters && tok < t1;
        const size_t off = ((size_t)r * a.N + (valid ? tok : 0)) * a.D;
        float v[CPL * VEC];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int ch = lane + c * LPT;
            unpack16<T>(rw.xa[c], &v[c * VEC]);
            if (xb != nullptr) {
                float w[VEC];
                unpack16<T>(rw.xb[c], w);
#pragma unroll
                for (int i = 0; i < VEC; i += 2) {
                    v[c * VEC + i] += w[i];
                    v[c * VEC + i + 1] += w[i + 1];
                    round2_to<T>(v[c * VEC + i], v[c * VEC + i + 1]);
                }
                if (xsum != nullptr && valid && ch < nchunks) st16(xsum + off + (size_t)ch * VEC, pack16<T>(&v[c * VEC]));
            }
        }
        if (ln_w != nullptr) layer_norm_row<T, LPT, CPL>(v, ln_w, ln_b, nchunks, lane, a.D, a.eps);
        if (a.c_out != nullptr && valid) {
            T* c_out = static_cast<T*>(a.c_out);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int ch = lane + c * LPT;
                if (ch < nchunks) st16(c_out + off + (size_t)ch * VEC, pack16<T>(&v[c * VEC]));
            }
        }
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            if (lane + c * LPT < nchunks) {
                float q[VEC];
                if (p != nullptr) unpack16<T>(rw.p[c], q);
#pragma unroll
                for (int i = 0; i < VEC; i += 2) {
                    float e0 = v[c * VEC + i], e1 = v[c * VEC + i + 1];
                    if (p != nullptr) {
                        e0 -= q[i];
                        e1 -= q[i + 1];
                        round2_to<T>(e0, e1);
                    }
                    ss = fmaf(e0, e0, ss);
                    ss = fmaf(e1, e1, ss);
                }
            }
        }
        ss = group_sum<LPT>(ss);
        if (valid && lane == 0) a.norm[(size_t)r * a.N + tok] = round_to<T>(sqrtf(ss));
    };

    Raw ra, rb;
    issue(0, ra);
    issue(1, rb);
    if (stage_ln) {  // one copy of the LayerNorm affine parameters per CTA (overlaid on the key buffer of the selection stage)
        T* sw = reinterpret_cast<T*>(s_keys);
        T* sb = sw + a.D;
        for (int c = threadIdx.x; c < nchunks; c += kGateThreads) {
            st16(sw + (size_t)c * VEC, ld16(ln_w + (size_t)c * VEC));
            st16(sb + (size_t)c * VEC, ld16(ln_b + (size_t)c * VEC));
        }
        __syncthreads();
        ln_w = sw;
        ln_b = sb;
    }
    for (int it = 0; it < iters; it += 2) {
        process(it, ra);
        if (it + 2 < iters) issue(it + 2, ra);
        process(it + 1, rb);
        if (it + 3 < iters) issue(it + 3, rb);
    }

    if (a.dbg != nullptr && threadIdx.x == 0) atomicMax(a.dbg + 1, gtime());
    // ---- last CTA of this row runs the selection (threadFenceReduction pattern)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(a.ticket + r, 1);
        s_misc[2] = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_misc[2]) return;
    __threadfence();
    if (a.dbg != nullptr && threadIdx.x == 0) a.dbg[2] = gtime();
    if (a.N <= 16 * kGateThreads) select_row<16>(a, r, s_hist, s_keys);
    else select_row<32>(a, r, s_hist, s_keys);
    if (a.dbg != nullptr && threadIdx.x == 0) a.dbg[5] = gtime();
    if (threadIdx.x == 0) a.ticket[r] = 0;  // self-reset so the workspace is reusable / graph-replayable
}

// ------------------------------------------------------------------------------------------
struct GatherArgs {
    const void* x;
    const void* ln_w;
    const void* ln_b;
    float eps;
    int ln_after;
    void* p;
    const long long* idx;
    const int* count;
    int N, D, k;
    void* c_tilde;
    void* e_tilde;
    int total_rows;
};

template <typename T, int LPT, int CPL>
__global__ void __launch_bounds__(kGateThreads) gate_gather_kernel(const GatherArgs a) {
    et_pdl_prologue();
    constexpr int VEC = ElemTraits<T>::VEC;
    constexpr int GROUPS = kGateThreads / LPT;
    const int lane = threadIdx.x % LPT;
    const int group = threadIdx.x / LPT;
    const int nchunks = a.D / VEC;
    const int row = blockIdx.x * GROUPS + group;  // selected row m = r * k + j
    const bool in_range = row < a.total_rows;
    const int r = in_range ? row / a.k : 0;
    const int j = in_range ? row % a.k : 0;
    const bool valid = in_range && (a.count == nullptr || j < a.count[r]);
    long long tok = valid ? a.idx[(size_t)r * a.k + j] : 0;
    const size_t src = ((size_t)r * a.N + (size_t)tok) * a.D;
    const size_t dst = (size_t)row * a.D;
    const T* x = static_cast<const T*>(a.x);
    T* p = static_cast<T*>(a.p);
    float v[CPL * VEC];
    load_row<T, LPT, CPL>(x, nullptr, nullptr, src, nchunks, lane, valid, v);
    const bool ln = a.ln_w != nullptr;
    if (ln && !a.ln_after)
        layer_norm_row<T, LPT, CPL>(v, static_cast<const T*>(a.ln_w), static_cast<const T*>(a.ln_b), nchunks, lane,
                                    a.D, a.eps);
    // v now holds the gate input rows c[idx]
    if (valid) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int ch = lane + c * LPT;
            if (ch < nchunks) {
                const uint4 packed = pack16<T>(&v[c * VEC]);
                if (a.e_tilde != nullptr) {
                    float q[VEC], e[VEC];
                    unpack16<T>(ld16(p + src + (size_t)ch * VEC), q);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) e[i] = v[c * VEC + i] - q[i];
                    st16(static_cast<T*>(a.e_tilde) + dst + (size_t)ch * VEC, pack16<T>(e));
                }
                if (p != nullptr) st16(p + src + (size_t)ch * VEC, packed);
                if (!(ln && a.ln_after)) st16(static_cast<T*>(a.c_tilde) + dst + (size_t)ch * VEC, packed);
            }
        }
    }
    if (ln && a.ln_after) {
        layer_norm_row<T, LPT, CPL>(v, static_cast<const T*>(a.ln_w), static_cast<const T*>(a.ln_b), nchunks, lane,
                                    a.D, a.eps);
        if (valid) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int ch = lane + c * LPT;
                if (ch < nchunks) st16(static_cast<T*>(a.c_tilde) + dst + (size_t)ch * VEC, pack16<T>(&v[c * VEC]));
            }
        }
    }
}

// p <- LN?(x) for every token (SimpleSTGTGate, modules.py:44) / plain vector copy
template <typename T, int LPT, int CPL>
__global__ void __launch_bounds__(kGateThreads) gate_replace_kernel(const GatherArgs a) {
    et_pdl_prologue();
    constexpr int VEC = ElemTraits<T>::VEC;
    constexpr int GROUPS = kGateThreads / LPT;
    const int lane = threadIdx.x % LPT;
    const int group = threadIdx.x / LPT;
    const int nchunks = a.D / VEC;
    const int row = blockIdx.x * GROUPS + group;
    const bool valid = row < a.total_rows;
    const size_t off = (size_t)(valid ? row : 0) * a.D;
    float v[CPL * VEC];
    load_row<T, LPT, CPL>(static_cast<const T*>(a.x), nullptr, nullptr, off, nchunks, lane, valid, v);
    if (a.ln_w != nullptr && !a.ln_after)
        layer_norm_row<T, LPT, CPL>(v, static_cast<const T*>(a.ln_w), static_cast<const T*>(a.ln_b), nchunks, lane,
                                    a.D, a.eps);
    if (valid) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int ch = lane + c * LPT;
            if (ch < nchunks) st16(static_cast<T*>(a.p) + off + (size_t)ch * VEC, pack16<T>(&v[c * VEC]));
        }
    }
}

// ------------------------------------------------------------------------------------------
// TokenBuffer scatter: rows (16-byte vectors) or columns (element granularity).
template <typename T>
__global__ void __launch_bounds__(256) scatter_rows_kernel(T* buf, const T* x, const long long* idx,
                                                           const int* count, int rows_per_index, int N, int D,
                                                           int k, long long total_vec) {
    et_pdl_prologue();
    constexpr int VEC = ElemTraits<T>::VEC;
    const int nchunks = D / VEC;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total_vec;
         g += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(g % nchunks);
        const long long row = g / nchunks;  // r * k + j
        const long long r = row / k;
        const int j = (int)(row % k);
        const long long ir = r / rows_per_index;
        if (count != nullptr && j >= count[ir]) continue;
        const long long tok = idx[ir * k + j];
        st16(buf + ((size_t)r * N + (size_t)tok) * D + (size_t)ch * VEC, ld_stream16(x + (size_t)row * D + (size_t)ch * VEC));
    }
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_cols_kernel(T* buf, const T* x, const long long* idx,
                                                           int rows_per_index, int N, int M, int k,
                                                           long long total) {
    et_pdl_prologue();
    // buf (R, N, M), x (R, N, k): buf[r, n, idx[j]] = x[r, n, j]
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(g % k);
        const long long rn = g / k;
        const long long r = rn / N;
        const long long col = idx[(r / rows_per_index) * k + j];
        buf[(size_t)rn * M + (size_t)col] = x[g];
    }
}

// column gate: c~ = c[..., idx], e~ = c~ - p[..., idx], p[..., idx] = c~  on (R, N, M)
template <typename T>
__global__ void __launch_bounds__(256) gather_cols_kernel(const T* c, T* p, const long long* idx,
                                                          int rows_per_index, int N, int M, int k, T* c_tilde,
                                                          T* e_tilde, long long total) {
    et_pdl_prologue();
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(g % k);
        const long long rn = g / k;
        const long long r = rn / N;
        const long long col = idx[(r / rows_per_index) * k + j];
        const size_t at = (size_t)rn * M + (size_t)col;
        const T cv = c[at];
        if (e_tilde != nullptr)
            e_tilde[g] = ElemTraits<T>::from_float(ElemTraits<T>::to_float(cv) - ElemTraits<T>::to_float(p[at]));
        if (p != nullptr) p[at] = cv;
        c_tilde[g] = cv;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* a, const T* b, T* out, long long nvec, float sign) {
    et_pdl_prologue();
    constexpr int VEC = ElemTraits<T>::VEC;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < nvec;
         g += (long long)gridDim.x * blockDim.x) {
        float x[VEC], y[VEC];
        unpack16<T>(ld_stream16(a + g * VEC), x);
        unpack16<T>(ld_stream16(b + g * VEC), y);
#pragma unroll
        for (int i = 0; i < VEC; ++i) x[i] = fmaf(sign, y[i], x[i]);
        st16(out + g * VEC, pack16<T>(x));
    }
}

// ---------------------------------------------------------------- row-shape dispatch
// A token row of D elements is spread over LPT lanes x CPL 16-byte chunks.
template <typename T, template <typename, int, int> class Launcher, typename... Args>
int dispatch_row_shape(int D, Args&&... args) {
    constexpr int VEC = ElemTraits<T>::VEC;
    const int nchunks = D / VEC;
    if (nchunks <= 1) return Launcher<T, 1, 1>::run(args...);
    if (nchunks <= 2) return Launcher<T, 2, 1>::run(args...);
    if (nchunks <= 4) return Launcher<T, 4, 1>::run(args...);
    if (nchunks <= 8) return Launcher<T, 8, 1>::run(args...);
    if (nchunks <= 16) return Launcher<T, 16, 1>::run(args...);
    if (nchunks <= 32) return Launcher<T, 32, 1>::run(args...);
    if (nchunks <= 64) return Launcher<T, 32, 2>::run(args...);
    if (nchunks <= 96) return Launcher<T, 32, 3>::run(args...);
    if (nchunks <= 128) return Launcher<T, 32, 4>::run(args...);
    if (nchunks <= 192) return Launcher<T, 32, 6>::run(args...);
    if (nchunks <= 256) return Launcher<T, 32, 8>::run(args...);
    if constexpr (sizeof(T) == 4) {
        if (nchunks <= 384) return Launcher<T, 32, 12>::run(args...);
        if (nchunks <= 512) return Launcher<T, 32, 16>::run(args...);
    }
    return et_fail(ET_ERR_UNSUPPORTED, "token width D=%d not supported by the gate kernels (max 2048)", D);
}

template <typename T, int LPT, int CPL>
struct SelectLauncher {
    static int run(const GateArgs& a, int ctas_per_row, int R, cudaStream_t s) {
        et_launch(gate_select_kernel<T, LPT, CPL>, dim3(dim3(ctas_per_row, R)), dim3(kGateThreads), 0, s, a);
        ET_COUNT_LAUNCH(1);
        return 0;
    }
};
template <typename T, int LPT, int CPL>
struct GatherLauncher {
    static int run(const GatherArgs& a, cudaStream_t s) {
        constexpr int GROUPS = kGateThreads / LPT;
        et_launch(gate_gather_kernel<T, LPT, CPL>, dim3((a.total_rows + GROUPS - 1) / GROUPS), dim3(kGateThreads), 0, s, a);
        ET_COUNT_LAUNCH(1);
        return 0;
    }
};
template <typename T, int LPT, int CPL>
struct ReplaceLauncher {
    static int run(const GatherArgs& a, cudaStream_t s) {
        constexpr int GROUPS = kGateThreads / LPT;
        et_launch(gate_replace_kernel<T, LPT, CPL>, dim3((a.total_rows + GROUPS - 1) / GROUPS), dim3(kGateThreads), 0, s, a);
        ET_COUNT_LAUNCH(1);
        return 0;
    }
};

int grid_for(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)et_sm_count() * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" {

int et_version(void) { return 200; }

long long et_launch_count(void) { return g_et_launches; }

const char* et_last_error(void) { return g_et_error; }

int et_device_info(int device, int* cc_major, int* cc_minor, int* sm_count) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return et_fail(ET_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (prop.major != 10)
        return et_fail(ET_ERR_UNSUPPORTED, "device %d is sm_%d%d; libeventful_b200 is built for sm_100a only", device,
                       prop.major, prop.minor);
    return ET_OK;
}

int et_gate_select(const void* xa, const void* xb, void* xsum_out, void* c_out, const void* ln_w, const void* ln_b, float ln_eps,
                   const void* p, int64_t R, int64_t N, int64_t D, int dtype, int mode, int64_t k, float threshold,
                   float* norm_out, int64_t* idx_out, int32_t* count_out, int32_t* ticket, void* stream) {
    ET_CHECK_ARG(xa && norm_out && idx_out && ticket, "et_gate_select: null pointer");
    ET_CHECK_ARG(R > 0 && N > 0 && D > 0 && R <= 65535 && N <= 524280, "et_gate_select: bad shape R=%lld N=%lld D=%lld",
                 (long long)R, (long long)N, (long long)D);
    ET_CHECK_ARG((ln_w == nullptr) == (ln_b == nullptr), "et_gate_select: ln_w / ln_b must both be set or both null");
    ET_CHECK_ARG(et_aligned16(xa) && et_aligned16(xb) && et_aligned16(xsum_out) && et_aligned16(c_out) && et_aligned16(p) &&
                     et_aligned16(ln_w) && et_aligned16(ln_b),
                 "et_gate_select: pointers must be 16-byte aligned");
    if (mode == ET_SELECT_TOPK) {
        // torch.topk raises for k > N (policies.py:63)
        ET_CHECK_ARG(k >= 0 && k <= N, "et_gate_select: selected index k out of range (k=%lld, N=%lld)", (long long)k,
                     (long long)N);
    } else {
        ET_CHECK_ARG(mode == ET_SELECT_THRESHOLD && count_out, "et_gate_select: bad mode / missing count_out");
    }
    GateArgs a;
    a.xa = xa; a.xb = xb; a.xsum = xsum_out; a.c_out = c_out; a.ln_w = ln_w; a.ln_b = ln_b; a.eps = ln_eps; a.p = p;
    a.N = (int)N; a.D = (int)D; a.mode = mode; a.k = (int)k; a.norm = norm_out;
    a.idx = reinterpret_cast<long long*>(idx_out); a.count = count_out; a.ticket = ticket;
    a.dbg = g_gate_dbg;
    int rc = ET_OK;
    ET_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = ElemTraits<T>::VEC;
        ET_CHECK_ARG(D % VEC == 0, "et_gate_select: D=%lld must be a multiple of %d", (long long)D, VEC);
        const int nchunks = (int)D / VEC;
        const int lpt = nchunks <= 1 ? 1 : nchunks <= 2 ? 2 : nchunks <= 4 ? 4 : nchunks <= 8 ? 8 : nchunks <= 16 ? 16 : 32;
        const int groups = kGateThreads / lpt;
        // ~4 CTAs per SM over the whole launch (all resident: more loads in flight), each CTA at least one pass
        long long want = ((long long)g_gate_cta_waves * et_sm_count() + R - 1) / R;
        long long max_ctas = (N + groups - 1) / groups;
        long long ctas = want < 1 ? 1 : (want > max_ctas ? max_ctas : want);
        a.tokens_per_cta = (int)((N + ctas - 1) / ctas);
        ctas = (N + a.tokens_per_cta - 1) / a.tokens_per_cta;
        a.low_bit = sizeof(T) == 4 ? 0 : (dtype == ET_BF16 ? 16 : 13);
        a.thr = (dtype == ET_F32) ? threshold
                                  : (dtype == ET_BF16 ? __bfloat162float(__float2bfloat16_rn(threshold))
                                                      : __half2float(__float2half_rn(threshold)));
        rc = dispatch_row_shape<T, SelectLauncher>((int)D, a, (int)ctas, (int)R, et_stream(stream));
    });
    if (rc) return rc;
    ET_CHECK_LAUNCH("et_gate_select");
    return ET_OK;
}

int et_gate_gather(const void* x, const void* ln_w, const void* ln_b, float ln_eps, int ln_after, void* p,
                   const int64_t* idx, const int32_t* count, int64_t R, int64_t N, int64_t D, int64_t k, int dtype,
                   void* c_tilde, void* e_tilde, int full_replace, void* stream) {
    ET_CHECK_ARG(x && idx && c_tilde, "et_gate_gather: null pointer");
    ET_CHECK_ARG(R > 0 && N > 0 && D > 0 && k >= 0 && R * k < (1LL << 31), "et_gate_gather: bad shape");
    ET_CHECK_ARG(e_tilde == nullptr || p != nullptr, "et_gate_gather: e_tilde needs the reference state p");
    ET_CHECK_ARG(et_aligned16(x) && et_aligned16(p) && et_aligned16(c_tilde) && et_aligned16(e_tilde) &&
                     et_aligned16(ln_w) && et_aligned16(ln_b),
                 "et_gate_gather: pointers must be 16-byte aligned");
    GatherArgs a;
    a.x = x; a.ln_w = ln_w; a.ln_b = ln_b; a.eps = ln_eps; a.ln_after = ln_after; a.p = p;
    a.idx = reinterpret_cast<const long long*>(idx); a.count = count; a.N = (int)N; a.D = (int)D; a.k = (int)k;
    a.c_tilde = c_tilde; a.e_tilde = e_tilde; a.total_rows = (int)(R * k);
    int rc = ET_OK;
    ET_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = ElemTraits<T>::VEC;
        ET_CHECK_ARG(D % VEC == 0, "et_gate_gather: D=%lld must be a multiple of %d", (long long)D, VEC);
        if (a.total_rows > 0) rc = dispatch_row_shape<T, GatherLauncher>((int)D, a, et_stream(stream));
        if (!rc && full_replace) {
            ET_CHECK_ARG(p != nullptr, "et_gate_gather: full_replace needs p");
            GatherArgs b = a;
            b.total_rows = (int)(R * N);
            rc = dispatch_row_shape<T, ReplaceLauncher>((int)D, b, et_stream(stream));
        }
    });
    if (rc) return rc;
    ET_CHECK_LAUNCH("et_gate_gather");
    return ET_OK;
}

int et_gate_gather_cols(const void* c, void* p, const int64_t* idx, int64_t R, int64_t rows_per_index, int64_t N,
                        int64_t M, int64_t k, int dtype, void* c_tilde, void* e_tilde, void* stream) {
    ET_CHECK_ARG(c && idx && c_tilde && rows_per_index > 0, "et_gate_gather_cols: null pointer");
    ET_CHECK_ARG(e_tilde == nullptr || p != nullptr, "et_gate_gather_cols: e_tilde needs p");
    const long long total = (long long)R * N * k;
    if (total == 0) return ET_OK;
    ET_DISPATCH_DTYPE(dtype, T, {
        et_launch(gather_cols_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, et_stream(stream), static_cast<const T*>(c), static_cast<T*>(p), reinterpret_cast<const long long*>(idx), (int)rows_per_index,
            (int)N, (int)M, (int)k, static_cast<T*>(c_tilde), static_cast<T*>(e_tilde), total);
    });
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_gate_gather_cols");
    return ET_OK;
}

int et_buffer_scatter(void* buf, const void* x, const int64_t* idx, const int32_t* count, int64_t R,
                      int64_t rows_per_index, int64_t N, int64_t D, int64_t k, int dtype, int structure,
                      void* stream) {
    ET_CHECK_ARG(buf && x && idx && rows_per_index > 0, "et_buffer_scatter: null pointer");
    ET_CHECK_ARG(structure == 0 || structure == 1, "et_buffer_scatter: structure must be 0 (row) or 1 (col)");
    if (R * k == 0) return ET_OK;
    ET_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = ElemTraits<T>::VEC;
        if (structure == 0) {
            ET_CHECK_ARG(D % VEC == 0 && et_aligned16(buf) && et_aligned16(x),
                         "et_buffer_scatter: rows must be 16-byte multiples and aligned");
            const long long total = (long long)R * k * (D / VEC);
            et_launch(scatter_rows_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, et_stream(stream), static_cast<T*>(buf), static_cast<const T*>(x), reinterpret_cast<const long long*>(idx), count,
                (int)rows_per_index, (int)N, (int)D, (int)k, total);
        } else {
            ET_CHECK_ARG(count == nullptr, "et_buffer_scatter: device-side counts are row-structure only");
            const long long total = (long long)R * N * k;  // buf (R, N, D): D is the indexed axis
            et_launch(scatter_cols_kernel<T>, dim3(grid_for(total, 256)), dim3(256), 0, et_stream(stream), static_cast<T*>(buf), static_cast<const T*>(x), reinterpret_cast<const long long*>(idx),
                (int)rows_per_index, (int)N, (int)D, (int)k, total);
        }
    });
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_buffer_scatter");
    return ET_OK;
}

static int add_impl(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream, float sign);

int et_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
    return add_impl(a, b, out, n, dtype, stream, 1.f);
}

int et_sub(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
    return add_impl(a, b, out, n, dtype, stream, -1.f);
}

static int add_impl(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream, float sign) {
    ET_CHECK_ARG(a && b && out && n >= 0, "et_add: null pointer");
    ET_CHECK_ARG(et_aligned16(a) && et_aligned16(b) && et_aligned16(out), "et_add: pointers must be 16-byte aligned");
    if (n == 0) return ET_OK;
    ET_DISPATCH_DTYPE(dtype, T, {
        constexpr int VEC = ElemTraits<T>::VEC;
        ET_CHECK_ARG(n % VEC == 0, "et_add: element count must be a multiple of %d", VEC);
        et_launch(add_kernel<T>, dim3(grid_for(n / VEC, 256)), dim3(256), 0, et_stream(stream), static_cast<const T*>(a), static_cast<const T*>(b), static_cast<T*>(out), n / VEC, sign);
    });
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_add");
    return ET_OK;
}

}  // extern "C"
