// Generic strided batched matmul C = alpha * A @ B (+ C) for the stand-alone MatmulBuffer / MatmulDeltaAccumulator /
// CountedMatmul modules (modules.py:204-299, counting.py:165-175 used outside the fused block path; operands are
// arbitrary strided views such as the q / k^T views of a QKV buffer).  Register-tiled on the CUDA cores: 64 x 64 output
// tile per CTA, 4 x 4 per thread, fp32 accumulate; the fused attention kernels are the performance path.
#include "et_common.cuh"

namespace {

constexpr int BT = 64, BKK = 16, BTHREADS = 256, BLD = BT + 4;

struct BmmArgs {
    const void* A;
    const void* B;
    void* C;
    long long M, N, K;
    long long sab, sai, sam, sak, sbb, sbi, sbk, sbn, scb, sci, scm, scn;  // outer-batch, inner-batch, row, column strides
    int inner;
    int accumulate;
    float alpha;
};

template <typename T>
__global__ void __launch_bounds__(BTHREADS) bmm_kernel(const BmmArgs a) {
    et_pdl_prologue();
    __shared__ __align__(16) float As[BKK][BLD];  // As[k][m]
    __shared__ __align__(16) float Bs[BKK][BLD];  // Bs[k][n]
    const long long bo = blockIdx.z / a.inner, bi = blockIdx.z - bo * a.inner;
    const T* A = static_cast<const T*>(a.A) + bo * a.sab + bi * a.sai;
    const T* B = static_cast<const T*>(a.B) + bo * a.sbb + bi * a.sbi;
    T* C = static_cast<T*>(a.C) + bo * a.scb + bi * a.sci;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long m0 = (long long)blockIdx.y * BT, n0 = (long long)blockIdx.x * BT;
    // loader index order follows the unit-stride axis of each operand so that global reads coalesce
    const bool a_k_fast = a.sak == 1, b_n_fast = a.sbn == 1;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (long long k0 = 0; k0 < a.K; k0 += BKK) {
        for (int i = tid; i < BT * BKK; i += BTHREADS) {
            const int m = a_k_fast ? i / BKK : i % BT, kk = a_k_fast ? i % BKK : i / BT;
            As[kk][m] = (m0 + m < a.M && k0 + kk < a.K) ? ElemTraits<T>::to_float(A[(m0 + m) * a.sam + (k0 + kk) * a.sak]) : 0.f;
            const int n = b_n_fast ? i % BT : i / BKK, kb = b_n_fast ? i / BT : i % BKK;
            Bs[kb][n] = (n0 + n < a.N && k0 + kb < a.K) ? ElemTraits<T>::to_float(B[(k0 + kb) * a.sbk + (n0 + n) * a.sbn]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BKK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            T* c = C + m * a.scm + n * a.scn;
            // `product += matmul(...)`: the matmul result is rounded to dtype before the in-place add (modules.py:293)
            float v = round_to<T>(acc[i][j] * a.alpha);
            if (a.accumulate) v += ElemTraits<T>::to_float(*c);
            *c = ElemTraits<T>::from_float(v);
        }
    }
}

}  // namespace

extern "C" int et_bmm(const void* A, const void* Bm, void* C, int64_t batch_outer, int64_t batch_inner, int64_t M, int64_t N,
                      int64_t K, const int64_t* strides_a, const int64_t* strides_b, const int64_t* strides_c, int accumulate,
                      float alpha, int dtype, void* stream) {
    ET_CHECK_ARG(A && Bm && C && strides_a && strides_b && strides_c, "et_bmm: null pointer");
    const int64_t batch = batch_outer * batch_inner;
    ET_CHECK_ARG(batch_outer >= 0 && batch_inner >= 0 && M >= 0 && N >= 0 && K >= 0 && batch <= 65535, "et_bmm: bad shape");
    if (batch == 0 || M == 0 || N == 0) return ET_OK;
    BmmArgs a;
    a.A = A; a.B = Bm; a.C = C; a.M = M; a.N = N; a.K = K; a.inner = (int)batch_inner;
    a.sab = strides_a[0]; a.sai = strides_a[1]; a.sam = strides_a[2]; a.sak = strides_a[3];
    a.sbb = strides_b[0]; a.sbi = strides_b[1]; a.sbk = strides_b[2]; a.sbn = strides_b[3];
    a.scb = strides_c[0]; a.sci = strides_c[1]; a.scm = strides_c[2]; a.scn = strides_c[3];
    a.accumulate = accumulate; a.alpha = alpha;
    const dim3 grid((unsigned)((N + BT - 1) / BT), (unsigned)((M + BT - 1) / BT), (unsigned)batch);
    ET_CHECK_ARG(grid.y <= 65535, "et_bmm: M too large");
    ET_DISPATCH_DTYPE(dtype, T, { et_launch(bmm_kernel<T>, dim3(grid), dim3(BTHREADS), 0, et_stream(stream), a); });
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_bmm");
    return ET_OK;
}

// ---------------------------------------------------------------- patch / tubelet extraction for the embedding GEMM
// LinearEmbedding (models/vitdet.py:17-52: Conv2d with kernel = stride = patch) and TubeletEmbedding (models/vivit.py:
// 153-192: Conv3d with kernel = stride = tubelet) are GEMMs over non-overlapping patches.  This kernel lays the patches out
// as rows -- out[b, t', n, (c, dt, dy, dx)] = x[b, t' pt + dt, c, y ph + dy, x pw + dx], the order of the flattened conv
// weight (dim, C, pt, ph, pw) -- so that the projection itself runs on et_linear.
namespace {
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const T* x, T* out, int Tn, int C, int H, int W, int pt, int ph, int pw,
                                                       long long total) {
    et_pdl_prologue();
    const int gw = W / pw, gh = H / ph, tn = Tn / pt;
    const int F = C * pt * ph * pw;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(gi % F);
        long long r = gi / F;
        const int n = (int)(r % (gh * gw));
        r /= gh * gw;
        const int tt = (int)(r % tn);
        const long long b = r / tn;
        const int dx = f % pw, dy = (f / pw) % ph, dt = (f / (pw * ph)) % pt, c = f / (pw * ph * pt);
        const int py = n / gw, px = n - py * gw;
        out[gi] = x[(((b * Tn + (long long)tt * pt + dt) * C + c) * H + py * ph + dy) * W + px * pw + dx];
    }
}
}  // namespace

extern "C" int et_patchify(const void* x, void* out, int64_t B, int64_t T, int64_t C, int64_t H, int64_t W, int64_t pt, int64_t ph,
                           int64_t pw, int dtype, void* stream) {
    ET_CHECK_ARG(x && out, "et_patchify: null pointer");
    ET_CHECK_ARG(pt > 0 && ph > 0 && pw > 0 && T % pt == 0 && H % ph == 0 && W % pw == 0,
                 "et_patchify: input (%lld, %lld, %lld) is not a multiple of the patch (%lld, %lld, %lld)", (long long)T, (long long)H,
                 (long long)W, (long long)pt, (long long)ph, (long long)pw);
    const long long total = (long long)B * T * C * H * W;
    if (total == 0) return ET_OK;
    const long long blocks = (total + 255) / 256, cap = (long long)et_sm_count() * 16;
    const dim3 grid((unsigned)(blocks > cap ? cap : blocks));
    ET_DISPATCH_DTYPE(dtype, Tp, {
        et_launch(patchify_kernel<Tp>, grid, dim3(256), 0, et_stream(stream), static_cast<const Tp*>(x), static_cast<Tp*>(out), (int)T, (int)C,
                  (int)H, (int)W, (int)pt, (int)ph, (int)pw, total);
    });
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_patchify");
    return ET_OK;
}
