// Generic strided batched matmul for the stand-alone MatmulBuffer / MatmulDeltaAccumulator modules
// (modules.py:204-299 used outside the fused block path).  Plain shared-memory tiling on CUDA cores,
// fp32 accumulate; the fused attention kernels in et_attn.cu are the performance path.
#include "et_common.cuh"

namespace {

constexpr int TILE = 16;

struct BmmArgs {
    const void* A;
    const void* B;
    void* C;
    long long M, N, K;
    long long sab, sam, sak, sbb, sbk, sbn, scb, scm, scn;
    int accumulate;
};

template <typename T>
__global__ void __launch_bounds__(TILE * TILE) bmm_kernel(const BmmArgs a) {
    et_pdl_prologue();
    __shared__ float As[TILE][TILE + 1];
    __shared__ float Bs[TILE][TILE + 1];
    const T* A = static_cast<const T*>(a.A) + (long long)blockIdx.z * a.sab;
    const T* B = static_cast<const T*>(a.B) + (long long)blockIdx.z * a.sbb;
    T* C = static_cast<T*>(a.C) + (long long)blockIdx.z * a.scb;
    const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
    const long long m = (long long)blockIdx.y * TILE + ty, n = (long long)blockIdx.x * TILE + tx;
    float acc = 0.f;
    for (long long k0 = 0; k0 < a.K; k0 += TILE) {
        As[ty][tx] = (m < a.M && k0 + tx < a.K) ? ElemTraits<T>::to_float(A[m * a.sam + (k0 + tx) * a.sak]) : 0.f;
        Bs[ty][tx] = (k0 + ty < a.K && n < a.N) ? ElemTraits<T>::to_float(B[(k0 + ty) * a.sbk + n * a.sbn]) : 0.f;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TILE; ++kk) acc = fmaf(As[ty][kk], Bs[kk][tx], acc);
        __syncthreads();
    }
    if (m < a.M && n < a.N) {
        T* c = C + m * a.scm + n * a.scn;
        // `product += matmul(...)`: the matmul result is rounded to dtype before the in-place add (modules.py:293)
        float v = round_to<T>(acc);
        if (a.accumulate) v += ElemTraits<T>::to_float(*c);
        *c = ElemTraits<T>::from_float(v);
    }
}

}  // namespace

extern "C" int et_bmm(const void* A, const void* Bm, void* C, int64_t batch, int64_t M, int64_t N, int64_t K, int64_t sab,
                      int64_t sam, int64_t sak, int64_t sbb, int64_t sbk, int64_t sbn, int64_t scb, int64_t scm,
                      int64_t scn, int accumulate, int dtype, void* stream) {
    ET_CHECK_ARG(A && Bm && C, "et_bmm: null pointer");
    ET_CHECK_ARG(batch >= 0 && M >= 0 && N >= 0 && K >= 0 && batch <= 65535, "et_bmm: bad shape");
    if (batch == 0 || M == 0 || N == 0) return ET_OK;
    BmmArgs a;
    a.A = A; a.B = Bm; a.C = C; a.M = M; a.N = N; a.K = K;
    a.sab = sab; a.sam = sam; a.sak = sak; a.sbb = sbb; a.sbk = sbk; a.sbn = sbn; a.scb = scb; a.scm = scm; a.scn = scn;
    a.accumulate = accumulate;
    const dim3 grid((unsigned)((N + TILE - 1) / TILE), (unsigned)((M + TILE - 1) / TILE), (unsigned)batch);
    ET_CHECK_ARG(grid.y <= 65535, "et_bmm: M too large");
    ET_DISPATCH_DTYPE(dtype, T, { et_launch(bmm_kernel<T>, dim3(grid), dim3(TILE * TILE), 0, et_stream(stream), a); });
    ET_COUNT_LAUNCH(1);
    ET_CHECK_LAUNCH("et_bmm");
    return ET_OK;
}
