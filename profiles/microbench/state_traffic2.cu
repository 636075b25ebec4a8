// Micro-benchmark 2: which thread structure moves the A-gate state tiles fastest?  Same traffic as tc_apply_kernel
// (per 64-column tile and CTA: 64 x 256 B read + 64 x 256 B written, columns picked by idx), one CTA per SM, no compute.
//   variant 0: T threads, each does reads (cp.async ring, 4 deep) and writes        (T = 128 / 256 / 512)
//   variant 1: T/2 reader threads + T/2 writer threads (separate warps)
//   variant 2: like 0 with st.global.cs (evict-first) stores
//   variant 3: writes as 256-byte cp.async.bulk (smem -> global), one per column, reads as in 0
//   variant 4: reads AND writes as 256-byte bulk copies (mbarrier complete_tx / bulk_group)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 state_traffic2.cu -o state_traffic2
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

constexpr int ROWS = 128, SEGS = 16, TILE = 64 * 256, STAGES = 4;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

template <int VARIANT>
__global__ void __launch_bounds__(512) traffic2(uint16_t* state, const long long* idx, int k, int N, int NP, int nthreads) {
    extern __shared__ __align__(128) uint8_t ring[];  // [STAGES][TILE]
    __shared__ uint64_t full[STAGES], done[STAGES];
    const int q0 = blockIdx.x * ROWS, h = blockIdx.y, tid = threadIdx.x;
    uint16_t* base = state + (size_t)h * N * NP;
    const int T = k / 64;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(VARIANT == 4 ? 1 : (VARIANT == 1 ? nthreads / 2 : nthreads)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&done[s])), "r"(VARIANT == 1 ? nthreads / 2 : nthreads));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nread = VARIANT == 1 ? nthreads / 2 : nthreads;   // threads taking part in reads
    const int nwrite = nread;
    const bool is_reader = VARIANT == 1 ? tid < nread : true;
    const bool is_writer = VARIANT == 1 ? tid >= nread : true;
    const int rt = tid, wt = VARIANT == 1 ? tid - nread : tid;

    auto issue_read = [&](int t) {
        uint8_t* dst = ring + (t % STAGES) * TILE;
        if (VARIANT == 4) {
            if (rt == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[t % STAGES])), "r"(TILE) : "memory");
            __syncwarp();
            for (int c = rt; c < 64; c += nread) {
                const long long tok = idx[t * 64 + c];
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst + c * 256)),
                             "l"(base + (size_t)tok * NP + q0), "r"(256), "r"(s32(&full[t % STAGES])) : "memory");
            }
        } else {
            for (int ch = rt; ch < 64 * SEGS; ch += nread) {
                const int c = ch / SEGS, sg = ch % SEGS;
                const long long tok = idx[t * 64 + c];
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst + c * 256 + sg * 16)), "l"(base + (size_t)tok * NP + q0 + sg * 8) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(&full[t % STAGES])) : "memory");
        }
    };
    if (is_reader) {
        if (VARIANT == 4) { if (rt < 64) for (int t = 0; t < STAGES && t < T; ++t) issue_read(t); }
        else for (int t = 0; t < STAGES && t < T; ++t) issue_read(t);
    }
    for (int t = 0; t < T; ++t) {
        const int s = t % STAGES;
        const uint32_t ph = (t / STAGES) & 1;
        if (is_writer) {
            mbar_wait(s32(&full[s]), ph);
            const uint8_t* src = ring + s * TILE;
            if (VARIANT == 3 || VARIANT == 4) {
                for (int c = wt; c < 64; c += nwrite) {
                    const long long tok = idx[t * 64 + c];
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)tok * NP + q0), "r"(s32(src + c * 256)), "r"(256) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem source may be overwritten
            } else {
                for (int ch = wt; ch < 64 * SEGS; ch += nwrite) {
                    const int c = ch / SEGS, sg = ch % SEGS;
                    const long long tok = idx[t * 64 + c];
                    const uint4 v = *reinterpret_cast<const uint4*>(src + c * 256 + sg * 16);
                    uint16_t* dstp = base + (size_t)tok * NP + q0 + sg * 8;
                    if (VARIANT == 2) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dstp), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    else *reinterpret_cast<uint4*>(dstp) = v;
                }
            }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&done[s])) : "memory");
        }
        if (is_reader && t + STAGES < T) {
            if (VARIANT == 4 && rt >= 64) continue;
            mbar_wait(s32(&done[s]), ph);   // every writer has copied tile t out of slot s
            issue_read(t + STAGES);
        }
    }
    if (VARIANT == 3 || VARIANT == 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int VARIANT>
float run(uint16_t* state, const long long* idx, int k, int N, int H, int nthreads) {
    const int dyn = 120 * 1024;  // one CTA per SM
    cudaFuncSetAttribute(traffic2<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        traffic2<VARIANT><<<dim3(N / ROWS, H), nthreads, dyn>>>(state, idx, k, N, N, nthreads);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return best * 1e3f;
}

int main() {
    const int N = 4096, H = 12, k = 2048;
    uint16_t* state; long long* idx;
    cudaMalloc(&state, (size_t)H * N * N * 2);
    cudaMemset(state, 0, (size_t)H * N * N * 2);
    cudaMalloc(&idx, k * 8);
    std::vector<long long> all(N), sel;
    for (int i = 0; i < N; ++i) all[i] = i;
    std::mt19937 rng(1);
    std::shuffle(all.begin(), all.end(), rng);
    sel.assign(all.begin(), all.begin() + k);
    std::sort(sel.begin(), sel.end());
    cudaMemcpy(idx, sel.data(), k * 8, cudaMemcpyHostToDevice);
    const double mb = 2.0 * H * N * k * 2 / 1e6;
    printf("state traffic %.0f MB (read + write), ascending index order, 1 CTA/SM, 4-deep ring\n", mb);
    for (int nt : {128, 256, 512}) {
        const float a = run<0>(state, idx, k, N, H, nt), b = run<1>(state, idx, k, N, H, nt), c = run<2>(state, idx, k, N, H, nt),
                    d = run<3>(state, idx, k, N, H, nt), e = run<4>(state, idx, k, N, H, nt);
        printf("  %3d threads: r+w same threads %6.1f us (%4.0f GB/s) | split reader/writer warps %6.1f | st.cs %6.1f | bulk stores %6.1f | bulk loads+stores %6.1f\n",
               nt, a, mb / a * 1e3, b, c, d, e);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
