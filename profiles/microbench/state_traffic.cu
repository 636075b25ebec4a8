// Micro-benchmark: the A-gate state traffic of tc_apply_kernel WITHOUT any compute.
// State (H, N, NP) bf16, column-major per head (a selected column = NP contiguous rows).  Grid (N / rows, H) CTAs of 256
// threads; per 64-column tile each thread moves 4 x 16 B: read a_state[:, idx] (cp.async into a 3-deep smem ring, like the
// kernel) and write the same bytes back.  Answers: is ~240 us the floor of this access pattern, or is the kernel slow?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 state_traffic.cu -o state_traffic
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

template <int ROWS, int MODE>  // MODE 0 = read + write, 1 = read only, 2 = write only
__global__ void __launch_bounds__(256) traffic_kernel(uint16_t* state, const long long* idx, int k, int N, int NP, int smem_pad) {
    extern __shared__ uint4 ring[];  // [3][64 * ROWS / 8]
    constexpr int SEGS = ROWS / 8;           // 16-byte segments per column
    constexpr int CPT = 64 * SEGS / 256;     // chunks per thread per tile
    constexpr int CSTEP = 256 / SEGS;        // columns covered by one pass of the 256 threads
    const int q0 = blockIdx.x * ROWS, h = blockIdx.y;
    const int st = threadIdx.x, seg = (st % SEGS) * 8, col0 = st / SEGS;
    uint16_t* base = state + (size_t)h * N * NP;
    const int T = k / 64;
    uint4 acc = make_uint4(0, 0, 0, 0);
    auto load = [&](int t) {
        uint4* dst = ring + (t % 3) * (64 * SEGS);
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = col0 + CSTEP * i;
            const long long tok = idx[t * 64 + c];
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + c * SEGS + (st % SEGS));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(base + (size_t)tok * NP + q0 + seg) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (MODE != 2) for (int t = 0; t < 3 && t < T; ++t) load(t);
    for (int t = 0; t < T; ++t) {
        uint4 v[CPT];
        if (MODE != 2) {
            if (t + 2 < T) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            const uint4* src = ring + (t % 3) * (64 * SEGS);
#pragma unroll
            for (int i = 0; i < CPT; ++i) v[i] = src[(col0 + CSTEP * i) * SEGS + (st % SEGS)];
            __syncthreads();
            if (t + 3 < T) load(t + 3);
        } else {
#pragma unroll
            for (int i = 0; i < CPT; ++i) v[i] = make_uint4(t, i, st, 0);
        }
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const long long tok = idx[t * 64 + col0 + CSTEP * i];
                *reinterpret_cast<uint4*>(base + (size_t)tok * NP + q0 + seg) = v[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < CPT; ++i) acc.x ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
        }
    }
    if (MODE == 1 && acc.x == 0x12345678u) state[0] = 1;
}

template <int ROWS, int MODE>
float run(uint16_t* state, const long long* idx, int k, int N, int H, int ctas_per_sm) {
    const int smem = 3 * 64 * ROWS * 2;
    // pad dynamic smem so that only `ctas_per_sm` CTAs fit on an SM
    const int want = ctas_per_sm == 1 ? 120 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : smem);
    const int dyn = std::max(smem, want);
    cudaFuncSetAttribute(traffic_kernel<ROWS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        traffic_kernel<ROWS, MODE><<<dim3(N / ROWS, H), 256, dyn>>>(state, idx, k, N, N, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
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
    for (int order = 0; order < 2; ++order) {
        std::shuffle(all.begin(), all.end(), rng);
        sel.assign(all.begin(), all.begin() + k);
        if (order == 0) std::sort(sel.begin(), sel.end());
        cudaMemcpy(idx, sel.data(), k * 8, cudaMemcpyHostToDevice);
        const double mb = 2.0 * H * N * k * 2 / 1e6;  // read + write
        printf("index order %s; state traffic %.0f MB (read + write)\n", order == 0 ? "ascending" : "random", mb);
        for (int cps : {1, 2, 4}) {
            const float rw = run<128, 0>(state, idx, k, N, H, cps), r = run<128, 1>(state, idx, k, N, H, cps), w = run<128, 2>(state, idx, k, N, H, cps);
            printf("  rows/CTA 128 (256 B segments), %d CTA/SM: read+write %6.1f us (%5.0f GB/s)  read only %6.1f us  write only %6.1f us\n",
                   cps, rw, mb / rw * 1e3, r, w);
            const float rw2 = run<256, 0>(state, idx, k, N, H, cps), r2 = run<256, 1>(state, idx, k, N, H, cps), w2 = run<256, 2>(state, idx, k, N, H, cps);
            printf("  rows/CTA 256 (512 B segments), %d CTA/SM: read+write %6.1f us (%5.0f GB/s)  read only %6.1f us  write only %6.1f us\n",
                   cps, rw2, mb / rw2 * 1e3, r2, w2);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
