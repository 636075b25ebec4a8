// MUFU.EX2 issue rate per SM for the fp32, f16x2 and bf16x2 forms (sm_100a): is a packed exp2 two results per MUFU slot?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu ; run: ./mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) rate_kernel(uint32_t* out, long long* cycles, int iters, uint32_t seed) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
            if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
            if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int results_per_op) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    rate_kernel<MODE><<<148, 1024>>>(out, cyc, iters, 0x3c003c00u);
    rate_kernel<MODE><<<148, 1024>>>(out, cyc, iters, 0x3c003c00u);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double ops = 1024.0 * 8 * iters;
    printf("%-24s %s  %.2f thread-ops/clk/SM  = %.2f results/clk/SM\n", name, cudaGetErrorString(e), ops / avg, results_per_op * ops / avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.ftz.bf16x2", 2);
    run<2>("ex2.approx.f16x2", 2);
    return 0;
}
