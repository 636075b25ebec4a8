// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, cta_group::1) on resident smem operands.
// Varies N, the number of interleaved independent accumulators and the operand majors.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../eventful-transformer_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include "et_tcgen05.cuh"
thread_local char g_et_error[512];
long long g_et_launches = 0;
using namespace et_tc;

// A operand read from tensor memory instead of shared memory (column address a_tmem; 8 columns per K = 16 step)
__device__ __forceinline__ void mma_f16_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.u32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int n, int n_acc, int a_mn, int b_mn, int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = umma_idesc_ex(128, n, 1, b_mn);
        if (a_mn == 1) idesc |= 1u << 15;
        uint64_t da = a_mn == 1 ? umma_smem_desc_mn_a(smem_u32(smem)) : umma_smem_desc(smem_u32(smem));
        uint64_t db = umma_smem_desc(smem_u32(smem + 32768));
        const int astep = a_mn == 1 ? 128 : 2, bstep = b_mn ? 128 : 2;
        // warm-up
        for (int i = 0; i < 8; ++i) tcgen05_mma_f16(tmem, da, db, idesc, i > 0);
        tcgen05_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t0 = clock64();
        if (a_mn == 2) {  // A from TMEM columns 448..479, accumulators below
            for (int r = 0; r < reps; ++r) {
#pragma unroll 4
                for (int kk = 0; kk < 4; ++kk)
                    for (int acc = 0; acc < n_acc; ++acc)
                        mma_f16_tmem_a(tmem + acc * (n > 128 ? 256 : 128), tmem + 448 + 8 * kk, db + (uint64_t)(bstep * kk), idesc, 1u);
            }
        } else
        for (int r = 0; r < reps; ++r) {
#pragma unroll 4
            for (int kk = 0; kk < 4; ++kk)
                for (int acc = 0; acc < n_acc; ++acc)
                    tcgen05_mma_f16(tmem + acc * (512 / 4), da + (uint64_t)(astep * kk), db + (uint64_t)(bstep * kk), idesc, 1u);
        }
        const long long t1 = clock64();
        tcgen05_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 1);
        const long long t2 = clock64();
        out[0] = t1 - t0;  // issue time
        out[1] = t2 - t0;  // completion time
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 64;
    printf("%5s %5s %4s %4s | %12s %12s  (cycles per MMA: issue, complete; floor = N/2)\n", "N", "n_acc", "a_mn", "b_mn", "issue", "complete");
    for (int grid : {1, 148})
        for (int a_mn = 0; a_mn < 3; ++a_mn)   // 0 = smem K-major, 1 = smem MN-major, 2 = tensor memory
            for (int b_mn = 0; b_mn < 2; ++b_mn) {
                if (a_mn != b_mn && grid == 148 && a_mn != 2) continue;
                for (int n : {64, 128, 256})
                    for (int n_acc : {1, 2, 4}) {
                        if (n * n_acc > 512 || (n > 128 && b_mn)) continue;
                        if (a_mn == 2 && (n_acc > 2 || (n > 128 && n_acc > 1))) continue;
                        mma_rate_kernel<<<grid, 64, 100 * 1024>>>(n, n_acc, a_mn, b_mn, reps, d);
                        long long h[2];
                        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                        const double m = (double)reps * 4 * n_acc;
                        printf("%5d %5d %4d %4d | %12.1f %12.1f   grid=%d\n", n, n_acc, a_mn, b_mn, h[0] / m, h[1] / m, grid);
                    }
            }
    return 0;
}
