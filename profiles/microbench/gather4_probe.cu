// Probe of the TMA tile::gather4 load on sm_100a: which box shape the tensor map needs and how the four gathered rows land in
// shared memory under the 128-byte swizzle (the layout the tcgen05 K-major operand descriptors expect).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 gather4_probe.cu -o gather4_probe -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, const int* rows, uint16_t* out, int col0, int groups) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(groups * 4 * 128) : "memory");
        for (int g = 0; g < groups; ++g)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                :
                : "r"(smem_u32(smem + g * 512)), "l"(&tm), "r"(smem_u32(&bar)), "r"(col0), "r"(rows[4 * g]), "r"(rows[4 * g + 1]),
                  "r"(rows[4 * g + 2]), "r"(rows[4 * g + 3])
                : "memory");
    }
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("timeout: the transaction count never completed\n"); return; }
    }
    for (int i = threadIdx.x; i < groups * 4 * 64; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
    const int R = 256, C = 128;
    std::vector<uint16_t> h(R * C);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 256 + c);  // raw 16-bit tags
    uint16_t *d, *out;
    cudaMalloc(&d, h.size() * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    const int groups = 8;
    std::vector<int> rows = {5, 9, 2, 60, 100, 101, 102, 103, 7, 7, 255, 0, 31, 30, 29, 28, 200, 1, 150, 3, 64, 65, 66, 67, 11, 12, 13, 14, 250, 251, 252, 253};
    int* drows;
    cudaMalloc(&drows, rows.size() * 4);
    cudaMemcpy(drows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, groups * 4 * 64 * 2);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Fn encode = (Fn)p;
    for (int box_rows : {1, 4}) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box rows %d: encode -> %d\n", box_rows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        cudaMemset(out, 0xff, groups * 4 * 64 * 2);
        probe<<<1, 128, groups * 512 + 1024>>>(tm, drows, out, 64, groups);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint16_t> o(groups * 4 * 64);
        cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
        // expected: smem row j (128 B) holds source row rows[j], columns 64..127, 16-byte chunk c stored at chunk c ^ (j & 7)
        int bad = 0;
        for (int j = 0; j < groups * 4; ++j)
            for (int c = 0; c < 8; ++c)
                for (int e2 = 0; e2 < 8; ++e2) {
                    const uint16_t want = (uint16_t)(rows[j] * 256 + 64 + c * 8 + e2);
                    const uint16_t got = o[j * 64 + ((c ^ (j & 7)) * 8) + e2];
                    if (want != got && bad++ < 4) printf("  mismatch row %d chunk %d: want %04x got %04x\n", j, c, want, got);
                }
        printf("  %s (%d mismatches): smem row j = source row idx[j], 128B-swizzled by (j & 7)\n", bad ? "DIFFERENT LAYOUT" : "layout as expected", bad);
    }
    return 0;
}
