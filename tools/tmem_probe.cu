// tools/tmem_probe.cu -- does tensor memory work as a per-thread parking space?  Every warp of a 128-thread CTA stores NV
// 32-bit values per thread into its own 32-lane quarter of the CTA's TMEM allocation (tcgen05.st.32x32b), reads them back
// (tcgen05.ld.32x32b) in a timed loop, and checks them.  Prints cycles per LDTM of 16 columns and the verdict.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                    "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}

constexpr int kCols = 128;

__global__ void __launch_bounds__(128, 4) probe(int* bad, long long* cycles, int reps) {
    __shared__ uint32_t tm_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tm_base)), "n"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tm_base + ((uint32_t)(warp * 32) << 16);   // this warp's lane quarter
    uint32_t v[16];
    for (int c = 0; c < kCols; c += 16) {
        for (int i = 0; i < 16; ++i) v[i] = (blockIdx.x << 20) ^ (threadIdx.x << 8) ^ (c + i);
        tm_st16(base + c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    int nbad = 0;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
        for (int c = 0; c < kCols; c += 16) {
            tm_ld16(base + c, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            for (int i = 0; i < 16; ++i) {
                acc += v[i];
                if (r == 0 && v[i] != ((blockIdx.x << 20) ^ (threadIdx.x << 8) ^ (uint32_t)(c + i))) ++nbad;
            }
        }
    const long long t1 = clock64();
    if (nbad) atomicAdd(bad, nbad);
    if (acc == 0x12345678u) atomicAdd(bad, 1);
    if (blockIdx.x == 0 && lane == 0) cycles[warp] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm_base), "n"(kCols));
}

int main() {
    int* bad; long long* cyc;
    cudaMalloc(&bad, 4); cudaMalloc(&cyc, 32); cudaMemset(bad, 0, 4);
    const int reps = 200;
    probe<<<148 * 4, 128>>>(bad, cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    int hb = -1; long long hc[4] = {0, 0, 0, 0};
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost);
    printf("{\"cuda\": \"%s\", \"mismatches\": %d, \"cycles_per_ldtm_x16_with_wait\": [%.1f, %.1f, %.1f, %.1f], \"ctas_per_sm\": 4, \"warps_per_sm\": 16}\n",
           cudaGetErrorString(e), hb, hc[0] / (double)(reps * kCols / 16), hc[1] / (double)(reps * kCols / 16),
           hc[2] / (double)(reps * kCols / 16), hc[3] / (double)(reps * kCols / 16));
    return e != cudaSuccess || hb != 0;
}
