// pipe_microbench.cu -- measures the sm_100a issue rates that bound the ADMM kernel (K2):
// scalar FFMA, packed FFMA2 (fma.rn.f32x2), FADD/FMUL, FMNMX, SHFL, REDUX, LDS, and their latencies.
// MEASURED_PEAKS.json has no CUDA-core figure; this supplies the "FP32 pipe" roofline denominator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipe_microbench tools/pipe_microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kIters = 4096;

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// MODE: 0 FFMA, 1 FFMA2, 2 FADD, 3 FMUL, 4 FMNMX, 5 SHFL.up, 6 REDUX.max, 7 LDS, 8 FFMA+SHFL 3:1, 9 FFMA2+SHFL 3:2,
//       10 FADD2, 11 FFMA + FMNMX 1:1, 12 MUFU.RSQ, 13 FFMA2 + FMNMX 1:1, 14 SHFL.bfly
template <int MODE, int ILP>
__global__ void bench_kernel(float* out, long long* cyc, float seed) {
    __shared__ float sh[1024];
    float a[ILP], b = seed, c = seed * 0.5f;
    unsigned long long p[ILP];
    const unsigned long long pb = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed * 0.25f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed + i + threadIdx.x;
        p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f);
    }
    sh[threadIdx.x & 1023] = seed;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);
            if (MODE == 1) p[i] = fma2(p[i], pb, pb);
            if (MODE == 2) a[i] = a[i] + b;
            if (MODE == 3) a[i] = a[i] * b;
            if (MODE == 4) a[i] = fmaxf(a[i], b + i);
            if (MODE == 5) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
            if (MODE == 6) a[i] = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(a[i])) + i);
            if (MODE == 7) a[i] = sh[(__float_as_uint(a[i]) + threadIdx.x) & 1023];
            if (MODE == 8) { a[i] = (i & 3) == 3 ? __shfl_up_sync(0xffffffffu, a[i], 1) : fmaf(a[i], b, c); }
            if (MODE == 9) { if ((i % 5) >= 3) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1); else p[i] = fma2(p[i], pb, pb); }
            if (MODE == 10) p[i] = add2(p[i], pb);
            if (MODE == 11) { a[i] = (i & 1) ? fmaxf(a[i], b + i) : fmaf(a[i], b, c); }
            if (MODE == 12) a[i] = rsqrtf(a[i]);
            if (MODE == 13) { if (i & 1) a[i] = fmaxf(a[i], b + i); else p[i] = fma2(p[i], pb, pb); }
            if (MODE == 14) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 16);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + __uint_as_float((unsigned)(p[i] & 0xffffffffu)) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int ILP>
static void run(const char* name, int warps_per_sm, float* out, long long* cyc, int nsm) {
    const int threads = 32 * warps_per_sm;
    bench_kernel<MODE, ILP><<<nsm, threads>>>(out, cyc, 1.0001f);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    bench_kernel<MODE, ILP><<<nsm, threads>>>(out, cyc, 1.0001f);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < nsm; ++i) mean += (double)h[i];
    mean /= nsm;
    const double instr = (double)kIters * ILP * warps_per_sm;  // warp-instructions per SM
    printf("{\"op\": \"%s\", \"ilp\": %d, \"warps_per_sm\": %d, \"cycles\": %.0f, \"warp_instr_per_clk_per_sm\": %.3f, "
           "\"cycles_per_instr_per_warp\": %.3f, \"ms\": %.4f}\n",
           name, ILP, warps_per_sm, mean, instr / mean, mean / ((double)kIters * ILP), ms);
}

int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", pr.name, nsm, pr.clockRate);
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * nsm * 1024));
    CK(cudaMalloc(&cyc, sizeof(long long) * 256));
    const int ws[] = {1, 4, 8, 16, 32};
    for (int w : ws) {
        run<0, 8>("FFMA", w, out, cyc, nsm);
        run<1, 8>("FFMA2", w, out, cyc, nsm);
        run<2, 8>("FADD", w, out, cyc, nsm);
        run<3, 8>("FMUL", w, out, cyc, nsm);
        run<10, 8>("FADD2", w, out, cyc, nsm);
        run<4, 8>("FMNMX", w, out, cyc, nsm);
        run<5, 8>("SHFL.up", w, out, cyc, nsm);
        run<14, 8>("SHFL.bfly", w, out, cyc, nsm);
        run<6, 8>("REDUX.max", w, out, cyc, nsm);
        run<7, 8>("LDS", w, out, cyc, nsm);
        run<12, 8>("MUFU.RSQ", w, out, cyc, nsm);
        run<8, 8>("FFMA+SHFL 3:1", w, out, cyc, nsm);
        run<9, 10>("FFMA2+SHFL 3:2", w, out, cyc, nsm);
        run<11, 8>("FFMA+FMNMX 1:1", w, out, cyc, nsm);
        run<13, 8>("FFMA2+FMNMX 1:1", w, out, cyc, nsm);
    }
    // latencies: one warp, one dependent chain
    run<0, 1>("lat FFMA", 1, out, cyc, nsm);
    run<1, 1>("lat FFMA2", 1, out, cyc, nsm);
    run<2, 1>("lat FADD", 1, out, cyc, nsm);
    run<4, 1>("lat FMNMX", 1, out, cyc, nsm);
    run<5, 1>("lat SHFL.up", 1, out, cyc, nsm);
    run<6, 1>("lat REDUX.max", 1, out, cyc, nsm);
    run<7, 1>("lat LDS", 1, out, cyc, nsm);
    run<12, 1>("lat MUFU.RSQ", 1, out, cyc, nsm);
    // two and four chains per warp (what 2 warps/SMSP of a dependent solver see)
    run<0, 2>("FFMA ilp2", 8, out, cyc, nsm);
    run<0, 4>("FFMA ilp4", 8, out, cyc, nsm);
    run<1, 2>("FFMA2 ilp2", 8, out, cyc, nsm);
    run<1, 4>("FFMA2 ilp4", 8, out, cyc, nsm);
    return 0;
}
