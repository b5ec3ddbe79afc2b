// admm_tm.cu -- kernel entry points of the tensor-memory variant of the paired-stage fp32 path (see admm_tm.cuh).
#include "launch_util.h"
#include "engine.h"
#include "admm_epilogue.cuh"
#include "admm_tm.cuh"
#include <cstdio>
#include <cstdlib>

namespace mpcb {

// from admm_pair.cu (same epilogue: MPC.get_control after the solve, MPC.py:185-222)
template <int LPS>
__device__ __forceinline__ void control_epilogue_tm(const GroupComm<LPS>& cm, const MpcParams& mp, const f2 w[5], const SolveResult& r,
                                                    double* cc, int* infeas, double* u_out, int* iters, int* qp_status, int* flags,
                                                    int b, int fl, const RolloutArgs& ro) {
    const int N = mp.N, gl = cm.gl, kA = 2 * gl, kB = kA + 1;
    const bool ok = !(r.status == -3 || r.status == -4 || r.status == -7 || r.status == 3 || r.status == 4);
    int inf = infeas[b];
    if (ok) {
        if (kA < N) { cc[2 * kA] = (double)w[3].x; cc[2 * kA + 1] = atan((double)w[4].x * mp.L); }  // MPC.py:187-189
        if (kB < N) { cc[2 * kB] = (double)w[3].y; cc[2 * kB + 1] = atan((double)w[4].y * mp.L); }
        if (gl == 0) {
            u_out[2 * (size_t)b] = (double)w[3].x;
            u_out[2 * (size_t)b + 1] = atan((double)w[4].x * mp.L);
            inf = 0;  // MPC.py:206
            fl &= ~MPC_ST_QP_FALLBACK;
        }
    } else if (gl == 0) {
        const int id = 2 * (inf + 1);  // MPC.py:212-213
        u_out[2 * (size_t)b] = cc[id];
        u_out[2 * (size_t)b + 1] = cc[id + 1];
        inf += 1;  // MPC.py:216
        fl |= MPC_ST_QP_FALLBACK;
    }
    if (gl == 0) {
        if (inf == N - 1) fl |= MPC_ST_DEAD;  // MPC.py:218-220
        if (ro.state && !(fl & MPC_ST_DEAD))
            drive_one(ro.state, b, ro.B, ro.spatial[b], ro.spatial[(size_t)ro.B + b], ro.kappa[ro.wp],
                      u_out[2 * (size_t)b], u_out[2 * (size_t)b + 1], mp.L, ro.Ts);
        infeas[b] = inf;
        if (flags) flags[b] = fl;
        if (iters) iters[b] = r.iters;
        if (qp_status) qp_status[b] = r.status;
        store_host_results(ro, u_out, b, fl);
    }
}

constexpr int kTmAllocCols = 128;  // = kTmCols, a power of two >= 32
constexpr int kTmWarps = 4;      // warps per CTA: warp w owns TMEM lanes 32 w .. 32 w + 31 of the CTA's columns
constexpr int kTmMinBlocks = 4;  // 4 CTAs x 128 columns = the SM's 512 columns; 16 warps x 128 registers = its register file

template <int LPS>
__global__ void __launch_bounds__(32 * kTmWarps, kTmMinBlocks)
assemble_solve_tm_kernel(MpcParams mp, AdmmSettings st, const f2 al2, const f2 nal2, PathView pv, const double* __restrict__ spatial,
                         const int* __restrict__ wp_id, double* __restrict__ control, const double* __restrict__ ub,
                         const double* __restrict__ lb, int* __restrict__ infeas, double* __restrict__ u_out,
                         double* __restrict__ x_out, int* __restrict__ iters, int* __restrict__ qp_status, int* __restrict__ flags,
                         int B, double* __restrict__ rollout_state, double Ts, const int* __restrict__ order,
                         f2* __restrict__ cold_base) {
    constexpr int G = 32 / LPS;
    __shared__ uint32_t tm_base_s;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tm_base_s)), "n"(kTmAllocCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const TmStore tm{tm_base_s + ((uint32_t)(warp * 32) << 16)};
    const int gwarp = blockIdx.x * kTmWarps + warp;
    const int slot = gwarp * G + lane / LPS;
    int b = slot < B ? (order ? order[slot] : slot) : B;
    if ((unsigned)b >= (unsigned)B) b = B;
    const int fl = (b < B && flags) ? flags[b] : 0;
    const bool live = b < B && !(fl & (MPC_ST_DEAD | MPC_ST_FINISHED));
    if (__any_sync(kFull, live)) {
        const GroupComm<LPS> cm;
        const int N = mp.N, n = 5 * N + 3;
        double* cc = control + (size_t)(live ? b : 0) * 2 * N;
        const int wp = live ? wp_id[b] : 0;
        Stage2 s;
        {
            Stage<float> sA, sB;
            stage_zero(sA); stage_zero(sB);
            if (live) {
                const double e_y = spatial[b], e_psi = spatial[(size_t)B + b];
                assemble_stage<float>(sA, mp, pv, 2 * cm.gl, wp, e_y, e_psi, cc, ub + (size_t)b * N, lb + (size_t)b * N);
                assemble_stage<float>(sB, mp, pv, 2 * cm.gl + 1, wp, e_y, e_psi, cc, ub + (size_t)b * N, lb + (size_t)b * N);
            }
            pack_stages(s, sA, sB);
        }
        f2* sm = cold_base + ((size_t)gwarp * G + lane / LPS) * kPairRows * LPS;   // the check-only rows: global memory
        float4* cf = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * PcrCoef<LPS>::kF4 * 32 + lane;
        auto emit = [&](const f2 w[5], const SolveResult& r) {
            if (x_out) {
                double* xo = x_out + (size_t)b * n;
                const int kA = 2 * cm.gl, kB = kA + 1;
                if (kA <= N) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) xo[3 * kA + i] = (double)w[i].x;
                    if (kA < N) { xo[3 * (N + 1) + 2 * kA] = (double)w[3].x; xo[3 * (N + 1) + 2 * kA + 1] = (double)w[4].x; }
                }
                if (kB <= N) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) xo[3 * kB + i] = (double)w[i].y;
                    if (kB < N) { xo[3 * (N + 1) + 2 * kB] = (double)w[3].y; xo[3 * (N + 1) + 2 * kB + 1] = (double)w[4].y; }
                }
            }
            const RolloutArgs ro{rollout_state, spatial, pv.kappa, wp, Ts, B};
            control_epilogue_tm<LPS>(cm, mp, w, r, cc, infeas, u_out, iters, qp_status, flags, b, fl, ro);
        };
        admm_solve_tm<LPS>(cm, s, st, al2, nal2, n, sm, cf, tm, live, emit);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();   // every warp is done with its columns
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm_base_s), "n"(kTmAllocCols));
}

constexpr size_t kTmColdBytesPerWarp = (size_t)kPairRows * 32 * sizeof(f2);
static f2* g_tm_cold = nullptr;
static size_t g_tm_cold_warps = 0;
int reserve_tm_scratch(int B) {
    const size_t warps = (((size_t)(B > 0 ? B : 0) + 1) / 2 + kTmWarps - 1) / kTmWarps * kTmWarps;
    if (warps <= g_tm_cold_warps) return 0;
    f2* p = nullptr;
    if (cudaMalloc(&p, warps * kTmColdBytesPerWarp) != cudaSuccess) { (void)cudaGetLastError(); return MPC_E_CUDA; }
    if (g_tm_cold) cudaFree(g_tm_cold);
    g_tm_cold = p;
    g_tm_cold_warps = warps;
    return 0;
}

void preload_tm_kernels(int N) {
    cudaFuncAttributes fa;
    if (N + 1 <= 32 && N + 1 > 16) cudaFuncGetAttributes(&fa, assemble_solve_tm_kernel<16>);
}

// horizons of 17 .. 32 stages with the reference's unbounded e_psi / t rows; MPC_E_UNSUPPORTED otherwise (the caller then uses
// the paired kernel)
int launch_assemble_solve_tm(const MpcParams& mp, const AdmmSettings& st, const PathView& pv, const double* spatial,
                             const int* wp_id, double* control, const double* ub, const double* lb, int* infeas, double* u_out,
                             double* x_out, int* iters, int* qp_status, int* flags, int B, cudaStream_t s, double* rollout_state,
                             double Ts, const int* order) {
    const int ns = mp.N + 1;
    const bool loose = mp.xmin[1] <= -kOsqpInfty && mp.xmax[1] >= kOsqpInfty && mp.xmin[2] <= -kOsqpInfty &&
                       mp.xmax[2] >= kOsqpInfty;
    if (ns > 32 || ns <= 16 || !loose) return MPC_E_UNSUPPORTED;
    constexpr int LPS = 16, per_block = kTmWarps * (32 / LPS);
    const int grid = (B + per_block - 1) / per_block;
    if ((size_t)grid * kTmWarps > g_tm_cold_warps) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return MPC_E_UNSUPPORTED;
        if (reserve_tm_scratch(B)) return MPC_E_UNSUPPORTED;
    }
    NvtxRange nvtx_("mpc:K1+K2 assemble_solve (paired fp32, constants in tensor memory)");
    const size_t smem = (size_t)kTmWarps * PcrCoef<LPS>::kF4 * 32 * sizeof(float4);
    { static int have_ = 0; ensure_dynamic_smem(assemble_solve_tm_kernel<LPS>, have_, smem); }
    assemble_solve_tm_kernel<LPS><<<grid, 32 * kTmWarps, smem, s>>>(
        mp, st, make_float2((float)st.alpha, (float)st.alpha), make_float2(-(float)st.alpha, -(float)st.alpha), pv, spatial, wp_id,
        control, ub, lb, infeas, u_out, x_out, iters, qp_status, flags, B, rollout_state, Ts, order, g_tm_cold);
    return 0;
}

}  // namespace mpcb
