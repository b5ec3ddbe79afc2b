// admm_quad.cu -- kernel entry points of the four-stages-per-lane fp32 path (see admm_quad.cuh).  FMA contraction is enabled
// here: the QP solution is compared within a tolerance, not bit-for-bit.
#include "launch_util.h"
#include "engine.h"
#include "admm_epilogue.cuh"
#include "admm_quad.cuh"
#include <cstdio>
#include <cstdlib>

namespace mpcb {

constexpr int kQuadMinBlocks = 8;
template <int LPS> constexpr size_t quad_smem_bytes() { return (size_t)QuadHot<LPS>::kF4 * 32 * sizeof(float4); }  // HOT columns
constexpr size_t kQuadColdBytesPerWarp = (size_t)kQuadCold * 32 * sizeof(float4);

// COLD columns of every warp of a launch: one process drives one GPU, so a single grow-only device buffer serves all
// engines of the process (never shrunk; released with the context).  Grown outside stream capture only.
static float4* g_cold = nullptr;
static size_t g_cold_warps = 0;
int reserve_quad_scratch(int B) {
    const size_t warps = ((size_t)(B > 0 ? B : 0) + 3) / 4;
    if (warps <= g_cold_warps) return 0;
    float4* p = nullptr;
    if (cudaMalloc(&p, warps * kQuadColdBytesPerWarp) != cudaSuccess) { (void)cudaGetLastError(); return MPC_E_CUDA; }
    if (g_cold) cudaFree(g_cold);  // implicit device synchronisation: no launch is still using the old buffer
    g_cold = p;
    g_cold_warps = warps;
    return 0;
}

// stage indices of the lane: slice 0 (E) = (4 gl, 4 gl + 2), slice 1 (O) = (4 gl + 1, 4 gl + 3)
__device__ __forceinline__ int quad_stage(int gl, int sl, int half) { return 4 * gl + sl + 2 * half; }

__device__ __forceinline__ void write_solution4(int N, int gl, const f2 (&w)[2][5], double* xo) {
    if (!xo) return;
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = quad_stage(gl, sl, h);
            if (k > N) continue;
#pragma unroll
            for (int i = 0; i < 3; ++i) xo[3 * k + i] = (double)(h ? w[sl][i].y : w[sl][i].x);
            if (k < N) {
                xo[3 * (N + 1) + 2 * k] = (double)(h ? w[sl][3].y : w[sl][3].x);
                xo[3 * (N + 1) + 2 * k + 1] = (double)(h ? w[sl][4].y : w[sl][4].x);
            }
        }
}

// MPC.get_control after the solve (MPC.py:185-222), four stages per lane
template <int LPS>
__device__ __forceinline__ void control_epilogue4(const QuadComm<LPS>& cm, const MpcParams& mp, const f2 (&w)[2][5],
                                                  const SolveResult& r, double* cc, int* infeas, double* u_out, int* iters,
                                                  int* qp_status, int* flags, int b, int fl, const RolloutArgs& ro) {
    const int N = mp.N, gl = cm.gl;
    const bool ok = !(r.status == -3 || r.status == -4 || r.status == -7 || r.status == 3 || r.status == 4);  // MPC.py:185-206
    int inf = infeas[b];
    if (ok) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = quad_stage(gl, sl, h);
                if (k < N) {
                    cc[2 * k] = (double)(h ? w[sl][3].y : w[sl][3].x);                       // MPC.py:187
                    cc[2 * k + 1] = atan((double)(h ? w[sl][4].y : w[sl][4].x) * mp.L);      // MPC.py:188-189
                }
            }
        if (gl == 0) {
            u_out[2 * (size_t)b] = (double)w[0][3].x;
            u_out[2 * (size_t)b + 1] = atan((double)w[0][4].x * mp.L);
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

template <int LPS, int MINB>
__global__ void __launch_bounds__(32, MINB)
assemble_solve_quad_kernel(MpcParams mp, AdmmSettings st, const f2 al2, const f2 nal2, PathView pv, const double* __restrict__ spatial,
                           const int* __restrict__ wp_id, double* __restrict__ control, const double* __restrict__ ub,
                           const double* __restrict__ lb, int* __restrict__ infeas, double* __restrict__ u_out,
                           double* __restrict__ x_out, int* __restrict__ iters, int* __restrict__ qp_status,
                           int* __restrict__ flags, int B, double* __restrict__ rollout_state, double Ts,
                           const int* __restrict__ order, float4* __restrict__ cold_base, int zero) {
    constexpr int G = 32 / LPS;
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * G + lane / LPS;
    int b = slot < B ? (order ? order[slot] : slot) : B;
    if ((unsigned)b >= (unsigned)B) b = B;
    const int fl = (b < B && flags) ? flags[b] : 0;
    const bool live = b < B && !(fl & (MPC_ST_DEAD | MPC_ST_FINISHED));
    if (!__any_sync(kFull, live)) return;
    const QuadComm<LPS> cm;
    const int N = mp.N, n = 5 * N + 3;
    double* cc = control + (size_t)(live ? b : 0) * 2 * N;
    const int wp = live ? wp_id[b] : 0;
    Stage2 s[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        Stage<float> sA, sB;
        stage_zero(sA); stage_zero(sB);
        if (live) {
            const double e_y = spatial[b], e_psi = spatial[(size_t)B + b];
            assemble_stage<float>(sA, mp, pv, quad_stage(cm.gl, sl, 0), wp, e_y, e_psi, cc, ub + (size_t)b * N, lb + (size_t)b * N);
            assemble_stage<float>(sB, mp, pv, quad_stage(cm.gl, sl, 1), wp, e_y, e_psi, cc, ub + (size_t)b * N, lb + (size_t)b * N);
        }
        pack_stages(s[sl], sA, sB);
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* hot = reinterpret_cast<float4*>(smem_raw) + lane;
    float4* cold = cold_base + (size_t)blockIdx.x * kQuadCold * 32 + lane;
    auto emit = [&](const f2 (&w)[2][5], const SolveResult& r) {
        write_solution4(N, cm.gl, w, x_out ? x_out + (size_t)b * n : nullptr);
        const RolloutArgs ro{rollout_state, spatial, pv.kappa, wp, Ts, B};
        control_epilogue4<LPS>(cm, mp, w, r, cc, infeas, u_out, iters, qp_status, flags, b, fl, ro);
    };
    admm_solve4<LPS>(cm, s, st, al2, nal2, n, hot, cold, live, zero, emit);
}

void preload_quad_kernels(int N) {
    cudaFuncAttributes fa;
    if (N + 1 <= 32) cudaFuncGetAttributes(&fa, assemble_solve_quad_kernel<8, kQuadMinBlocks>);
}

// four stages per lane: horizons up to 31 intervals with the reference's unbounded e_psi / t rows; returns MPC_E_UNSUPPORTED
// otherwise (the caller then uses the paired kernel)
int launch_assemble_solve_quad(const MpcParams& mp, const AdmmSettings& st, const PathView& pv, const double* spatial,
                               const int* wp_id, double* control, const double* ub, const double* lb, int* infeas,
                               double* u_out, double* x_out, int* iters, int* qp_status, int* flags, int B, cudaStream_t s,
                               double* rollout_state, double Ts, const int* order) {
    const int ns = mp.N + 1;
    const bool loose = mp.xmin[1] <= -kOsqpInfty && mp.xmax[1] >= kOsqpInfty && mp.xmin[2] <= -kOsqpInfty &&
                       mp.xmax[2] >= kOsqpInfty;
    if (ns > 32 || !loose) return MPC_E_UNSUPPORTED;
    if (((size_t)B + 3) / 4 > g_cold_warps) {  // direct ABI call with a batch nobody reserved for: grow unless capturing
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return MPC_E_UNSUPPORTED;
        if (reserve_quad_scratch(B)) return MPC_E_UNSUPPORTED;
    }
    NvtxRange nvtx_("mpc:K1+K2 assemble_solve (four stages per lane, fp32)");
    constexpr int LPS = 8, per_block = 32 / LPS;
    const size_t smem = quad_smem_bytes<LPS>();
    { static int have_ = 0; ensure_dynamic_smem(assemble_solve_quad_kernel<LPS, kQuadMinBlocks>, have_, smem); }
    assemble_solve_quad_kernel<LPS, kQuadMinBlocks><<<(B + per_block - 1) / per_block, 32, smem, s>>>(
        mp, st, make_float2((float)st.alpha, (float)st.alpha), make_float2(-(float)st.alpha, -(float)st.alpha), pv, spatial, wp_id,
        control, ub, lb, infeas, u_out, x_out, iters, qp_status, flags, B, rollout_state, Ts, order, g_cold, 0);
    return 0;
}

}  // namespace mpcb
