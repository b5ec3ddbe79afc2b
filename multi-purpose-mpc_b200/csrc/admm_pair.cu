// admm_pair.cu -- kernel entry points of the paired-stage fp32 production path (see admm_pair.cuh).  FMA contraction
// is enabled here: the QP solution is compared within a tolerance, not bit-for-bit.
#include "launch_util.h"
#include "engine.h"
#include "admm_epilogue.cuh"
#include "admm_pair.cuh"
#include <cstdio>
#include <cstdlib>

namespace mpcb {

// ------------------------------------------------------------------------------------------------
// paired-stage kernels (fp32 production path, N + 1 <= 64): a group of LPS lanes per scenario, two stages per
// lane, 32 / LPS scenarios per warp (admm_pair.cuh)
// ------------------------------------------------------------------------------------------------
constexpr int kPairWarpsPerBlock = 1;
template <int LPS> constexpr size_t pair_smem_bytes() {
    return (size_t)kPairWarpsPerBlock * 32 * (kPairRows * sizeof(f2) + MPC_PCR_COEF_SMEM * PcrCoef<LPS>::kF4 * sizeof(float4));
}   // (32 / LPS) groups x [kPairRows][LPS] f2, then per warp [pcr_coef_f4][32 lanes] float4 (PCR coefficients)
template <int LPS> __device__ __forceinline__ float4* pair_coef_ptr(unsigned char* smem_raw, int warp, int lane) {
    return reinterpret_cast<float4*>(smem_raw + (size_t)kPairWarpsPerBlock * 32 * kPairRows * sizeof(f2)) +
           (size_t)warp * PcrCoef<LPS>::kF4 * 32 + lane;
}

__device__ __forceinline__ void write_solution2(int N, int gl, const f2 w[5], double* xo) {
    if (!xo) return;
    const int kA = 2 * gl, kB = kA + 1;
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

// MPC.get_control after the solve (MPC.py:185-222) for the paired layout
template <int LPS>
__device__ __forceinline__ void control_epilogue2(const GroupComm<LPS>& cm, const MpcParams& mp, const f2 w[5],
                                                  const SolveResult& r, double* cc, int* infeas, double* u_out, int* iters,
                                                  int* qp_status, int* flags, int b, int fl, const RolloutArgs& ro) {
    const int N = mp.N, gl = cm.gl, kA = 2 * gl, kB = kA + 1;
    const bool ok = !(r.status == -3 || r.status == -4 || r.status == -7 || r.status == 3 || r.status == 4);  // OSQP returns x (MPC.py:185-206)
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

template <int LPS, int MINB>
__global__ void __launch_bounds__(32 * kPairWarpsPerBlock, MINB)
solve_qp_pair_kernel(int N, AdmmSettings st, const f2 al2, const f2 nal2, const double* __restrict__ Pd, const double* __restrict__ q,
                     const double* __restrict__ Ax, const double* __restrict__ l, const double* __restrict__ u,
                     double* __restrict__ x_out, int* __restrict__ iters, int* __restrict__ status, int B) {
    constexpr int G = 32 / LPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = (blockIdx.x * kPairWarpsPerBlock + warp) * G + lane / LPS;
    const bool live = b < B;  // a group without a scenario runs an all-zero QP alongside (warp-uniform collectives)
    if (!__any_sync(kFull, live)) return;
    const GroupComm<LPS> cm;
    const int n = 5 * N + 3, m = 8 * N + 6, nnz = 16 * N + 6;
    Stage2 s;
    {
        Stage<float> sA, sB;
        stage_zero(sA); stage_zero(sB);
        if (live) {
            load_stage_qp<float>(sA, N, 2 * cm.gl, Pd + (size_t)b * n, q + (size_t)b * n, Ax + (size_t)b * nnz,
                                 l + (size_t)b * m, u + (size_t)b * m);
            load_stage_qp<float>(sB, N, 2 * cm.gl + 1, Pd + (size_t)b * n, q + (size_t)b * n, Ax + (size_t)b * nnz,
                                 l + (size_t)b * m, u + (size_t)b * m);
        }
        pack_stages(s, sA, sB);
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    f2* sm = reinterpret_cast<f2*>(smem_raw) + ((size_t)warp * G + lane / LPS) * kPairRows * LPS;
    auto emit = [&](const f2 w[5], const SolveResult& r) {
        write_solution2(N, cm.gl, w, x_out ? x_out + (size_t)b * n : nullptr);
        if (cm.gl == 0) {
            if (iters) iters[b] = r.iters;
            if (status) status[b] = r.status;
        }
    };
    // rows 1 and 2 (e_psi, t) unbounded at every stage of every scenario of the warp -> the cheaper loop
    const float big = (float)(kOsqpInfty * 0.5);
    const bool tight = live && 2 * cm.gl <= N &&
                       !(s.lo[1].x < -big && s.hi[1].x > big && s.lo[2].x < -big && s.hi[2].x > big &&
                         (2 * cm.gl + 1 > N || (s.lo[1].y < -big && s.hi[1].y > big && s.lo[2].y < -big && s.hi[2].y > big)));
    float4* cf = pair_coef_ptr<LPS>(smem_raw, warp, lane);
    if (!__any_sync(kFull, tight)) admm_solve2<LPS, true>(cm, s, st, al2, nal2, n, sm, cf, live, emit);
    else admm_solve2<LPS, false>(cm, s, st, al2, nal2, n, sm, cf, live, emit);
}

template <int LPS, bool LOOSE, int MINB>
__global__ void __launch_bounds__(32 * kPairWarpsPerBlock, MINB)
assemble_solve_pair_kernel(MpcParams mp, AdmmSettings st, const f2 al2, const f2 nal2, PathView pv, const double* __restrict__ spatial,
                           const int* __restrict__ wp_id, double* __restrict__ control, const double* __restrict__ ub,
                           const double* __restrict__ lb, int* __restrict__ infeas, double* __restrict__ u_out,
                           double* __restrict__ x_out, int* __restrict__ iters, int* __restrict__ qp_status,
                           int* __restrict__ flags, int B, double* __restrict__ rollout_state, double Ts,
                           const int* __restrict__ order, HostIO hio) {
    constexpr int G = 32 / LPS;
    MPC_PHASE_MARK(0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = (blockIdx.x * kPairWarpsPerBlock + warp) * G + lane / LPS;
    // `order` (closed-loop path): scenarios sorted by the length of their previous solve, see geometry.cu::plan_solve_order
    int b = slot < B ? (order ? order[slot] : slot) : B;
    if ((unsigned)b >= (unsigned)B) b = B;  // an `order` entry outside 0..B-1 can only be a stale / corrupt plan: idle group
    const int fl = (b < B && flags) ? flags[b] : 0;
    const bool live = b < B && !(fl & (MPC_ST_DEAD | MPC_ST_FINISHED));
    if (!__any_sync(kFull, live)) return;
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
    f2* sm = reinterpret_cast<f2*>(smem_raw) + ((size_t)warp * G + lane / LPS) * kPairRows * LPS;
    auto emit = [&](const f2 w[5], const SolveResult& r) {
        write_solution2(N, cm.gl, w, x_out ? x_out + (size_t)b * n : nullptr);
        const RolloutArgs ro{rollout_state, spatial, pv.kappa, wp, Ts, B, hio.state, hio.u, hio.flags};
        control_epilogue2<LPS>(cm, mp, w, r, cc, infeas, u_out, iters, qp_status, flags, b, fl, ro);
    };
    admm_solve2<LPS, LOOSE>(cm, s, st, al2, nal2, n, sm, pair_coef_ptr<LPS>(smem_raw, warp, lane), live, emit);
#ifdef MPC_PHASE_CLOCK
    MPC_PHASE_MARK(6);
    if (lane == 0 && (blockIdx.x & 127) == 0) {
        const long long* t = phase_clock_buf();
        printf("cta %4d  assemble %6lld  scale %6lld  factorise %6lld  pass1 %6lld  to-first-result %7lld  emit %6lld  rest %7lld\n",
               (int)blockIdx.x, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[7] - t[3], t[4] - t[3], t[5] - t[4], t[6] - t[5]);
    }
#endif
}

constexpr int kPairMinBlocks = 8;

template <int LPS>
static void solve_qp_pair_launch(int N, const AdmmSettings& st, const double* Pd, const double* q, const double* Ax,
                                 const double* l, const double* u, double* x_out, int* iters, int* status, int B,
                                 cudaStream_t s) {
    constexpr int per_block = kPairWarpsPerBlock * (32 / LPS);
    const size_t smem = pair_smem_bytes<LPS>();
    { static int have_ = 0; ensure_dynamic_smem(solve_qp_pair_kernel<LPS, kPairMinBlocks>, have_, smem); }
    solve_qp_pair_kernel<LPS, kPairMinBlocks><<<(B + per_block - 1) / per_block, 32 * kPairWarpsPerBlock, smem, s>>>(
        N, st, make_float2((float)st.alpha, (float)st.alpha), make_float2(-(float)st.alpha, -(float)st.alpha), Pd, q, Ax, l, u,
        x_out, iters, status, B);
}

template <int LPS, bool LOOSE>
static void assemble_solve_pair_launch(const MpcParams& mp, const AdmmSettings& st, const PathView& pv,
                                       const double* spatial, const int* wp_id, double* control, const double* ub,
                                       const double* lb, int* infeas, double* u_out, double* x_out, int* iters,
                                       int* qp_status, int* flags, int B, cudaStream_t s, double* rs, double Ts,
                                       const int* order, HostIO hio) {
    constexpr int per_block = kPairWarpsPerBlock * (32 / LPS);
    const size_t smem = pair_smem_bytes<LPS>();
    { static int have_ = 0; ensure_dynamic_smem(assemble_solve_pair_kernel<LPS, LOOSE, kPairMinBlocks>, have_, smem); }
    assemble_solve_pair_kernel<LPS, LOOSE, kPairMinBlocks><<<(B + per_block - 1) / per_block, 32 * kPairWarpsPerBlock, smem, s>>>(
        mp, st, make_float2((float)st.alpha, (float)st.alpha), make_float2(-(float)st.alpha, -(float)st.alpha), pv, spatial, wp_id,
        control, ub, lb, infeas, u_out, x_out, iters, qp_status, flags, B, rs, Ts, order, hio);
}

void preload_pair_kernels(int N) {
    const int ns = N + 1;
    cudaFuncAttributes fa;
    if (ns <= 16) { cudaFuncGetAttributes(&fa, assemble_solve_pair_kernel<8, true, kPairMinBlocks>); cudaFuncGetAttributes(&fa, assemble_solve_pair_kernel<8, false, kPairMinBlocks>); }
    else if (ns <= 32) { cudaFuncGetAttributes(&fa, assemble_solve_pair_kernel<16, true, kPairMinBlocks>); cudaFuncGetAttributes(&fa, assemble_solve_pair_kernel<16, false, kPairMinBlocks>); }
    else if (ns <= 64) { cudaFuncGetAttributes(&fa, assemble_solve_pair_kernel<32, true, kPairMinBlocks>); cudaFuncGetAttributes(&fa, assemble_solve_pair_kernel<32, false, kPairMinBlocks>); }
}

int launch_solve_qp_pair(int N, const AdmmSettings& st, const double* Pd, const double* q, const double* Ax, const double* l,
                         const double* u, double* x_out, int* iters, int* status, int B, cudaStream_t s) {
    NvtxRange nvtx_("mpc:K2 solve_qp (paired fp32)");
    const int ns = N + 1;
    if (ns > 64) return MPC_E_UNSUPPORTED;
    if (ns <= 16) solve_qp_pair_launch<8>(N, st, Pd, q, Ax, l, u, x_out, iters, status, B, s);
    else if (ns <= 32) solve_qp_pair_launch<16>(N, st, Pd, q, Ax, l, u, x_out, iters, status, B, s);
    else solve_qp_pair_launch<32>(N, st, Pd, q, Ax, l, u, x_out, iters, status, B, s);
    return 0;
}

int launch_assemble_solve_pair(const MpcParams& mp, const AdmmSettings& st, const PathView& pv, const double* spatial,
                               const int* wp_id, double* control, const double* ub, const double* lb, int* infeas,
                               double* u_out, double* x_out, int* iters, int* qp_status, int* flags, int B, cudaStream_t s,
                               double* rollout_state, double Ts, const int* order, const HostIO* host_io) {
    NvtxRange nvtx_("mpc:K1+K2 assemble_solve (paired fp32)");
    const HostIO hio = host_io ? *host_io : HostIO{nullptr, nullptr, nullptr};
    const int ns = mp.N + 1;
    if (ns > 64) return MPC_E_UNSUPPORTED;
    // e_psi and t unbounded (the reference's StateConstraints): OSQP's "loose" rows, skipped by the loop
    const bool loose = mp.xmin[1] <= -kOsqpInfty && mp.xmax[1] >= kOsqpInfty && mp.xmin[2] <= -kOsqpInfty &&
                       mp.xmax[2] >= kOsqpInfty;
#define PAIR_GO(LPS_) do { if (loose) assemble_solve_pair_launch<LPS_, true>(mp, st, pv, spatial, wp_id, control, ub, lb, infeas, u_out, x_out, iters, qp_status, flags, B, s, rollout_state, Ts, order, hio); \
                           else assemble_solve_pair_launch<LPS_, false>(mp, st, pv, spatial, wp_id, control, ub, lb, infeas, u_out, x_out, iters, qp_status, flags, B, s, rollout_state, Ts, order, hio); } while (0)
    if (ns <= 16) PAIR_GO(8); else if (ns <= 32) PAIR_GO(16); else PAIR_GO(32);
#undef PAIR_GO
    return 0;
}

}  // namespace mpcb
