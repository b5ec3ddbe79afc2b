// admm.cu -- kernel entry points for K1+K2 (see admm.cuh).  FMA contraction is enabled here: the QP
// solution is compared within a tolerance, not bit-for-bit.
#include "engine.h"

namespace mpcb {

constexpr int kWarpsPerBlock = 2;

template <typename T, int NLEV>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
solve_qp_kernel(int N, AdmmSettings st, const double* __restrict__ Pd, const double* __restrict__ q,
                const double* __restrict__ Ax, const double* __restrict__ l, const double* __restrict__ u,
                double* __restrict__ x_out, int* __restrict__ iters, int* __restrict__ status, int B) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (b >= B) return;
    const int n = 5 * N + 3, m = 8 * N + 6, nnz = 16 * N + 6;
    Stage<T> s;
    load_stage_qp<T>(s, N, lane, Pd + (size_t)b * n, q + (size_t)b * n, Ax + (size_t)b * nnz, l + (size_t)b * m,
                     u + (size_t)b * m);
    T w[5];
    const SolveResult r = admm_solve<T, NLEV>(s, st, lane, N + 1, n, w);
    if (x_out && lane <= N) {
        double* xo = x_out + (size_t)b * n;
#pragma unroll
        for (int i = 0; i < 3; ++i) xo[3 * lane + i] = (double)w[i];
        if (lane < N) { xo[3 * (N + 1) + 2 * lane] = (double)w[3]; xo[3 * (N + 1) + 2 * lane + 1] = (double)w[4]; }
    }
    if (lane == 0) {
        if (iters) iters[b] = r.iters;
        if (status) status[b] = r.status;
    }
}

template <typename T, int NLEV>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
assemble_solve_kernel(MpcParams mp, AdmmSettings st, PathView pv, const double* __restrict__ spatial,
                      const int* __restrict__ wp_id, double* __restrict__ control, const double* __restrict__ ub,
                      const double* __restrict__ lb, int* __restrict__ infeas, double* __restrict__ u_out,
                      double* __restrict__ x_out, int* __restrict__ iters, int* __restrict__ qp_status,
                      int* __restrict__ flags, int B) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (b >= B) return;
    int fl = flags ? flags[b] : 0;
    if (fl & (MPC_ST_DEAD | MPC_ST_FINISHED)) return;
    const int N = mp.N, n = 5 * N + 3;
    double* cc = control + (size_t)b * 2 * N;
    Stage<T> s;
    assemble_stage<T>(s, mp, pv, lane, wp_id[b], spatial[b], spatial[(size_t)B + b], cc, ub + (size_t)b * N,
                      lb + (size_t)b * N);
    T w[5];
    const SolveResult r = admm_solve<T, NLEV>(s, st, lane, N + 1, n, w);
    const bool ok = !(r.status == -3 || r.status == -4 || r.status == -7);  // OSQP returns x (MPC.py:185-206)
    if (x_out && lane <= N) {
        double* xo = x_out + (size_t)b * n;
#pragma unroll
        for (int i = 0; i < 3; ++i) xo[3 * lane + i] = (double)w[i];
        if (lane < N) { xo[3 * (N + 1) + 2 * lane] = (double)w[3]; xo[3 * (N + 1) + 2 * lane + 1] = (double)w[4]; }
    }
    int inf = infeas[b];
    __syncwarp();
    if (ok) {
        if (lane < N) {
            cc[2 * lane] = (double)w[3];                         // MPC.py:187
            cc[2 * lane + 1] = atan((double)w[4] * mp.L);        // MPC.py:188-189
        }
        if (lane == 0) {
            u_out[2 * (size_t)b] = (double)w[3];
            u_out[2 * (size_t)b + 1] = atan((double)w[4] * mp.L);
            inf = 0;                                             // MPC.py:206
            fl &= ~MPC_ST_QP_FALLBACK;
        }
    } else if (lane == 0) {
        const int id = 2 * (inf + 1);                            // MPC.py:212-213
        u_out[2 * (size_t)b] = cc[id];
        u_out[2 * (size_t)b + 1] = cc[id + 1];
        inf += 1;                                                // MPC.py:216
        fl |= MPC_ST_QP_FALLBACK;
    }
    if (lane == 0) {
        if (inf == N - 1) fl |= MPC_ST_DEAD;                     // MPC.py:218-220
        infeas[b] = inf;
        if (flags) flags[b] = fl;
        if (iters) iters[b] = r.iters;
        if (qp_status) qp_status[b] = r.status;
    }
}

template <typename T>
static int solve_qp_dispatch(int N, const AdmmSettings& st, const double* Pd, const double* q, const double* Ax,
                             const double* l, const double* u, double* x_out, int* iters, int* status, int B,
                             cudaStream_t s) {
    const int grid = (B + kWarpsPerBlock - 1) / kWarpsPerBlock, block = 32 * kWarpsPerBlock;
    if (N + 1 <= 16) solve_qp_kernel<T, 4><<<grid, block, 0, s>>>(N, st, Pd, q, Ax, l, u, x_out, iters, status, B);
    else if (N + 1 <= 32) solve_qp_kernel<T, 5><<<grid, block, 0, s>>>(N, st, Pd, q, Ax, l, u, x_out, iters, status, B);
    else return MPC_E_UNSUPPORTED;
    return 0;
}

int launch_solve_qp(int precision, int N, const AdmmSettings& st, const double* Pd, const double* q, const double* Ax,
                    const double* l, const double* u, double* x_out, int* iters, int* status, int B, cudaStream_t s) {
    if (precision == 1) return solve_qp_dispatch<double>(N, st, Pd, q, Ax, l, u, x_out, iters, status, B, s);
    return solve_qp_dispatch<float>(N, st, Pd, q, Ax, l, u, x_out, iters, status, B, s);
}

template <typename T>
static int assemble_solve_dispatch(const MpcParams& mp, const AdmmSettings& st, const PathView& pv, const double* spatial,
                                   const int* wp_id, double* control, const double* ub, const double* lb, int* infeas,
                                   double* u_out, double* x_out, int* iters, int* qp_status, int* flags, int B,
                                   cudaStream_t s) {
    const int grid = (B + kWarpsPerBlock - 1) / kWarpsPerBlock, block = 32 * kWarpsPerBlock;
    if (mp.N + 1 <= 16)
        assemble_solve_kernel<T, 4><<<grid, block, 0, s>>>(mp, st, pv, spatial, wp_id, control, ub, lb, infeas, u_out,
                                                           x_out, iters, qp_status, flags, B);
    else if (mp.N + 1 <= 32)
        assemble_solve_kernel<T, 5><<<grid, block, 0, s>>>(mp, st, pv, spatial, wp_id, control, ub, lb, infeas, u_out,
                                                           x_out, iters, qp_status, flags, B);
    else return MPC_E_UNSUPPORTED;
    return 0;
}

int launch_assemble_solve(int precision, const MpcParams& mp, const AdmmSettings& st, const PathView& pv,
                          const double* spatial, const int* wp_id, double* control, const double* ub, const double* lb,
                          int* infeas, double* u_out, double* x_out, int* iters, int* qp_status, int* flags, int B,
                          cudaStream_t s) {
    if (precision == 1)
        return assemble_solve_dispatch<double>(mp, st, pv, spatial, wp_id, control, ub, lb, infeas, u_out, x_out, iters,
                                               qp_status, flags, B, s);
    return assemble_solve_dispatch<float>(mp, st, pv, spatial, wp_id, control, ub, lb, infeas, u_out, x_out, iters,
                                          qp_status, flags, B, s);
}

}  // namespace mpcb
