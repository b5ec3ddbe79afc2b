// admm.cuh -- K1 (LTV assembly) + K2 (OSQP-equivalent ADMM) for sm_100a.
//
// One warp per scenario, one lane per horizon stage (N + 1 <= 32).  Lane j keeps the whole of
// stage j of the QP that MPC._init_problem builds (reference: src/MPC.py:61-159) in registers:
//   variables w_j = (e_y, e_psi, t, v, kappa)_j              (stage N has no inputs)
//   rows      dynamics block j (3 equalities, MPC.py:128-131,142-147) + 5 bound rows (MPC.py:133)
// The OSQP iteration (restated in oracle/osqp_oracle.c; executable model in
// tools/admm_pcr_model.py) runs entirely in registers; neighbouring stages talk through warp
// shuffles.  The reduced KKT system (P + sigma I + A' diag(rho) A) x = b is block tridiagonal over
// stages: the two inputs of a stage are eliminated locally (their 2x2 block is diagonal), and the
// remaining chain of 3x3 blocks is solved by parallel cyclic reduction (log2(32) = 5 levels), so
// one linear solve is ~40 shuffles + ~130 FMAs per lane instead of a 31-step serial Riccati sweep.
// The iteration is written in increment form (see admm_solve) so that fp32 reproduces OSQP's
// iteration counts and infeasibility certificates at the reference's default tolerance; fp64 is
// the validation path (and the one to use for eps < 1e-4).
// Norms for termination / rho adaptation are warp reductions (redux.sync on the float bit pattern
// for fp32).  No tensor cores: per-instance 3x3 / 5x5 blocks are not a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mpcb {

constexpr double kOsqpInfty = 1e30;
constexpr double kMinScaling = 1e-4;
constexpr double kMaxScaling = 1e4;
constexpr double kRhoMin = 1e-6;
constexpr double kRhoMax = 1e6;
constexpr double kRhoEqOverIneq = 1e3;
constexpr double kRhoTol = 1e-4;
constexpr unsigned kFull = 0xffffffffu;

struct AdmmSettings {
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
    int max_iter, scaling, check_termination, adaptive_rho_interval;
};

struct MpcParams {  // MPC.__init__ arguments (MPC.py:15-59) + car geometry
    int N;
    double Q[3], R[2], QN[3], xmin[3], xmax[3], umin[2], umax[2];
    double ay_max, L;
};

struct PathView {  // device tables, one row per quantity (set by mpc_set_path)
    int n_wp, circular;
    const double *x, *y, *psi, *kappa, *v_ref, *ds_next, *cos_psi, *sin_psi, *cos_ub, *sin_ub, *cos_lb, *sin_lb;
    const double *length_cum;
    const double *border;  // [n_wp][4]
    // [n_wp][kStageTab]: what K1 needs per waypoint and what depends on the path, v_ref and R only -- the entries of A_lin,
    // B_lin, uq and q of MPC.py:96-108,125-131 -- evaluated once in fp64 by stage_table_kernel (admm.cu) instead of twice
    // per lane in front of every solve (four fp64 divisions per stage on the critical path of the solve kernel)
    const double *stage_tab;
};
// row of PathView::stage_tab for waypoint w:  ds | -(kappa^2) ds | -kappa / v_ref ds | -1 / v_ref^2 ds | -R0 v_ref | -R1 kappa
// | ds kappa | (-1 / v_ref^2 ds) v_ref - (1 / v_ref ds)
constexpr int kStageTab = 8;
__host__ __device__ __forceinline__ void stage_coefficients(double ds, double kap, double vr, double R0, double R1, double* o) {
    o[0] = ds;
    o[1] = -(kap * kap) * ds;               // A_lin[1][0]   sbm.py:404
    o[2] = -kap / vr * ds;                  // A_lin[2][0]   sbm.py:406
    o[3] = -1 / (vr * vr) * ds;             // B_lin[2][0]   sbm.py:410
    o[4] = -R0 * vr;                        // q of v        MPC.py:129-131
    o[5] = -R1 * kap;                       // q of kappa
    o[6] = ds * kap;                        // uq[1] = B_lin[1][1] kappa_ref            MPC.py:107-108
    o[7] = (-1 / (vr * vr) * ds) * vr - (1 / vr * ds);  // uq[2] = B_lin[2][0] v_ref - f[2]
}

// ------------------------------------------------------------------------------------------------
// neighbour exchange + reductions between the stages of one scenario
// ------------------------------------------------------------------------------------------------
// WarpComm : one warp per scenario, stage = lane (N + 1 <= 32): shuffles and redux.sync.
// BlockComm: one CTA of NT threads per scenario, stage = thread (N + 1 <= NT <= 128): double-buffered
//            shared-memory exchange, one __syncthreads per exchange.  Same math, longer horizons.
struct WarpComm {
    static constexpr int W = 32;
    __device__ __forceinline__ explicit WarpComm(void*) {}
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    template <typename T> __device__ __forceinline__ T up(T v, int s) { return __shfl_up_sync(kFull, v, s); }
    template <typename T> __device__ __forceinline__ T dn(T v, int s) { return __shfl_down_sync(kFull, v, s); }
    __device__ __forceinline__ float max(float v) {  // v >= 0: order of the bit pattern = order of the value
        return __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(v)));
    }
    __device__ __forceinline__ double max(double v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, s));
        return v;
    }
    template <typename T> __device__ __forceinline__ T sum(T v) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(kFull, v, s);
        return v;
    }
    __device__ __forceinline__ bool any(bool p) { return __any_sync(kFull, p); }
    __device__ __forceinline__ void sync() { __syncwarp(); }
};

template <int NT> struct BlockComm {
    static constexpr int W = NT;
    double* buf;  // [2][NT] exchange slots + [2][NT / 32] reduction slots (8-byte slots for either precision)
    int flip;
    __device__ __forceinline__ explicit BlockComm(void* smem) : buf(reinterpret_cast<double*>(smem)), flip(0) {}
    __host__ __device__ static constexpr int slots() { return 2 * NT + 2 * (NT / 32); }
    __device__ __forceinline__ int lane() const { return threadIdx.x; }
    template <typename T> __device__ __forceinline__ T up(T v, int s) {
        T* b = reinterpret_cast<T*>(buf + flip * NT);
        b[threadIdx.x] = v;
        __syncthreads();
        const T r = (int)threadIdx.x >= s ? b[threadIdx.x - s] : v;
        flip ^= 1;
        return r;
    }
    template <typename T> __device__ __forceinline__ T dn(T v, int s) {
        T* b = reinterpret_cast<T*>(buf + flip * NT);
        b[threadIdx.x] = v;
        __syncthreads();
        const T r = (int)threadIdx.x + s < NT ? b[threadIdx.x + s] : v;
        flip ^= 1;
        return r;
    }
    template <typename T, typename Op> __device__ __forceinline__ T reduce(T v, Op op) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v = op(v, __shfl_xor_sync(kFull, v, s));
        T* b = reinterpret_cast<T*>(buf + 2 * NT + flip * (NT / 32));
        if ((threadIdx.x & 31) == 0) b[threadIdx.x >> 5] = v;
        __syncthreads();
        T r = b[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) r = op(r, b[w]);
        flip ^= 1;
        return r;
    }
    __device__ __forceinline__ float max(float v) { return reduce(v, [](float a, float b) { return fmaxf(a, b); }); }
    __device__ __forceinline__ double max(double v) { return reduce(v, [](double a, double b) { return fmax(a, b); }); }
    template <typename T> __device__ __forceinline__ T sum(T v) { return reduce(v, [](T a, T b) { return a + b; }); }
    __device__ __forceinline__ bool any(bool p) { return __syncthreads_or(p) != 0; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

__device__ __forceinline__ float tabs(float v) { return fabsf(v); }   // clears the sign of -0.0 too (redux.max on bits)
__device__ __forceinline__ double tabs(double v) { return fabs(v); }
__device__ __forceinline__ float tmax(float a, float b) { return fmaxf(a, b); }   // FMNMX, no compare+select
__device__ __forceinline__ double tmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float tmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double tmin(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float trsqrt(float v) {  // bare MUFU.RSQ (2 ulp): scaling factors only; the argument is clamped
    float r;                                         // to [1e-4, 1e4], so rsqrtf()'s denormal fix-up is dead weight
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ double trsqrt(double v) { return 1.0 / sqrt(v); }
template <typename T> __device__ __forceinline__ T limit_scaling(T v) {
    v = v < T(kMinScaling) ? T(1) : v;
    return v > T(kMaxScaling) ? T(kMaxScaling) : v;
}

// ------------------------------------------------------------------------------------------------
// per-lane stage data
// ------------------------------------------------------------------------------------------------
// a[8]: nonzeros of [A_j B_j] (rows of dynamics block j+1, sbm.py:404-410):
//   a0=A00 a1=A01 | a2=A10 a3=A11 a6=B11(kappa) | a4=A20 a5=A22 a7=B20(v)
// c[3]: the -I entries of dynamics block j on x_j;  e[5]: identity (bound) rows.
template <typename T> struct Stage {
    T a[8], c[3], e[5];
    T P[5], q[5];
    T d[3];          // rhs of dynamics block j (l = u)
    T lo[5], hi[5];  // bound rows
    T D[5], Ed[3], Eb[5];
    T cs;            // cost scaling c
    int ctype[5];    // -1 loose, 0 inequality, 1 equality (bound rows; dynamics rows are always 1)
};

// PCR factor.  Levels [0, RLEV) live in registers, levels [RLEV, NLEV) in shared memory laid out
// [coefficient][lane] (conflict-free, one 128 B / 256 B wavefront per coefficient); the split trades
// registers (occupancy) against shared-memory bandwidth and is chosen per precision at compile time.
template <typename T, int NLEV, int RLEV> struct Factor {
    T al[RLEV > 0 ? RLEV : 1][9], be[RLEV > 0 ? RLEV : 1][9];
    T* fs;                 // this warp's shared slab: [(NLEV - RLEV) * 18][32]
    T Dinv[6];             // symmetric: 00 01 02 11 12 22
    T iv, ik;              // 1 / S_vv, 1 / S_kk
    T sxv0, sxv2, sxk0, sxk1;  // S_xu nonzeros
    T fv, fk;              // coupling of (v, kappa)_j to (t, e_psi)_{j+1}
};

// per-scenario shared constants (read only at termination checks): [29][W]
//   0..2 d | 3..7 D | 8..10 Ed | 11..15 Eb | 16..20 1/D | 21..23 1/Ed | 24..28 1/Eb
//   29..33 alpha D | 34..36 dy (dynamics rows) | 37..41 dy (bound rows) of the LAST pass, for OSQP's end-of-loop checks
constexpr int kConstRows = 42;
template <int NLEV, int RLEV> __host__ __device__ constexpr int smem_rows() { return kConstRows + 18 * (NLEV - RLEV); }

template <typename T> __device__ __forceinline__ T tfma(T a, T b, T c);
template <> __device__ __forceinline__ float tfma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double tfma<double>(double a, double b, double c) { return fma(a, b, c); }

// z = A w for this lane: zd (dynamics block j) and zb (bound rows)
template <typename T, typename Comm>
__device__ __forceinline__ void A_apply(Comm& cm, const Stage<T>& s, const T w[5], int lane, T zd[3], T zb[5]) {
    T o0 = s.a[0] * w[0] + s.a[1] * w[1];
    T o1 = s.a[2] * w[0] + s.a[3] * w[1] + s.a[6] * w[4];
    T o2 = s.a[4] * w[0] + s.a[5] * w[2] + s.a[7] * w[3];
    o0 = cm.up(o0, 1); o1 = cm.up(o1, 1); o2 = cm.up(o2, 1);
    if (lane == 0) { o0 = T(0); o1 = T(0); o2 = T(0); }
    zd[0] = s.c[0] * w[0] + o0; zd[1] = s.c[1] * w[1] + o1; zd[2] = s.c[2] * w[2] + o2;
#pragma unroll
    for (int i = 0; i < 5; ++i) zb[i] = s.e[i] * w[i];
}

// r = A' y for this lane's 5 variables
template <typename T, typename Comm>
__device__ __forceinline__ void At_apply(Comm& cm, const Stage<T>& s, const T yd[3], const T yb[5], T r[5]) {
    T g0 = cm.dn(yd[0], 1), g1 = cm.dn(yd[1], 1), g2 = cm.dn(yd[2], 1);
    // lanes whose successor is outside the chain have a == 0, so the shuffled value is harmless
    r[0] = s.c[0] * yd[0] + s.a[0] * g0 + s.a[2] * g1 + s.a[4] * g2 + s.e[0] * yb[0];
    r[1] = s.c[1] * yd[1] + s.a[1] * g0 + s.a[3] * g1 + s.e[1] * yb[1];
    r[2] = s.c[2] * yd[2] + s.a[5] * g2 + s.e[2] * yb[2];
    r[3] = s.a[7] * g2 + s.e[3] * yb[3];
    r[4] = s.a[6] * g1 + s.e[4] * yb[4];
}

// OSQP scale_data (Ruiz equilibration of the KKT matrix + cost scaling), stage layout.
template <typename T, typename Comm>
__device__ __forceinline__ void ruiz_scale(Comm& cm, Stage<T>& s, int iters, int nvar) {
#pragma unroll
    for (int i = 0; i < 5; ++i) { s.D[i] = T(1); s.Eb[i] = T(1); }
#pragma unroll
    for (int i = 0; i < 3; ++i) s.Ed[i] = T(1);
    s.cs = T(1);
    for (int it = 0; it < iters; ++it) {
        T aa[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) aa[i] = tabs(s.a[i]);
        T col[5];
        col[0] = tmax(tmax(aa[0], aa[2]), aa[4]);
        col[1] = tmax(aa[1], aa[3]);
        col[2] = aa[5];
        col[3] = aa[7];
        col[4] = aa[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) col[i] = tmax(col[i], tabs(s.c[i]));
#pragma unroll
        for (int i = 0; i < 5; ++i) col[i] = tmax(tmax(col[i], tabs(s.e[i])), tabs(s.P[i]));
        T ro[3];
        ro[0] = tmax(aa[0], aa[1]);
        ro[1] = tmax(tmax(aa[2], aa[3]), aa[6]);
        ro[2] = tmax(tmax(aa[4], aa[5]), aa[7]);
        const int lane = cm.lane();
        T Dt[5], Edt[3], Ebt[5];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            T r = cm.up(ro[i], 1);
            if (lane == 0) r = T(0);
            Edt[i] = trsqrt(limit_scaling(tmax(tabs(s.c[i]), r)));
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            Dt[i] = trsqrt(limit_scaling(col[i]));
            Ebt[i] = trsqrt(limit_scaling(tabs(s.e[i])));
        }
        T En[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) En[i] = cm.dn(Edt[i], 1);
#pragma unroll
        for (int i = 0; i < 5; ++i) s.P[i] = s.P[i] * Dt[i] * Dt[i];
        s.a[0] = s.a[0] * En[0] * Dt[0]; s.a[1] = s.a[1] * En[0] * Dt[1];
        s.a[2] = s.a[2] * En[1] * Dt[0]; s.a[3] = s.a[3] * En[1] * Dt[1];
        s.a[4] = s.a[4] * En[2] * Dt[0]; s.a[5] = s.a[5] * En[2] * Dt[2];
        s.a[6] = s.a[6] * En[1] * Dt[4]; s.a[7] = s.a[7] * En[2] * Dt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { s.c[i] = s.c[i] * Edt[i] * Dt[i]; s.Ed[i] *= Edt[i]; }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            s.e[i] = s.e[i] * Ebt[i] * Dt[i];
            s.q[i] *= Dt[i];
            s.D[i] *= Dt[i];
            s.Eb[i] *= Ebt[i];
        }
        // cost scaling
        T sp = T(0), mq = T(0);
#pragma unroll
        for (int i = 0; i < 5; ++i) { sp += tabs(s.P[i]); mq = tmax(mq, tabs(s.q[i])); }
        sp = cm.sum(sp) / T(nvar);
        mq = limit_scaling(cm.max(mq));
        T ct = limit_scaling(tmax(sp, mq));
        ct = T(1) / ct;
#pragma unroll
        for (int i = 0; i < 5; ++i) { s.P[i] *= ct; s.q[i] *= ct; }
        s.cs *= ct;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) s.d[i] *= s.Ed[i];
#pragma unroll
    for (int i = 0; i < 5; ++i) { s.lo[i] *= s.Eb[i]; s.hi[i] *= s.Eb[i]; }
}

// 3x3 helpers (row-major)
template <typename T> __device__ __forceinline__ void mm3(const T* A, const T* B, T* C) {  // C = A B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <typename T> __device__ __forceinline__ void inv3sym(const T* M, T* R) {  // full 3x3 in / out
    T a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[8];
    T A = d * f - e * e, B = c * e - b * f, C = b * e - c * d;
    T r = T(1) / (a * A + b * B + c * C);
    R[0] = A * r; R[1] = R[3] = B * r; R[2] = R[6] = C * r;
    R[4] = (a * f - c * c) * r; R[5] = R[7] = (b * c - a * e) * r; R[8] = (a * d - b * b) * r;
}

// Build S = P + sigma I + A' R A for this lane's stage, eliminate the inputs, PCR-factorise.
template <typename T, int NLEV, int RLEV, typename Comm>
__device__ __forceinline__ void factorize(Comm& cm, const Stage<T>& s, Factor<T, NLEV, RLEV>& f, T sigma, T rd,
                                          const T rb[5], int lane, int nstage) {
    constexpr int W = Comm::W;
    const T* a = s.a;
    T diag[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) diag[i] = s.P[i] + sigma + rb[i] * s.e[i] * s.e[i];
    T cn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) cn[i] = cm.dn(s.c[i], 1);
    T Dm[9], U[9], Lo[9];
    Dm[0] = diag[0] + rd * (s.c[0] * s.c[0] + a[0] * a[0] + a[2] * a[2] + a[4] * a[4]);
    Dm[4] = diag[1] + rd * (s.c[1] * s.c[1] + a[1] * a[1] + a[3] * a[3]);
    Dm[8] = diag[2] + rd * (s.c[2] * s.c[2] + a[5] * a[5]);
    Dm[1] = Dm[3] = rd * (a[0] * a[1] + a[2] * a[3]);
    Dm[2] = Dm[6] = rd * (a[4] * a[5]);
    Dm[5] = Dm[7] = T(0);
    T Svv = diag[3] + rd * a[7] * a[7];
    T Skk = diag[4] + rd * a[6] * a[6];
    f.sxv0 = rd * a[4] * a[7]; f.sxv2 = rd * a[5] * a[7];
    f.sxk0 = rd * a[2] * a[6]; f.sxk1 = rd * a[3] * a[6];
    f.fv = rd * a[7] * cn[2];
    f.fk = rd * a[6] * cn[1];
    f.iv = T(1) / Svv;
    f.ik = T(1) / Skk;
    // coupling block (row j, col j+1): F_x[i][r] = rd * (coef of x_i in row r of block j+1) * c_{j+1}[r]
    U[0] = rd * a[0] * cn[0]; U[1] = rd * a[2] * cn[1]; U[2] = rd * a[4] * cn[2];
    U[3] = rd * a[1] * cn[0]; U[4] = rd * a[3] * cn[1]; U[5] = T(0);
    U[6] = T(0);              U[7] = T(0);              U[8] = rd * a[5] * cn[2];
    // Schur complement of the (diagonal) input block
    {
        T sxv[3] = {f.sxv0, T(0), f.sxv2}, sxk[3] = {f.sxk0, f.sxk1, T(0)};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) Dm[3 * i + k] -= f.iv * sxv[i] * sxv[k] + f.ik * sxk[i] * sxk[k];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            U[3 * i + 2] -= f.iv * f.fv * sxv[i];
            U[3 * i + 1] -= f.ik * f.fk * sxk[i];
        }
        T add1 = cm.up(f.ik * f.fk * f.fk, 1), add2 = cm.up(f.iv * f.fv * f.fv, 1);
        if (lane > 0) { Dm[4] -= add1; Dm[8] -= add2; }
    }
    // Lo = coupling block (row j, col j-1) = U_{j-1}'
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T v = cm.up(U[3 * k + i], 1);
            Lo[3 * i + k] = lane > 0 ? v : T(0);
        }
    if (lane >= nstage - 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) U[i] = T(0);
    }
    if (lane >= nstage) {
#pragma unroll
        for (int i = 0; i < 9; ++i) Lo[i] = T(0);
    }
#pragma unroll
    for (int lev = 0; lev < NLEV; ++lev) {
        const int sft = 1 << lev;
        T Di[9];
        inv3sym(Dm, Di);
        T Dup[9], Ddn[9], Uup[9], Ldn[9], Lup[9], Udn[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            Dup[i] = cm.up(Di[i], sft); Ddn[i] = cm.dn(Di[i], sft);
            Uup[i] = cm.up(U[i], sft);  Ldn[i] = cm.dn(Lo[i], sft);
            Lup[i] = cm.up(Lo[i], sft); Udn[i] = cm.dn(U[i], sft);
        }
        const bool has_up = lane >= sft, has_dn = lane + sft < W;
        T al[9], be[9];
        mm3(Lo, Dup, al);
        mm3(U, Ddn, be);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (!has_up) al[i] = T(0);
            if (!has_dn) be[i] = T(0);
        }
        T t1[9], t2[9];
        mm3(al, Uup, t1);
        mm3(be, Ldn, t2);
#pragma unroll
        for (int i = 0; i < 9; ++i) Dm[i] -= t1[i] + t2[i];
        mm3(al, Lup, t1);
        mm3(be, Udn, t2);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            Lo[i] = -t1[i]; U[i] = -t2[i];
            if (lev < RLEV) { f.al[lev < RLEV ? lev : 0][i] = al[i]; f.be[lev < RLEV ? lev : 0][i] = be[i]; }
            else { f.fs[((lev - RLEV) * 18 + i) * W + lane] = al[i]; f.fs[((lev - RLEV) * 18 + 9 + i) * W + lane] = be[i]; }
        }
    }
    T Di[9];
    inv3sym(Dm, Di);
    f.Dinv[0] = Di[0]; f.Dinv[1] = Di[1]; f.Dinv[2] = Di[2]; f.Dinv[3] = Di[4]; f.Dinv[4] = Di[5]; f.Dinv[5] = Di[8];
    cm.sync();
}

// x = S^-1 b
template <typename T, int NLEV, int RLEV, typename Comm>
__device__ __forceinline__ void kkt_solve(Comm& cm, const Factor<T, NLEV, RLEV>& f, const T b[5], int lane, T x[5]) {
    constexpr int W = Comm::W;
    const T bv = f.iv * b[3], bk = f.ik * b[4];
    T bx0 = tfma(-bk, f.sxk0, tfma(-bv, f.sxv0, b[0]));
    T bx1 = tfma(-bk, f.sxk1, b[1]);
    T bx2 = tfma(-bv, f.sxv2, b[2]);
    T t1 = cm.up(bk * f.fk, 1), t2 = cm.up(bv * f.fv, 1);
    if (lane > 0) { bx1 -= t1; bx2 -= t2; }
#pragma unroll
    for (int lev = 0; lev < NLEV; ++lev) {
        const int sft = 1 << lev;
        const T u0 = cm.up(bx0, sft), u1 = cm.up(bx1, sft), u2 = cm.up(bx2, sft);
        const T d0 = cm.dn(bx0, sft), d1 = cm.dn(bx1, sft), d2 = cm.dn(bx2, sft);
        T al[9], be[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (lev < RLEV) { al[i] = f.al[lev < RLEV ? lev : 0][i]; be[i] = f.be[lev < RLEV ? lev : 0][i]; }
            else { al[i] = f.fs[((lev - RLEV) * 18 + i) * W + lane]; be[i] = f.fs[((lev - RLEV) * 18 + 9 + i) * W + lane]; }
        }
        // two independent FMA chains per row (up / down neighbours) keep the dependent depth at 3
        const T p0 = tfma(-al[2], u2, tfma(-al[1], u1, tfma(-al[0], u0, bx0)));
        const T q0 = tfma(be[2], d2, tfma(be[1], d1, be[0] * d0));
        const T p1 = tfma(-al[5], u2, tfma(-al[4], u1, tfma(-al[3], u0, bx1)));
        const T q1 = tfma(be[5], d2, tfma(be[4], d1, be[3] * d0));
        const T p2 = tfma(-al[8], u2, tfma(-al[7], u1, tfma(-al[6], u0, bx2)));
        const T q2 = tfma(be[8], d2, tfma(be[7], d1, be[6] * d0));
        bx0 = p0 - q0; bx1 = p1 - q1; bx2 = p2 - q2;
    }
    x[0] = tfma(f.Dinv[2], bx2, tfma(f.Dinv[1], bx1, f.Dinv[0] * bx0));
    x[1] = tfma(f.Dinv[4], bx2, tfma(f.Dinv[3], bx1, f.Dinv[1] * bx0));
    x[2] = tfma(f.Dinv[5], bx2, tfma(f.Dinv[4], bx1, f.Dinv[2] * bx0));
    const T xn1 = cm.dn(x[1], 1), xn2 = cm.dn(x[2], 1);  // fv, fk are 0 where there is no successor
    x[3] = f.iv * tfma(-f.fv, xn2, tfma(-f.sxv2, x[2], tfma(-f.sxv0, x[0], b[3])));
    x[4] = f.ik * tfma(-f.fk, xn1, tfma(-f.sxk1, x[1], tfma(-f.sxk0, x[0], b[4])));
}

template <typename T> __device__ __forceinline__ void set_rho(const Stage<T>& s, T rho, T& rd, T rb[5], T rbi[5]) {
    rd = T(kRhoEqOverIneq) * rho;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        rb[i] = s.ctype[i] < 0 ? T(kRhoMin) : (s.ctype[i] > 0 ? T(kRhoEqOverIneq) * rho : rho);
        rbi[i] = T(1) / rb[i];
    }
}

struct SolveResult {
    int iters, status;
};

// The OSQP loop for one scenario (all 32 lanes of the warp call this together), in "increment form":
// instead of x~ the linear solve returns D = x~ - x,
//     S D = -(q + P x + t + A'(R r)),   r = A x - z and t = A'y both TRACKED (r += ..., t += A'dy), never recomputed,
// and the row updates use  v - z = alpha (r + A D),  z+ = clip(z + (v - z) + y / rho),
// dy = rho ((v - z) - (z+ - z)),  r+ = r + alpha A D - (z+ - z).  This is the same iteration as
// oracle/osqp_oracle.c in exact arithmetic (tools/admm_pcr_model.py checks it), but every quantity
// that multiplies rho_eq = 1e3 rho is a small residual rather than a difference of O(1) numbers, which
// is what lets the fp32 path reproduce OSQP's iteration counts and infeasibility certificates.
// On return w[5] holds the UNSCALED primal stage vector (NaN when OSQP would return no solution).
// sm: this warp's shared slab, smem_rows_per_warp<NLEV, RLEV>() * 32 elements of T.
template <typename T, int NLEV, int RLEV, typename Comm>
__device__ __forceinline__ SolveResult admm_solve(Comm& cm, Stage<T>& s, const AdmmSettings& st, int lane, int nstage,
                                                  int nvar, T* sm, T w[5]) {
    constexpr int W = Comm::W;
    if (st.scaling > 0) ruiz_scale(cm, s, st.scaling, nvar);
    else {
#pragma unroll
        for (int i = 0; i < 5; ++i) { s.D[i] = T(1); s.Eb[i] = T(1); }
#pragma unroll
        for (int i = 0; i < 3; ++i) s.Ed[i] = T(1);
        s.cs = T(1);
    }
    const T thr = T(kOsqpInfty * kMinScaling);
#pragma unroll
    for (int i = 0; i < 5; ++i)
        s.ctype[i] = (s.lo[i] < -thr && s.hi[i] > thr) ? -1 : ((s.hi[i] - s.lo[i] < T(kRhoTol)) ? 1 : 0);
    // constants that are only needed at checks go to shared memory
#pragma unroll
    for (int i = 0; i < 3; ++i) { sm[i * W + lane] = s.d[i]; sm[(8 + i) * W + lane] = s.Ed[i]; sm[(21 + i) * W + lane] = T(1) / s.Ed[i]; }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        sm[(3 + i) * W + lane] = s.D[i]; sm[(11 + i) * W + lane] = s.Eb[i];
        sm[(16 + i) * W + lane] = T(1) / s.D[i]; sm[(24 + i) * W + lane] = T(1) / s.Eb[i];
    }
    T rho = T(st.rho), rd, rb[5], rbi[5];
    const T sigma = T(st.sigma), alpha = T(st.alpha);
    set_rho(s, rho, rd, rb, rbi);
    Factor<T, NLEV, RLEV> f;
    f.fs = sm + kConstRows * W;
    factorize<T, NLEV, RLEV>(cm, s, f, sigma, rd, rb, lane, nstage);
    // constant norms
    T nq_s = T(0), nq_u = T(0);
#pragma unroll
    for (int i = 0; i < 5; ++i) { nq_s = tmax(nq_s, tabs(s.q[i])); nq_u = tmax(nq_u, tabs(s.q[i]) * sm[(16 + i) * W + lane]); }
    nq_s = cm.max(nq_s); nq_u = cm.max(nq_u);
    const T cinv = T(1) / s.cs;
    // the loop keeps only a, c, e, P, q, lo, hi of the stage live
    T x[5] = {0, 0, 0, 0, 0}, zb[5] = {0, 0, 0, 0, 0}, yb[5] = {0, 0, 0, 0, 0};
    T ty[5] = {0, 0, 0, 0, 0};  // tracked A'y: accumulated from the small dual steps, never recomputed from a large y
    T rdy[3] = {0, 0, 0}, rbd[5] = {0, 0, 0, 0, 0};  // tracked residuals A x - z
    T stepd[3] = {s.d[0], s.d[1], s.d[2]};            // z jumps from the cold start 0 to d in iteration 1, then stays
    int status = 0, iter = 0;
    int chk = st.check_termination > 0 ? st.check_termination : -1;
    int adp = st.adaptive_rho_interval > 0 ? st.adaptive_rho_interval : -1;
    // The termination check / rho adaptation after a pass.  After max_iter passes OSQP (osqp.c, after its main loop) runs a
    // NORMAL termination check if the last pass was not a check pass (phase 1) and then the APPROXIMATE one (phase 2: every
    // tolerance x 10, statuses 2 / 3 / 4), else reports max-iter (-2): the same code, instantiated once more behind the loop
    // so that the loop itself carries nothing for it.  Returns the OSQP status, 0 = keep iterating.
    auto check = [&](const int phase, const T tol, const T* dl, const T* ed, const T* eb, const bool can_check,
                     const bool can_adapt) __attribute__((always_inline)) -> int {
        if (can_check || can_adapt) {
            T axd[3], axb[5], aty[5], zd[3], D[5], Ed[3], Eb[5], Di[5], Edi[3], Ebi[5], dx[5], dyd[3], dyb[5];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                zd[i] = sm[i * W + lane]; Ed[i] = sm[(8 + i) * W + lane]; Edi[i] = sm[(21 + i) * W + lane];
                dyd[i] = ed[i];
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                D[i] = sm[(3 + i) * W + lane]; Eb[i] = sm[(11 + i) * W + lane];
                Di[i] = sm[(16 + i) * W + lane]; Ebi[i] = sm[(24 + i) * W + lane];
                dx[i] = alpha * dl[i]; dyb[i] = eb[i]; aty[i] = ty[i];
            }
            A_apply(cm, s, x, lane, axd, axb);
            // scaled and unscaled infinity norms
            T pr_s = 0, pr_u = 0, nz_s = 0, nz_u = 0, nax_s = 0, nax_u = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                T r = tabs(axd[i] - zd[i]), ei = Edi[i];
                pr_s = tmax(pr_s, r); pr_u = tmax(pr_u, r * ei);
                nz_s = tmax(nz_s, tabs(zd[i])); nz_u = tmax(nz_u, tabs(zd[i]) * ei);
                nax_s = tmax(nax_s, tabs(axd[i])); nax_u = tmax(nax_u, tabs(axd[i]) * ei);
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                T r = tabs(axb[i] - zb[i]), ei = Ebi[i];
                pr_s = tmax(pr_s, r); pr_u = tmax(pr_u, r * ei);
                nz_s = tmax(nz_s, tabs(zb[i])); nz_u = tmax(nz_u, tabs(zb[i]) * ei);
                nax_s = tmax(nax_s, tabs(axb[i])); nax_u = tmax(nax_u, tabs(axb[i]) * ei);
            }
            T du_s = 0, du_u = 0, npx_s = 0, npx_u = 0, naty_s = 0, naty_u = 0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                T px = s.P[i] * x[i], di = Di[i];
                T r = tabs(px + s.q[i] + aty[i]);
                du_s = tmax(du_s, r); du_u = tmax(du_u, r * di);
                npx_s = tmax(npx_s, tabs(px)); npx_u = tmax(npx_u, tabs(px) * di);
                naty_s = tmax(naty_s, tabs(aty[i])); naty_u = tmax(naty_u, tabs(aty[i]) * di);
            }
            pr_s = cm.max(pr_s); pr_u = cm.max(pr_u); du_s = cm.max(du_s); du_u = cm.max(du_u) * cinv;
            nz_s = cm.max(nz_s); nz_u = cm.max(nz_u); nax_s = cm.max(nax_s); nax_u = cm.max(nax_u);
            npx_s = cm.max(npx_s); npx_u = cm.max(npx_u); naty_s = cm.max(naty_s); naty_u = cm.max(naty_u);
            if (can_check) {
                if (pr_u > T(kOsqpInfty) || du_u > T(kOsqpInfty)) return -7;
                const T eps_prim = tol * (T(st.eps_abs) + T(st.eps_rel) * tmax(nz_u, nax_u));
                const T eps_dual = tol * (T(st.eps_abs) + T(st.eps_rel) * cinv * tmax(tmax(nq_u, naty_u), npx_u));
                const bool prim_ok = pr_u < eps_prim, dual_ok = du_u < eps_dual;
                if (prim_ok && dual_ok) return phase == 2 ? 2 : 1;
                bool pinf = false, dinf = false;
                if (!prim_ok) {  // is_primal_infeasible
                    const T epi = tol * T(st.eps_prim_inf);
                    T pyb[5], ndy = 0, lhs = 0;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        ndy = tmax(ndy, tabs(Ed[i] * dyd[i]));
                        lhs += zd[i] * dyd[i];  // u*max(dy,0) + l*min(dy,0) with l = u = d
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        T d = dyb[i];
                        if (s.hi[i] > thr) d = (s.lo[i] < -thr) ? T(0) : tmin(d, T(0));
                        else if (s.lo[i] < -thr) d = tmax(d, T(0));
                        pyb[i] = d;
                        ndy = tmax(ndy, tabs(Eb[i] * d));
                        lhs += s.hi[i] * tmax(d, T(0)) + s.lo[i] * tmin(d, T(0));
                    }
                    ndy = cm.max(ndy);
                    lhs = cm.sum(lhs);
                    if (ndy > epi && lhs < -epi * ndy) {
                        T atdy[5], na = 0;
                        At_apply(cm, s, dyd, pyb, atdy);
#pragma unroll
                        for (int i = 0; i < 5; ++i) na = tmax(na, tabs(atdy[i] * Di[i]));
                        na = cm.max(na);
                        pinf = na < epi * ndy;
                    }
                }
                if (!dual_ok && !pinf) {  // is_dual_infeasible
                    const T edi = tol * T(st.eps_dual_inf);
                    T ndx = 0, qdx = 0;
#pragma unroll
                    for (int i = 0; i < 5; ++i) { ndx = tmax(ndx, tabs(D[i] * dx[i])); qdx += s.q[i] * dx[i]; }
                    ndx = cm.max(ndx);
                    qdx = cm.sum(qdx);
                    if (ndx > edi && qdx < -s.cs * edi * ndx) {
                        T npdx = 0;
#pragma unroll
                        for (int i = 0; i < 5; ++i) npdx = tmax(npdx, tabs(s.P[i] * dx[i] * Di[i]));
                        npdx = cm.max(npdx);
                        if (npdx < s.cs * edi * ndx) {
                            T adxd[3], adxb[5];
                            A_apply(cm, s, dx, lane, adxd, adxb);
                            int bad = 0;
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                T v = adxd[i] * Edi[i];
                                if (v > edi * ndx || v < -edi * ndx) bad = 1;  // equality rows have finite bounds
                            }
#pragma unroll
                            for (int i = 0; i < 5; ++i) {
                                T v = adxb[i] * Ebi[i];
                                if ((s.hi[i] < thr && v > edi * ndx) || (s.lo[i] > -thr && v < -edi * ndx)) bad = 1;
                            }
                            dinf = !cm.any(bad != 0);
                        }
                    }
                }
                if (pinf) return phase == 2 ? 3 : -3;
                if (dinf) return phase == 2 ? 4 : -4;
            }
            if (can_adapt) {  // adapt_rho / compute_rho_estimate on the scaled residuals
                T pn = pr_s / (tmax(nz_s, nax_s) + T(1e-10));
                T dn = du_s / (tmax(tmax(nq_s, naty_s), npx_s) + T(1e-10));
                T rnew = rho * sqrt(pn / (dn + T(1e-10)));
                rnew = tmin(tmax(rnew, T(kRhoMin)), T(kRhoMax));
                if (rnew > rho * T(st.adaptive_rho_tolerance) || rnew < rho / T(st.adaptive_rho_tolerance)) {
                    rho = rnew;
                    set_rho(s, rho, rd, rb, rbi);
                    factorize<T, NLEV, RLEV>(cm, s, f, sigma, rd, rb, lane, nstage);
                }
            }
        }
        return 0;
    };
    // one ADMM pass; leaves this pass's steps dl (primal), ed / eb (dual) for the checks
    auto pass = [&](T* dl, T* ed, T* eb) __attribute__((always_inline)) {
        T g[5], td[3], tb[5];
#pragma unroll
        for (int i = 0; i < 3; ++i) td[i] = rd * rdy[i];
#pragma unroll
        for (int i = 0; i < 5; ++i) tb[i] = rb[i] * rbd[i];
        At_apply(cm, s, td, tb, g);
#pragma unroll
        for (int i = 0; i < 5; ++i) g[i] = -((tfma(s.P[i], x[i], s.q[i]) + ty[i]) + g[i]);
        kkt_solve<T, NLEV, RLEV>(cm, f, g, lane, dl);
        T add[3], adb[5];
        A_apply(cm, s, dl, lane, add, adb);
        // ed, eb: dual steps dy = rho ((v - z_prev) - (z_new - z_prev))
#pragma unroll
        for (int i = 0; i < 5; ++i) x[i] = tfma(alpha, dl[i], x[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const T wv = alpha * (rdy[i] + add[i]);  // v - z_prev
            ed[i] = rd * (wv - stepd[i]);  // dy of the dynamics rows (their y itself is never needed)
            rdy[i] = tfma(alpha, add[i], rdy[i]) - stepd[i];
            stepd[i] = T(0);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const T wv = alpha * (rbd[i] + adb[i]);
            const T zn = tmin(tmax(tfma(rbi[i], yb[i], zb[i] + wv), s.lo[i]), s.hi[i]);
            const T step = zn - zb[i];
            eb[i] = rb[i] * (wv - step);  // dy of the bound rows
            yb[i] += eb[i];
            rbd[i] = tfma(alpha, adb[i], rbd[i]) - step;
            zb[i] = zn;
        }
        {
            T dty[5];
            At_apply(cm, s, ed, eb, dty);
#pragma unroll
            for (int i = 0; i < 5; ++i) ty[i] += dty[i];
        }
    };
    for (iter = 1; iter <= st.max_iter; ++iter) {
        T dl[5], ed[3], eb[5];
        // Passes after which nothing happens (no check, no rho adaptation, not the last one) run in a loop of their own: one
        // straight-line body and one backward branch instead of a round trip through the check code's branches.  Same
        // passes in the same order; the check lambda would only have decremented the two counters.
        {
            int quiet = st.max_iter - iter;
            if (chk > 0) quiet = min(quiet, chk - 1);
            if (adp > 0) quiet = min(quiet, adp - 1);
#pragma unroll 1
            for (int i = 0; i < quiet; ++i) pass(dl, ed, eb);
            if (quiet > 0) {
                iter += quiet;
                if (chk > 0) chk -= quiet;
                if (adp > 0) adp -= quiet;
            }
        }
        pass(dl, ed, eb);
        const bool can_check = (--chk == 0), can_adapt = (--adp == 0);
        if (can_check) chk = st.check_termination;
        if (can_adapt) adp = st.adaptive_rho_interval;
        status = check(0, T(1), dl, ed, eb, can_check, can_adapt);
        if (status != 0) break;
        if (iter == st.max_iter) {  // the end-of-loop checks need this pass's steps
#pragma unroll
            for (int i = 0; i < 5; ++i) { sm[(29 + i) * W + lane] = dl[i]; sm[(37 + i) * W + lane] = eb[i]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) sm[(34 + i) * W + lane] = ed[i];
        }
    }
    if (status == 0) {
        iter = st.max_iter;
        T dl[5], ed[3], eb[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) { dl[i] = sm[(29 + i) * W + lane]; eb[i] = sm[(37 + i) * W + lane]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) ed[i] = sm[(34 + i) * W + lane];
        // the last pass was a check pass iff check_termination divides max_iter
        const bool checked = st.check_termination > 0 && (st.max_iter % st.check_termination == 0);
        if (!checked) status = check(1, T(1), dl, ed, eb, true, false);
        if (status == 0) status = check(2, T(10), dl, ed, eb, true, false);
        if (status == 0) status = -2;
    }
    T D[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) D[i] = sm[(3 + i) * W + lane];
    // OSQP stores no solution for the (approximately) infeasible and the non-convex statuses (NaN vectors)
    const bool nan_out = (status == -3 || status == -4 || status == -7 || status == 3 || status == 4);
#pragma unroll
    for (int i = 0; i < 5; ++i) w[i] = nan_out ? T(NAN) : D[i] * x[i];
    SolveResult r;
    r.iters = iter;
    r.status = status;
    return r;
}

// ------------------------------------------------------------------------------------------------
// stage loaders
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void stage_zero(Stage<T>& s) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s.a[i] = T(0);
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.c[i] = T(0); s.d[i] = T(0); }
#pragma unroll
    for (int i = 0; i < 5; ++i) { s.e[i] = T(0); s.P[i] = T(0); s.q[i] = T(0); s.lo[i] = T(0); s.hi[i] = T(0); }
}

__device__ __forceinline__ double clip_inf(double v) { return fmin(fmax(v, -kOsqpInfty), kOsqpInfty); }

// K2-only entry: the QP arrives in the reference's layout (see mpc_b200.h::mpc_solve_qp).
// Offsets into the fixed CSC pattern: stage k < N occupies 16 values (x columns: 5, 4, 3; then the
// inputs live at the tail: 2 values per input column); stage N occupies 6.
template <typename T>
__device__ __forceinline__ void load_stage_qp(Stage<T>& s, int N, int lane, const double* Pd, const double* q,
                                              const double* Ax, const double* l, const double* u) {
    stage_zero(s);
    if (lane > N) return;
    const int neq = 3 * (N + 1);
    const int k = lane;
    if (k < N) {
        const double* v = Ax + 12 * k;  // x columns of stage k: [c0 a0 a2 a4 e0 | c1 a1 a3 e1 | c2 a5 e2]
        s.c[0] = T(v[0]); s.a[0] = T(v[1]); s.a[2] = T(v[2]); s.a[4] = T(v[3]); s.e[0] = T(v[4]);
        s.c[1] = T(v[5]); s.a[1] = T(v[6]); s.a[3] = T(v[7]); s.e[1] = T(v[8]);
        s.c[2] = T(v[9]); s.a[5] = T(v[10]); s.e[2] = T(v[11]);
        const double* w = Ax + 12 * N + 6 + 4 * k;  // input columns: [a7 e3 | a6 e4]
        s.a[7] = T(w[0]); s.e[3] = T(w[1]); s.a[6] = T(w[2]); s.e[4] = T(w[3]);
    } else {
        const double* v = Ax + 12 * N;  // last stage: [c0 e0 | c1 e1 | c2 e2]
        s.c[0] = T(v[0]); s.e[0] = T(v[1]); s.c[1] = T(v[2]); s.e[1] = T(v[3]); s.c[2] = T(v[4]); s.e[2] = T(v[5]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        s.P[i] = T(Pd[3 * k + i]); s.q[i] = T(q[3 * k + i]);
        s.d[i] = T(l[3 * k + i]);
        s.lo[i] = T(clip_inf(l[neq + 3 * k + i])); s.hi[i] = T(clip_inf(u[neq + 3 * k + i]));
    }
    if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            s.P[3 + i] = T(Pd[neq + 2 * k + i]); s.q[3 + i] = T(q[neq + 2 * k + i]);
            s.lo[3 + i] = T(clip_inf(l[2 * neq + 2 * k + i])); s.hi[3 + i] = T(clip_inf(u[2 * neq + 2 * k + i]));
        }
    }
}

// K1: MPC._init_problem for stage `lane` (MPC.py:86-155, sbm.py:404-412), fp64 then cast.
template <typename T>
__device__ __forceinline__ void assemble_stage(Stage<T>& s, const MpcParams& mp, const PathView& pv, int lane, int wp_id,
                                               double e_y, double e_psi, const double* cc /*2N*/, const double* ub,
                                               const double* lb) {
    stage_zero(s);
    const int N = mp.N;
    if (lane > N) return;
    const int k = lane;
#pragma unroll
    for (int i = 0; i < 3; ++i) s.c[i] = T(-1);
#pragma unroll
    for (int i = 0; i < 3; ++i) s.e[i] = T(1);
    if (k < N) {
        int w0 = wp_id + k;
        if (w0 >= pv.n_wp) w0 %= pv.n_wp;  // rp.py:364-365 (non-circular end-of-path is flagged by the caller)
        const double2* row = reinterpret_cast<const double2*>(pv.stage_tab) + (size_t)w0 * (kStageTab / 2);
        const double2 t01 = row[0], t23 = row[1], t45 = row[2];
        s.a[0] = T(1); s.a[1] = T(t01.x);
        s.a[2] = T(t01.y); s.a[3] = T(1); s.a[6] = T(t01.x);
        s.a[4] = T(t23.x); s.a[5] = T(1); s.a[7] = T(t23.y);
        s.e[3] = T(1); s.e[4] = T(1);
        s.P[0] = T(mp.Q[0]); s.P[1] = T(mp.Q[1]); s.P[2] = T(mp.Q[2]); s.P[3] = T(mp.R[0]); s.P[4] = T(mp.R[1]);
        s.q[3] = T(t45.x); s.q[4] = T(t45.y);
        // input bounds; speed limit from predicted curvature (MPC.py:86-87,111-113; quirk Q1)
        const double kp = tan(cc[3 + k] + cc[2 * N - 1]) / mp.L;
        const double vmax_dyn = sqrt(mp.ay_max / (fabs(kp) + 1e-12));
        s.lo[3] = T(clip_inf(mp.umin[0])); s.hi[3] = T(clip_inf(fmin(mp.umax[0], vmax_dyn)));
        s.lo[4] = T(clip_inf(mp.umin[1])); s.hi[4] = T(clip_inf(mp.umax[1]));
    } else {
        s.P[0] = T(mp.QN[0]); s.P[1] = T(mp.QN[1]); s.P[2] = T(mp.QN[2]);
    }
    if (k == 0) {
        s.d[0] = T(-e_y); s.d[1] = T(-e_psi); s.d[2] = T(-0.0);  // leq = -x0 (MPC.py:142-143)
        s.lo[0] = T(e_y); s.hi[0] = T(e_y);                      // MPC.py:119-120 (quirk Q6)
        s.q[0] = T(-mp.Q[0] * 0.0);
    } else {
        int wm = wp_id + k - 1;
        if (wm >= pv.n_wp) wm %= pv.n_wp;
        // uq = B_lin.dot([v_ref, kappa_ref]) - f  (MPC.py:107-108)
        const double2 t67 = reinterpret_cast<const double2*>(pv.stage_tab)[(size_t)wm * (kStageTab / 2) + 3];
        s.d[0] = T(0);
        s.d[1] = T(t67.x);
        s.d[2] = T(t67.y);
        const double l_ = lb[k - 1], u_ = ub[k - 1];
        s.lo[0] = T(clip_inf(l_)); s.hi[0] = T(clip_inf(u_));
        const double xr = (l_ + u_) / 2;  // MPC.py:125
        s.q[0] = T(-(k < N ? mp.Q[0] : mp.QN[0]) * xr);
    }
    s.lo[1] = T(clip_inf(mp.xmin[1])); s.hi[1] = T(clip_inf(mp.xmax[1]));
    s.lo[2] = T(clip_inf(mp.xmin[2])); s.hi[2] = T(clip_inf(mp.xmax[2]));
}

}  // namespace mpcb
