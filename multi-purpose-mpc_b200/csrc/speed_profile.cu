// speed_profile.cu -- ReferencePath.compute_speed_profile (reference: src/reference_path.py:289-354)
// on the device: the one-off QP
//     min 1/2 ||v||^2 - v_max' v   s.t.  a_min <= D1 v <= a_max,   v_min <= v <= v_max_dyn
// with D1[i] = [-1/(2 l_i), 1/(2 l_i)] on (v_i, v_{i+1}), solved by the same OSQP iteration as the MPC
// QP (oracle/osqp_oracle.c).  n = n_waypoints - 1 variables (199 on the sim track); the reduced KKT
// matrix P + sigma I + A' diag(rho) A is scalar tridiagonal.  One CTA: vector updates are
// thread-parallel, the tridiagonal LDL' sweep is done by one thread (the problem is solved once per
// track, so latency is irrelevant; what matters is that the iterates equal the reference solver's in
// fp64, because OSQP's default eps = 1e-3 answer -- not the exact minimiser -- becomes v_ref).
#include "engine.h"

#include <string>
#include <vector>

namespace mpcb {

struct SpShared {
    double *P, *q, *av, *aw, *eb;      // P diag; accel row i: av[i]*v_i + aw[i]*v_{i+1}; bound row coef
    double *la, *ua, *lb, *ub;         // bounds
    double *D, *Ea, *Eb;               // scalings
    double *x, *za, *zb, *ya, *yb, *xt, *rhs, *dx, *dya, *dyb;
    double *dg, *od, *ld;              // tridiagonal: diagonal, off-diagonal, LDL' multipliers
    double *ra, *rb;                   // rho per row
    double *tmp;                       // reduction scratch [blockDim]
    int *ta, *tb;                      // constraint types
};

__device__ double block_max(double v, double* tmp) {
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) tmp[threadIdx.x] = fmax(tmp[threadIdx.x], tmp[threadIdx.x + s]);
        __syncthreads();
    }
    const double r = tmp[0];
    __syncthreads();
    return r;
}
__device__ double block_sum(double v, double* tmp) {
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) tmp[threadIdx.x] += tmp[threadIdx.x + s];
        __syncthreads();
    }
    const double r = tmp[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double lim_scal(double v) {
    v = v < kMinScaling ? 1.0 : v;
    return v > kMaxScaling ? kMaxScaling : v;
}

// One CTA per track (blockIdx.x): the batched form serves scenario sets that randomise the track itself
// (SURVEY 8f-1).  off[t] .. off[t+1] are track t's waypoints in the concatenated li / vmax / v_out arrays (li holds
// n - 1 segment lengths per track, its last slot is unused).  The 27 n working doubles live in shared memory when they
// fit and otherwise in this CTA's slice of the global workspace `ws` (L2-resident: a long track just runs slower; the
// reference's compute_speed_profile has no size limit, rp.py:289-354).
__global__ void speed_profile_kernel(const int* __restrict__ off, const double* __restrict__ li_all,
                                     const double* __restrict__ vmax_all, double v_min, double a_min, double a_max,
                                     AdmmSettings st, double* __restrict__ v_out_all, int* __restrict__ info_all /*[T][2]*/,
                                     double* __restrict__ ws, size_t ws_stride) {
    extern __shared__ double sm_dyn[];
    __shared__ double tmp_red[256];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n = off[blockIdx.x + 1] - off[blockIdx.x], na = n - 1;
    const double* li = li_all + off[blockIdx.x];
    const double* vmax = vmax_all + off[blockIdx.x];
    double* v_out = v_out_all + off[blockIdx.x];
    int* info = info_all + 2 * blockIdx.x;
    double* sm = ws ? ws + (size_t)blockIdx.x * ws_stride : sm_dyn;
    SpShared S;
    double* p = sm;
    auto take = [&](int k) { double* r = p; p += k; return r; };
    S.P = take(n); S.q = take(n); S.av = take(n); S.aw = take(n); S.eb = take(n);
    S.la = take(n); S.ua = take(n); S.lb = take(n); S.ub = take(n);
    S.D = take(n); S.Ea = take(n); S.Eb = take(n);
    S.x = take(n); S.za = take(n); S.zb = take(n); S.ya = take(n); S.yb = take(n); S.xt = take(n); S.rhs = take(n);
    S.dx = take(n); S.dya = take(n); S.dyb = take(n);
    S.dg = take(n); S.od = take(n); S.ld = take(n); S.ra = take(n); S.rb = take(n);
    S.tmp = tmp_red;
    S.ta = reinterpret_cast<int*>(take((n + 1) / 2 + 1));
    S.tb = reinterpret_cast<int*>(take((n + 1) / 2 + 1));
    // ---- problem data (rp.py:300-344) ----
    for (int i = tid; i < n; i += nt) {
        S.P[i] = 1.0;
        S.q[i] = -1 * vmax[i];
        S.eb[i] = 1.0;
        S.lb[i] = fmax(v_min, -kOsqpInfty); S.ub[i] = fmin(vmax[i], kOsqpInfty);
        if (i < na) {
            S.av[i] = -1 / (2 * li[i]); S.aw[i] = 1 / (2 * li[i]);
            S.la[i] = fmax(a_min, -kOsqpInfty); S.ua[i] = fmin(a_max, kOsqpInfty);
        } else { S.av[i] = 0; S.aw[i] = 0; S.la[i] = 0; S.ua[i] = 0; }
        S.D[i] = 1.0; S.Ea[i] = 1.0; S.Eb[i] = 1.0;
        S.x[i] = 0; S.za[i] = 0; S.zb[i] = 0; S.ya[i] = 0; S.yb[i] = 0;
    }
    __syncthreads();
    // ---- Ruiz equilibration + cost scaling ----
    double cs = 1.0;
    for (int it = 0; it < st.scaling; ++it) {
        double Dt[1], Eat[1], Ebt[1];
        // one variable / row pair per thread-iteration; results staged in xt / rhs / dx as scratch
        for (int i = tid; i < n; i += nt) {
            double col = fmax(fabs(S.P[i]), fabs(S.eb[i]));
            if (i < na) col = fmax(col, fabs(S.av[i]));
            if (i > 0) col = fmax(col, fabs(S.aw[i - 1]));
            S.xt[i] = 1.0 / sqrt(lim_scal(col));
            S.rhs[i] = 1.0 / sqrt(lim_scal(i < na ? fmax(fabs(S.av[i]), fabs(S.aw[i])) : 0.0));
            S.dx[i] = 1.0 / sqrt(lim_scal(fabs(S.eb[i])));
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            Dt[0] = S.xt[i]; Eat[0] = S.rhs[i]; Ebt[0] = S.dx[i];
            S.P[i] *= Dt[0] * Dt[0];
            S.q[i] *= Dt[0];
            S.eb[i] *= Ebt[0] * Dt[0];
            if (i < na) { S.av[i] *= Eat[0] * Dt[0]; S.aw[i] *= Eat[0] * S.xt[i + 1]; }
            S.D[i] *= Dt[0]; S.Eb[i] *= Ebt[0];
            if (i < na) S.Ea[i] *= Eat[0];
        }
        __syncthreads();
        double sp = 0, mq = 0;
        for (int i = tid; i < n; i += nt) { sp += fabs(S.P[i]); mq = fmax(mq, fabs(S.q[i])); }
        sp = block_sum(sp, S.tmp) / n;
        mq = lim_scal(block_max(mq, S.tmp));
        double ct = 1.0 / lim_scal(fmax(sp, mq));
        for (int i = tid; i < n; i += nt) { S.P[i] *= ct; S.q[i] *= ct; }
        cs *= ct;
        __syncthreads();
    }
    const double cinv = 1.0 / cs, thr = kOsqpInfty * kMinScaling;
    for (int i = tid; i < n; i += nt) {
        S.lb[i] *= S.Eb[i]; S.ub[i] *= S.Eb[i];
        if (i < na) { S.la[i] *= S.Ea[i]; S.ua[i] *= S.Ea[i]; }
        S.tb[i] = (S.lb[i] < -thr && S.ub[i] > thr) ? -1 : (S.ub[i] - S.lb[i] < kRhoTol ? 1 : 0);
        S.ta[i] = i < na ? ((S.la[i] < -thr && S.ua[i] > thr) ? -1 : (S.ua[i] - S.la[i] < kRhoTol ? 1 : 0)) : 0;
    }
    __syncthreads();
    double rho = st.rho;
    const double sigma = st.sigma, alpha = st.alpha;
    double nq_s = 0, nq_u = 0;
    for (int i = tid; i < n; i += nt) { nq_s = fmax(nq_s, fabs(S.q[i])); nq_u = fmax(nq_u, fabs(S.q[i] / S.D[i])); }
    nq_s = block_max(nq_s, S.tmp); nq_u = block_max(nq_u, S.tmp);

    auto factorize = [&]() {
        for (int i = tid; i < n; i += nt) {
            S.rb[i] = S.tb[i] < 0 ? kRhoMin : (S.tb[i] > 0 ? kRhoEqOverIneq * rho : rho);
            S.ra[i] = S.ta[i] < 0 ? kRhoMin : (S.ta[i] > 0 ? kRhoEqOverIneq * rho : rho);
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            double d = S.P[i] + sigma + S.rb[i] * S.eb[i] * S.eb[i];
            if (i < na) d += S.ra[i] * S.av[i] * S.av[i];
            if (i > 0) d += S.ra[i - 1] * S.aw[i - 1] * S.aw[i - 1];
            S.dg[i] = d;
            S.od[i] = i < na ? S.ra[i] * S.av[i] * S.aw[i] : 0.0;
        }
        __syncthreads();
        if (tid == 0) {  // LDL' of the tridiagonal: dg <- pivots, ld <- multipliers
            for (int i = 0; i < n - 1; ++i) {
                S.ld[i] = S.od[i] / S.dg[i];
                S.dg[i + 1] -= S.ld[i] * S.od[i];
            }
        }
        __syncthreads();
    };
    factorize();
    int status = 0, iter = 0;
    for (iter = 1; iter <= st.max_iter; ++iter) {
        for (int i = tid; i < n; i += nt) {
            double r = sigma * S.x[i] - S.q[i] + S.eb[i] * (S.rb[i] * S.zb[i] - S.yb[i]);
            if (i < na) r += S.av[i] * (S.ra[i] * S.za[i] - S.ya[i]);
            if (i > 0) r += S.aw[i - 1] * (S.ra[i - 1] * S.za[i - 1] - S.ya[i - 1]);
            S.rhs[i] = r;
        }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < n; ++i) S.rhs[i] -= S.ld[i - 1] * S.rhs[i - 1];
            S.xt[n - 1] = S.rhs[n - 1] / S.dg[n - 1];
            for (int i = n - 2; i >= 0; --i) S.xt[i] = S.rhs[i] / S.dg[i] - S.ld[i] * S.xt[i + 1];
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const double xn = alpha * S.xt[i] + (1 - alpha) * S.x[i];
            S.dx[i] = xn - S.x[i];
            // bound row
            const double ztb = S.eb[i] * S.xt[i];
            const double vb = alpha * ztb + (1 - alpha) * S.zb[i];
            const double znb = fmin(fmax(vb + S.yb[i] / S.rb[i], S.lb[i]), S.ub[i]);
            S.dyb[i] = S.rb[i] * (vb - znb);
            S.yb[i] += S.dyb[i];
            S.zb[i] = znb;
            if (i < na) {
                const double zta = S.av[i] * S.xt[i] + S.aw[i] * S.xt[i + 1];
                const double va = alpha * zta + (1 - alpha) * S.za[i];
                const double zna = fmin(fmax(va + S.ya[i] / S.ra[i], S.la[i]), S.ua[i]);
                S.dya[i] = S.ra[i] * (va - zna);
                S.ya[i] += S.dya[i];
                S.za[i] = zna;
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) S.x[i] += S.dx[i];
        __syncthreads();
        const bool can_check = st.check_termination && (iter % st.check_termination == 0);
        const bool can_adapt = st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0);
        if (can_check || can_adapt) {
            double pr_s = 0, pr_u = 0, nz_s = 0, nz_u = 0, nax_s = 0, nax_u = 0, du_s = 0, du_u = 0, npx_s = 0, npx_u = 0,
                   naty_s = 0, naty_u = 0;
            for (int i = tid; i < n; i += nt) {
                const double axb = S.eb[i] * S.x[i], eib = 1.0 / S.Eb[i];
                pr_s = fmax(pr_s, fabs(axb - S.zb[i])); pr_u = fmax(pr_u, fabs(axb - S.zb[i]) * eib);
                nz_s = fmax(nz_s, fabs(S.zb[i])); nz_u = fmax(nz_u, fabs(S.zb[i]) * eib);
                nax_s = fmax(nax_s, fabs(axb)); nax_u = fmax(nax_u, fabs(axb) * eib);
                double aty = S.eb[i] * S.yb[i];
                if (i < na) {
                    const double axa = S.av[i] * S.x[i] + S.aw[i] * S.x[i + 1], eia = 1.0 / S.Ea[i];
                    pr_s = fmax(pr_s, fabs(axa - S.za[i])); pr_u = fmax(pr_u, fabs(axa - S.za[i]) * eia);
                    nz_s = fmax(nz_s, fabs(S.za[i])); nz_u = fmax(nz_u, fabs(S.za[i]) * eia);
                    nax_s = fmax(nax_s, fabs(axa)); nax_u = fmax(nax_u, fabs(axa) * eia);
                    aty += S.av[i] * S.ya[i];
                }
                if (i > 0) aty += S.aw[i - 1] * S.ya[i - 1];
                const double px = S.P[i] * S.x[i], di = 1.0 / S.D[i], r = fabs(px + S.q[i] + aty);
                du_s = fmax(du_s, r); du_u = fmax(du_u, r * di);
                npx_s = fmax(npx_s, fabs(px)); npx_u = fmax(npx_u, fabs(px) * di);
                naty_s = fmax(naty_s, fabs(aty)); naty_u = fmax(naty_u, fabs(aty) * di);
            }
            pr_s = block_max(pr_s, S.tmp); pr_u = block_max(pr_u, S.tmp); nz_s = block_max(nz_s, S.tmp);
            nz_u = block_max(nz_u, S.tmp); nax_s = block_max(nax_s, S.tmp); nax_u = block_max(nax_u, S.tmp);
            du_s = block_max(du_s, S.tmp); du_u = block_max(du_u, S.tmp) * cinv; npx_s = block_max(npx_s, S.tmp);
            npx_u = block_max(npx_u, S.tmp); naty_s = block_max(naty_s, S.tmp); naty_u = block_max(naty_u, S.tmp);
            if (can_check) {
                const double eps_prim = st.eps_abs + st.eps_rel * fmax(nz_u, nax_u);
                const double eps_dual = st.eps_abs + st.eps_rel * cinv * fmax(fmax(nq_u, naty_u), npx_u);
                if (pr_u < eps_prim && du_u < eps_dual) { status = 1; break; }
                // (infeasibility certificates: the speed-profile QP is always feasible for
                //  a_min <= 0 <= a_max and v_min <= v_max; omitted)
            }
            if (can_adapt) {
                const double pn = pr_s / (fmax(nz_s, nax_s) + 1e-10);
                const double dn = du_s / (fmax(fmax(nq_s, naty_s), npx_s) + 1e-10);
                double rnew = rho * sqrt(pn / (dn + 1e-10));
                rnew = fmin(fmax(rnew, kRhoMin), kRhoMax);
                if (rnew > rho * st.adaptive_rho_tolerance || rnew < rho / st.adaptive_rho_tolerance) {
                    rho = rnew;
                    factorize();
                }
            }
        }
    }
    if (status == 0) { status = -2; iter = st.max_iter; }
    for (int i = tid; i < n; i += nt) v_out[i] = S.D[i] * S.x[i];
    if (tid == 0) { info[0] = iter; info[1] = status; }
}

}  // namespace mpcb

using namespace mpcb;

static size_t sp_work_doubles(int n) { return (size_t)27 * n + 2 * ((size_t)(n + 1) / 2 + 1); }

extern "C" int mpc_speed_profile_batch(const double* h_li, const double* h_vmax, const int32_t* h_off, int32_t T,
                                       double v_min, double a_min, double a_max, const mpc_config* cfg, double* h_v_out,
                                       int32_t* h_iters, int32_t* h_status) {
    extern int mpc_set_error_(int code, const char* msg);
    if (!h_li || !h_vmax || !h_off || !h_v_out || T < 1) return mpc_set_error_(MPC_E_INVALID, "bad speed-profile arguments");
    int n_max = 0;
    for (int t = 0; t < T; ++t) {
        const int n = h_off[t + 1] - h_off[t];
        if (n < 2) return mpc_set_error_(MPC_E_INVALID, "a speed profile needs at least 2 waypoints (rp.py:296-297)");
        n_max = n > n_max ? n : n_max;
    }
    if (h_off[0] != 0) return mpc_set_error_(MPC_E_INVALID, "track offsets must start at 0");
    const int total = h_off[T];
    mpc_config c;
    if (cfg) c = *cfg; else mpc_config_default(&c);
    AdmmSettings st;
    st.rho = c.rho; st.sigma = c.sigma; st.alpha = c.alpha; st.eps_abs = c.eps_abs; st.eps_rel = c.eps_rel;
    st.eps_prim_inf = c.eps_prim_inf; st.eps_dual_inf = c.eps_dual_inf; st.adaptive_rho_tolerance = c.adaptive_rho_tolerance;
    st.max_iter = c.max_iter; st.scaling = c.scaling; st.check_termination = c.check_termination;
    st.adaptive_rho_interval = c.adaptive_rho_interval;
    double *d_li = nullptr, *d_vmax = nullptr, *d_v = nullptr, *d_ws = nullptr;
    int *d_info = nullptr, *d_off = nullptr;
    const int nt = 256;
    const size_t work = sp_work_doubles(n_max);
    size_t smem = work * sizeof(double);
    const bool spill = smem > 200 * 1024;  // longer tracks keep their vectors in a global workspace (one slice per CTA)
    if (spill) smem = 0;
    cudaError_t e = cudaMalloc(&d_li, (size_t)total * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_vmax, (size_t)total * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_v, (size_t)total * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_info, 2 * (size_t)T * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_off, ((size_t)T + 1) * sizeof(int));
    if (e == cudaSuccess && spill) e = cudaMalloc(&d_ws, work * (size_t)T * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_li, h_li, (size_t)total * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_vmax, h_vmax, (size_t)total * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_off, h_off, ((size_t)T + 1) * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && smem) e = cudaFuncSetAttribute(speed_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<int> info(2 * (size_t)T, 0);
    if (e == cudaSuccess) {
        speed_profile_kernel<<<T, nt, smem>>>(d_off, d_li, d_vmax, v_min, a_min, a_max, st, d_v, d_info, d_ws, work);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h_v_out, d_v, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(info.data(), d_info, info.size() * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_li); cudaFree(d_vmax); cudaFree(d_v); cudaFree(d_info); cudaFree(d_off); cudaFree(d_ws);
    if (e != cudaSuccess) return mpc_set_error_(MPC_E_CUDA, cudaGetErrorString(e));
    for (int t = 0; t < T; ++t) {
        if (h_iters) h_iters[t] = info[2 * t];
        if (h_status) h_status[t] = info[2 * t + 1];
    }
    return 0;
}

extern "C" int mpc_speed_profile(const double* h_li, const double* h_vmax, int32_t n, double v_min, double a_min,
                                 double a_max, const mpc_config* cfg, double* h_v_out, int32_t* h_iters,
                                 int32_t* h_status) {
    extern int mpc_set_error_(int code, const char* msg);
    if (n < 2) return mpc_set_error_(MPC_E_INVALID, "a speed profile needs at least 2 waypoints (rp.py:296-297)");
    const int32_t off[2] = {0, n};
    // the single-track ABI passes n - 1 segment lengths: pad to n so that both arrays share the offsets
    std::vector<double> li(h_li, h_li + (n - 1));
    li.push_back(0.0);
    return mpc_speed_profile_batch(li.data(), h_vmax, off, 1, v_min, a_min, a_max, cfg, h_v_out, h_iters, h_status);
}
