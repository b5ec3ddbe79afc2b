/*
 * oracle/osqp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU fp64 restatement of the OSQP algorithm (Stellato, Banjac, Goulart, Bemporad, Boyd:
 * "OSQP: an operator splitting solver for quadratic programs", Math. Prog. Comp. 12, 2020) as the
 * reference calls it: osqp.OSQP().setup(P, q, A, l, u, verbose=False); .solve()
 * (/root/reference/src/MPC.py:158-159,183 and reference_path.py:347-349), i.e. OSQP ~0.6.x with
 * default settings:
 *   rho=0.1 sigma=1e-6 alpha=1.6 eps_abs=eps_rel=1e-3 eps_prim_inf=eps_dual_inf=1e-4
 *   max_iter=4000 scaling=10 (Ruiz + cost scaling) adaptive_rho=1 adaptive_rho_tolerance=5
 *   check_termination=25 scaled_termination=0 polish=0, cold start, +-inf clipped to +-1e30.
 *
 * PARITY UNPINNED: the osqp package and its C source are not on this machine and the reference has
 * no tests, so the constants and the order of operations below are restated from the paper and
 * the documented defaults, not checked against an OSQP binary.  What is checked (tests/): the
 * returned point satisfies the QP's KKT conditions in fp64 to the requested tolerance, which
 * certifies the minimiser independently of which solver produced it.
 *
 * Deliberate, documented deviations from an OSQP binary:
 *   - adaptive_rho_interval: OSQP (adaptive_rho_interval=0) derives it from wall-clock timing of
 *     setup vs. iterations, then rounds to a multiple of check_termination.  Here it is a fixed
 *     setting (default 25 = what the timing rule yields for QPs this small).   
 *   - linear system: OSQP factorises the quasi-definite KKT matrix with QDLDL; here the
 *     algebraically identical reduced system (P + sigma I + A' diag(rho) A) x = rhs is solved by a
 *     banded Cholesky after an optional symmetric permutation.  Same iterates in exact arithmetic.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define RHO_TOL 1e-4

enum {
    OSQP_SOLVED = 1,
    OSQP_SOLVED_INACCURATE = 2,
    OSQP_MAX_ITER_REACHED = -2,
    OSQP_PRIMAL_INFEASIBLE = -3,
    OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3,
    OSQP_DUAL_INFEASIBLE = -4,
    OSQP_DUAL_INFEASIBLE_INACCURATE = 4,
    OSQP_NON_CVX = -7,
    OSQP_ORACLE_FACTOR_FAILED = -100
};

typedef struct {
    double rho, sigma, alpha;
    double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int max_iter, scaling, check_termination;
    int adaptive_rho, adaptive_rho_interval;
    double adaptive_rho_tolerance;
} orc_osqp_settings;

void orc_osqp_default_settings(orc_osqp_settings *s)
{
    s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6;
    s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
    s->max_iter = 4000; s->scaling = 10; s->check_termination = 25;
    s->adaptive_rho = 1; s->adaptive_rho_interval = 25; s->adaptive_rho_tolerance = 5.0;
}

typedef struct {
    int n, m, bw;
    /* scaled problem */
    double *Pp_x; /* P values (CSC upper, same pattern as input) */
    const int *Pp, *Pi;
    double *A_x;
    const int *Ap, *Ai;
    double *q, *l, *u;
    double *D, *E, *Dinv, *Einv;
    double c, cinv;
    /* CSR view of A for A'RA assembly */
    int *Rp, *Rj; int *Rk; /* Rk: index into A_x */
    /* rho */
    double rho; double *rho_vec, *rho_inv_vec; int *constr_type;
    /* band Cholesky of permuted reduced matrix: L[j*(bw+1) + (i-j)], i in [j, j+bw] */
    double *band; const int *perm; int *iperm;
    /* iterates */
    double *x, *z, *y, *x_prev, *z_prev, *xt, *zt, *dx, *dy, *Ax, *Px, *Aty, *rhs, *tmpn, *tmpm;
} work_t;

static double norm_inf(const double *v, int n) { double r = 0; for (int i = 0; i < n; ++i) { double a = fabs(v[i]); if (a > r) r = a; } return r; }
static double scaled_norm_inf(const double *s, const double *v, int n) { double r = 0; for (int i = 0; i < n; ++i) { double a = fabs(s[i] * v[i]); if (a > r) r = a; } return r; }

static void A_mul(const work_t *w, const double *x, double *y) /* y = A x */
{
    memset(y, 0, sizeof(double) * w->m);
    for (int j = 0; j < w->n; ++j)
        for (int k = w->Ap[j]; k < w->Ap[j + 1]; ++k) y[w->Ai[k]] += w->A_x[k] * x[j];
}
static void At_mul(const work_t *w, const double *y, double *x) /* x = A' y */
{
    for (int j = 0; j < w->n; ++j) {
        double s = 0;
        for (int k = w->Ap[j]; k < w->Ap[j + 1]; ++k) s += w->A_x[k] * y[w->Ai[k]];
        x[j] = s;
    }
}
static void P_mul(const work_t *w, const double *x, double *y) /* y = P x, P upper CSC symmetric */
{
    memset(y, 0, sizeof(double) * w->n);
    for (int j = 0; j < w->n; ++j)
        for (int k = w->Pp[j]; k < w->Pp[j + 1]; ++k) {
            int i = w->Pi[k];
            y[i] += w->Pp_x[k] * x[j];
            if (i != j) y[j] += w->Pp_x[k] * x[i];
        }
}

static void limit_scaling(double *v, int n)
{
    for (int i = 0; i < n; ++i) {
        v[i] = v[i] < MIN_SCALING ? 1.0 : v[i];
        v[i] = v[i] > MAX_SCALING ? MAX_SCALING : v[i];
    }
}

/* OSQP scale_data: `scaling` passes of Ruiz equilibration on the KKT matrix + cost scaling */
static void scale_data(work_t *w, int scaling)
{
    int n = w->n, m = w->m;
    double *Dt = w->tmpn, *Et = w->tmpm;
    for (int i = 0; i < n; ++i) w->D[i] = 1.0;
    for (int i = 0; i < m; ++i) w->E[i] = 1.0;
    w->c = 1.0;
    for (int it = 0; it < scaling; ++it) {
        /* inf-norm of the columns of [P A'; A 0] */
        for (int j = 0; j < n; ++j) Dt[j] = 0;
        for (int j = 0; j < n; ++j)
            for (int k = w->Pp[j]; k < w->Pp[j + 1]; ++k) {
                int i = w->Pi[k]; double a = fabs(w->Pp_x[k]);
                if (a > Dt[j]) Dt[j] = a;
                if (i != j && a > Dt[i]) Dt[i] = a;
            }
        for (int i = 0; i < m; ++i) Et[i] = 0;
        for (int j = 0; j < n; ++j)
            for (int k = w->Ap[j]; k < w->Ap[j + 1]; ++k) {
                double a = fabs(w->A_x[k]);
                if (a > Dt[j]) Dt[j] = a;
                if (a > Et[w->Ai[k]]) Et[w->Ai[k]] = a;
            }
        limit_scaling(Dt, n);
        limit_scaling(Et, m);
        for (int j = 0; j < n; ++j) Dt[j] = 1.0 / sqrt(Dt[j]);
        for (int i = 0; i < m; ++i) Et[i] = 1.0 / sqrt(Et[i]);
        /* P = Dt P Dt, A = Et A Dt, q = Dt q */
        for (int j = 0; j < n; ++j)
            for (int k = w->Pp[j]; k < w->Pp[j + 1]; ++k) w->Pp_x[k] *= Dt[j] * Dt[w->Pi[k]];
        for (int j = 0; j < n; ++j)
            for (int k = w->Ap[j]; k < w->Ap[j + 1]; ++k) w->A_x[k] *= Dt[j] * Et[w->Ai[k]];
        for (int j = 0; j < n; ++j) { w->q[j] *= Dt[j]; w->D[j] *= Dt[j]; }
        for (int i = 0; i < m; ++i) w->E[i] *= Et[i];
        /* cost scaling: c_temp = 1 / max(mean(col-norms of P), ||q||_inf) */
        for (int j = 0; j < n; ++j) Dt[j] = 0;
        for (int j = 0; j < n; ++j)
            for (int k = w->Pp[j]; k < w->Pp[j + 1]; ++k) {
                int i = w->Pi[k]; double a = fabs(w->Pp_x[k]);
                if (a > Dt[j]) Dt[j] = a;
                if (i != j && a > Dt[i]) Dt[i] = a;
            }
        double c_temp = 0;
        for (int j = 0; j < n; ++j) c_temp += Dt[j];
        c_temp /= n;
        double inf_norm_q = norm_inf(w->q, n);
        limit_scaling(&inf_norm_q, 1);
        c_temp = c_temp > inf_norm_q ? c_temp : inf_norm_q;
        limit_scaling(&c_temp, 1);
        c_temp = 1.0 / c_temp;
        for (int k = 0; k < w->Pp[n]; ++k) w->Pp_x[k] *= c_temp;
        for (int j = 0; j < n; ++j) w->q[j] *= c_temp;
        w->c *= c_temp;
    }
    for (int j = 0; j < n; ++j) w->Dinv[j] = 1.0 / w->D[j];
    for (int i = 0; i < m; ++i) w->Einv[i] = 1.0 / w->E[i];
    w->cinv = 1.0 / w->c;
    for (int i = 0; i < m; ++i) { w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
}

static void set_rho_vec(work_t *w)
{
    for (int i = 0; i < w->m; ++i) {
        if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) {
            w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN;
        } else if (w->u[i] - w->l[i] < RHO_TOL) {
            w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->rho;
        } else {
            w->constr_type[i] = 0; w->rho_vec[i] = w->rho;
        }
        w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
    }
}
static void update_rho_vec(work_t *w)
{
    for (int i = 0; i < w->m; ++i) {
        if (w->constr_type[i] == 0) w->rho_vec[i] = w->rho;
        else if (w->constr_type[i] == 1) w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->rho;
        w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
    }
}

/* assemble S = P + sigma I + A' R A in band storage (permuted), then Cholesky in place */
static int factorize(work_t *w, double sigma)
{
    int n = w->n, bw = w->bw, ld = bw + 1;
    memset(w->band, 0, sizeof(double) * (size_t)n * ld);
#define BAND(i, j) w->band[(size_t)(j) * ld + ((i) - (j))] /* i >= j */
    for (int j = 0; j < n; ++j) {
        int pj = w->iperm[j];
        BAND(pj, pj) += sigma;
        for (int k = w->Pp[j]; k < w->Pp[j + 1]; ++k) {
            int pi = w->iperm[w->Pi[k]];
            if (pi >= pj) BAND(pi, pj) += w->Pp_x[k]; else BAND(pj, pi) += w->Pp_x[k];
        }
    }
    for (int r = 0; r < w->m; ++r) {
        double rho = w->rho_vec[r];
        for (int a = w->Rp[r]; a < w->Rp[r + 1]; ++a)
            for (int b = w->Rp[r]; b < w->Rp[r + 1]; ++b) {
                int pi = w->iperm[w->Rj[a]], pj = w->iperm[w->Rj[b]];
                if (pi >= pj) BAND(pi, pj) += rho * w->A_x[w->Rk[a]] * w->A_x[w->Rk[b]];
            }
    }
    for (int j = 0; j < n; ++j) {
        double d = BAND(j, j);
        if (!(d > 0)) return -1;
        d = sqrt(d);
        BAND(j, j) = d;
        int imax = j + bw < n - 1 ? j + bw : n - 1;
        for (int i = j + 1; i <= imax; ++i) BAND(i, j) /= d;
        for (int k = j + 1; k <= imax; ++k) {
            double ljk = BAND(k, j);
            if (ljk == 0) continue;
            for (int i = k; i <= imax; ++i) BAND(i, k) -= BAND(i, j) * ljk;
        }
    }
    return 0;
}
static void chol_solve(const work_t *w, double *b /* permuted in/out */)
{
    int n = w->n, bw = w->bw, ld = bw + 1;
    for (int j = 0; j < n; ++j) {
        b[j] /= w->band[(size_t)j * ld];
        int imax = j + bw < n - 1 ? j + bw : n - 1;
        for (int i = j + 1; i <= imax; ++i) b[i] -= w->band[(size_t)j * ld + (i - j)] * b[j];
    }
    for (int j = n - 1; j >= 0; --j) {
        int imax = j + bw < n - 1 ? j + bw : n - 1;
        double s = b[j];
        for (int i = j + 1; i <= imax; ++i) s -= w->band[(size_t)j * ld + (i - j)] * b[i];
        b[j] = s / w->band[(size_t)j * ld];
    }
#undef BAND
}

static double compute_pri_res(work_t *w, int unscaled)
{
    A_mul(w, w->x, w->Ax);
    for (int i = 0; i < w->m; ++i) w->z_prev[i] = w->Ax[i] - w->z[i];
    return unscaled ? scaled_norm_inf(w->Einv, w->z_prev, w->m) : norm_inf(w->z_prev, w->m);
}
static double compute_dua_res(work_t *w, int unscaled)
{
    P_mul(w, w->x, w->Px);
    At_mul(w, w->y, w->Aty);
    for (int j = 0; j < w->n; ++j) w->x_prev[j] = w->q[j] + w->Px[j] + w->Aty[j];
    return unscaled ? w->cinv * scaled_norm_inf(w->Dinv, w->x_prev, w->n) : norm_inf(w->x_prev, w->n);
}

static int is_primal_infeasible(work_t *w, double eps)
{
    int m = w->m, n = w->n;
    double *dy = w->tmpm;
    for (int i = 0; i < m; ++i) {
        double d = w->dy[i];
        if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
            if (w->l[i] < -OSQP_INFTY * MIN_SCALING) d = 0.0; else d = d < 0 ? d : 0.0;
        } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) d = d > 0 ? d : 0.0;
        dy[i] = d;
    }
    double norm_dy = scaled_norm_inf(w->E, dy, m);
    if (norm_dy > eps) {
        double lhs = 0;
        for (int i = 0; i < m; ++i) lhs += w->u[i] * (dy[i] > 0 ? dy[i] : 0) + w->l[i] * (dy[i] < 0 ? dy[i] : 0);
        if (lhs < -eps * norm_dy) {
            At_mul(w, dy, w->tmpn);
            return scaled_norm_inf(w->Dinv, w->tmpn, n) < eps * norm_dy;
        }
    }
    return 0;
}
static int is_dual_infeasible(work_t *w, double eps)
{
    int m = w->m, n = w->n;
    double norm_dx = scaled_norm_inf(w->D, w->dx, n);
    double cs = w->c;
    if (norm_dx > eps) {
        double qdx = 0;
        for (int j = 0; j < n; ++j) qdx += w->q[j] * w->dx[j];
        if (qdx < -cs * eps * norm_dx) {
            P_mul(w, w->dx, w->tmpn);
            if (scaled_norm_inf(w->Dinv, w->tmpn, n) < cs * eps * norm_dx) {
                A_mul(w, w->dx, w->tmpm);
                for (int i = 0; i < m; ++i) {
                    double a = w->Einv[i] * w->tmpm[i];
                    if ((w->u[i] < OSQP_INFTY * MIN_SCALING && a > eps * norm_dx) ||
                        (w->l[i] > -OSQP_INFTY * MIN_SCALING && a < -eps * norm_dx)) return 0;
                }
                return 1;
            }
        }
    }
    return 0;
}

/* returns 0 = continue, else status */
static int check_termination(work_t *w, const orc_osqp_settings *s, double pri_res, double dua_res, int approximate)
{
    double eps_abs = s->eps_abs, eps_rel = s->eps_rel, epi = s->eps_prim_inf, edi = s->eps_dual_inf;
    int prim_ok = 0, dual_ok = 0, pinf = 0, dinf = 0;
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) return OSQP_NON_CVX;
    if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
    if (w->m == 0) prim_ok = 1;
    else {
        double a = scaled_norm_inf(w->Einv, w->z, w->m), b = scaled_norm_inf(w->Einv, w->Ax, w->m);
        double eps_prim = eps_abs + eps_rel * (a > b ? a : b);
        if (pri_res < eps_prim) prim_ok = 1; else pinf = is_primal_infeasible(w, epi);
    }
    {
        double a = scaled_norm_inf(w->Dinv, w->q, w->n), b = scaled_norm_inf(w->Dinv, w->Aty, w->n),
               c = scaled_norm_inf(w->Dinv, w->Px, w->n);
        double mx = a > b ? a : b; mx = mx > c ? mx : c;
        double eps_dual = eps_abs + eps_rel * w->cinv * mx;
        if (dua_res < eps_dual) dual_ok = 1; else dinf = is_dual_infeasible(w, edi);
    }
    if (prim_ok && dual_ok) return approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED;
    if (pinf) return approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE;
    if (dinf) return approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE;
    return 0;
}

static double compute_rho_estimate(work_t *w)
{
    /* z_prev, x_prev hold the scaled residual vectors from compute_pri_res/compute_dua_res */
    double pri = norm_inf(w->z_prev, w->m), dua = norm_inf(w->x_prev, w->n);
    double a = norm_inf(w->z, w->m), b = norm_inf(w->Ax, w->m);
    pri /= ((a > b ? a : b) + 1e-10);
    double c = norm_inf(w->q, w->n), d = norm_inf(w->Aty, w->n), e = norm_inf(w->Px, w->n);
    double mx = c > d ? c : d; mx = mx > e ? mx : e;
    dua /= (mx + 1e-10);
    double r = w->rho * sqrt(pri / (dua + 1e-10));
    r = r > RHO_MIN ? r : RHO_MIN;
    r = r < RHO_MAX ? r : RHO_MAX;
    return r;
}

/*
 * Solve  min 1/2 x'Px + q'x  s.t. l <= Ax <= u.
 * P: CSC upper triangle (n x n); A: CSC (m x n); perm: optional symmetric permutation for the
 * band Cholesky, perm[new] = old (NULL = identity).
 * info[8] = {pri_res, dua_res, obj_val, rho_final, rho_updates, n_factor, n_checks, bandwidth}.
 * Returns the OSQP status code; x_out/y_out unscaled (NaN-filled on infeasible, as OSQP does).
 */
int orc_osqp_solve(int n, int m, const int *Pp, const int *Pi, const double *Px_in, const double *q_in,
                   const int *Ap, const int *Ai, const double *Ax_in, const double *l_in,
                   const double *u_in, const orc_osqp_settings *s, const int *perm,
                   double *x_out, double *y_out, int *iters_out, double *info)
{
    work_t W, *w = &W;
    memset(w, 0, sizeof(W));
    w->n = n; w->m = m; w->Pp = Pp; w->Pi = Pi; w->Ap = Ap; w->Ai = Ai; w->perm = perm;
    int pnz = Pp[n], anz = Ap[n];
#define DALLOC(k) (double *)calloc((size_t)((k) > 0 ? (k) : 1), sizeof(double))
    w->Pp_x = DALLOC(pnz); memcpy(w->Pp_x, Px_in, sizeof(double) * pnz);
    w->A_x = DALLOC(anz); memcpy(w->A_x, Ax_in, sizeof(double) * anz);
    w->q = DALLOC(n); memcpy(w->q, q_in, sizeof(double) * n);
    w->l = DALLOC(m); w->u = DALLOC(m);
    for (int i = 0; i < m; ++i) { /* python interface: np.maximum(l, -OSQP_INFTY), np.minimum(u, OSQP_INFTY) */
        w->l[i] = l_in[i] > -OSQP_INFTY ? l_in[i] : -OSQP_INFTY;
        w->u[i] = u_in[i] < OSQP_INFTY ? u_in[i] : OSQP_INFTY;
    }
    w->D = DALLOC(n); w->Dinv = DALLOC(n); w->E = DALLOC(m); w->Einv = DALLOC(m);
    w->rho_vec = DALLOC(m); w->rho_inv_vec = DALLOC(m); w->constr_type = (int *)calloc(m > 0 ? m : 1, sizeof(int));
    w->x = DALLOC(n); w->z = DALLOC(m); w->y = DALLOC(m); w->x_prev = DALLOC(n); w->z_prev = DALLOC(m);
    w->xt = DALLOC(n); w->zt = DALLOC(m); w->dx = DALLOC(n); w->dy = DALLOC(m);
    w->Ax = DALLOC(m); w->Px = DALLOC(n); w->Aty = DALLOC(n); w->rhs = DALLOC(n);
    w->tmpn = DALLOC(n); w->tmpm = DALLOC(m);
    w->iperm = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) w->iperm[perm ? perm[i] : i] = i;
    /* CSR view */
    w->Rp = (int *)calloc(m + 1, sizeof(int)); w->Rj = (int *)malloc(sizeof(int) * (anz > 0 ? anz : 1));
    w->Rk = (int *)malloc(sizeof(int) * (anz > 0 ? anz : 1));
    for (int k = 0; k < anz; ++k) w->Rp[Ai[k] + 1]++;
    for (int i = 0; i < m; ++i) w->Rp[i + 1] += w->Rp[i];
    {
        int *fill = (int *)calloc(m > 0 ? m : 1, sizeof(int));
        for (int j = 0; j < n; ++j)
            for (int k = Ap[j]; k < Ap[j + 1]; ++k) {
                int r = Ai[k], pos = w->Rp[r] + fill[r]++;
                w->Rj[pos] = j; w->Rk[pos] = k;
            }
        free(fill);
    }
    /* bandwidth of permuted S */
    int bw = 0;
    for (int j = 0; j < n; ++j)
        for (int k = Pp[j]; k < Pp[j + 1]; ++k) { int d = abs(w->iperm[Pi[k]] - w->iperm[j]); if (d > bw) bw = d; }
    for (int r = 0; r < m; ++r)
        for (int a = w->Rp[r]; a < w->Rp[r + 1]; ++a)
            for (int b = a + 1; b < w->Rp[r + 1]; ++b) { int d = abs(w->iperm[w->Rj[a]] - w->iperm[w->Rj[b]]); if (d > bw) bw = d; }
    w->bw = bw;
    w->band = DALLOC((size_t)n * (bw + 1));

    int status = 0, iter = 0, n_factor = 0, n_checks = 0, rho_updates = 0;
    double pri_res = 0, dua_res = 0;
    if (s->scaling > 0) scale_data(w, s->scaling);
    else {
        for (int j = 0; j < n; ++j) w->D[j] = w->Dinv[j] = 1.0;
        for (int i = 0; i < m; ++i) w->E[i] = w->Einv[i] = 1.0;
        w->c = w->cinv = 1.0;
    }
    w->rho = s->rho;
    set_rho_vec(w);
    if (factorize(w, s->sigma)) { status = OSQP_ORACLE_FACTOR_FAILED; goto finish; }
    ++n_factor;

    for (iter = 1; iter <= s->max_iter; ++iter) {
        double *t;
        t = w->x; w->x = w->x_prev; w->x_prev = t;
        t = w->z; w->z = w->z_prev; w->z_prev = t;
        /* update_xz_tilde: (P + sigma I + A'RA) xt = sigma x_prev - q + A'(R z_prev - y); zt = A xt */
        for (int i = 0; i < m; ++i) w->tmpm[i] = w->rho_vec[i] * w->z_prev[i] - w->y[i];
        At_mul(w, w->tmpm, w->tmpn);
        for (int j = 0; j < n; ++j) w->rhs[w->iperm[j]] = s->sigma * w->x_prev[j] - w->q[j] + w->tmpn[j];
        chol_solve(w, w->rhs);
        for (int j = 0; j < n; ++j) w->xt[j] = w->rhs[w->iperm[j]];
        A_mul(w, w->xt, w->zt);
        /* update_x */
        for (int j = 0; j < n; ++j) {
            w->x[j] = s->alpha * w->xt[j] + (1.0 - s->alpha) * w->x_prev[j];
            w->dx[j] = w->x[j] - w->x_prev[j];
        }
        /* update_z */
        for (int i = 0; i < m; ++i) {
            double v = s->alpha * w->zt[i] + (1.0 - s->alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
            v = v > w->l[i] ? v : w->l[i];
            v = v < w->u[i] ? v : w->u[i];
            w->z[i] = v;
        }
        /* update_y */
        for (int i = 0; i < m; ++i) {
            double d = s->alpha * w->zt[i] + (1.0 - s->alpha) * w->z_prev[i] - w->z[i];
            d *= w->rho_vec[i];
            w->dy[i] = d;
            w->y[i] += d;
        }
        int can_check = s->check_termination && (iter % s->check_termination == 0);
        int have_info = 0;
        if (can_check) {
            pri_res = compute_pri_res(w, s->scaling > 0);
            dua_res = compute_dua_res(w, s->scaling > 0);
            have_info = 1; ++n_checks;
            status = check_termination(w, s, pri_res, dua_res, 0);
            if (status) break;
        }
        if (s->adaptive_rho && s->adaptive_rho_interval && (iter % s->adaptive_rho_interval == 0)) {
            if (!have_info) {
                pri_res = compute_pri_res(w, s->scaling > 0);
                dua_res = compute_dua_res(w, s->scaling > 0);
            }
            double rho_new = compute_rho_estimate(w);
            if (rho_new > w->rho * s->adaptive_rho_tolerance || rho_new < w->rho / s->adaptive_rho_tolerance) {
                w->rho = rho_new;
                update_rho_vec(w);
                if (factorize(w, s->sigma)) { status = OSQP_ORACLE_FACTOR_FAILED; break; }
                ++n_factor; ++rho_updates;
            }
        }
    }
    if (!status) { /* max_iter reached (osqp.c, after the main loop) */
        iter = s->max_iter;
        pri_res = compute_pri_res(w, s->scaling > 0);
        dua_res = compute_dua_res(w, s->scaling > 0);
        /* "if (!can_check_termination) { update_info(...); check_termination(work, 0); }": when the last iteration was not
         * a check iteration (max_iter not a multiple of check_termination, or checks disabled) OSQP first runs a NORMAL
         * termination check on the final iterate ... */
        if (!(s->check_termination && (s->max_iter % s->check_termination == 0))) {
            ++n_checks;
            status = check_termination(w, s, pri_res, dua_res, 0);
        }
        /* ... and only if that leaves the problem unsolved the approximate one (10x tolerances; it can also return the
         * inaccurate infeasibility statuses 3 / 4), else OSQP_MAX_ITER_REACHED */
        if (!status) status = check_termination(w, s, pri_res, dua_res, 1);
        if (!status) status = OSQP_MAX_ITER_REACHED;
    }
finish:
    if (status == OSQP_PRIMAL_INFEASIBLE || status == OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
        status == OSQP_DUAL_INFEASIBLE || status == OSQP_DUAL_INFEASIBLE_INACCURATE ||
        status == OSQP_NON_CVX || status == OSQP_ORACLE_FACTOR_FAILED) {
        for (int j = 0; j < n; ++j) x_out[j] = NAN;
        if (y_out) for (int i = 0; i < m; ++i) y_out[i] = NAN;
    } else {
        for (int j = 0; j < n; ++j) x_out[j] = w->D[j] * w->x[j];
        if (y_out) for (int i = 0; i < m; ++i) y_out[i] = w->cinv * w->E[i] * w->y[i];
    }
    if (iters_out) *iters_out = iter;
    if (info) {
        double obj = 0;
        P_mul(w, w->x, w->Px);
        for (int j = 0; j < n; ++j) obj += 0.5 * w->x[j] * w->Px[j] + w->q[j] * w->x[j];
        info[0] = pri_res; info[1] = dua_res; info[2] = obj * w->cinv; info[3] = w->rho;
        info[4] = rho_updates; info[5] = n_factor; info[6] = n_checks; info[7] = bw;
    }
    free(w->Pp_x); free(w->A_x); free(w->q); free(w->l); free(w->u); free(w->D); free(w->Dinv);
    free(w->E); free(w->Einv); free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
    free(w->x); free(w->z); free(w->y); free(w->x_prev); free(w->z_prev); free(w->xt); free(w->zt);
    free(w->dx); free(w->dy); free(w->Ax); free(w->Px); free(w->Aty); free(w->rhs); free(w->tmpn);
    free(w->tmpm); free(w->iperm); free(w->Rp); free(w->Rj); free(w->Rk); free(w->band);
    return status;
}
