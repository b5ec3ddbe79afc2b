/*
 * oracle/orc_batch.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * One full control step of the reference (MPC.get_control + BicycleModel.drive,
 * /root/reference/src/MPC.py:161-222, spatial_bicycle_models.py:221-244) restated on top of
 * mpc_oracle.c / osqp_oracle.c, plus OpenMP loops over independent scenarios.  The loops are what
 * bench.py times as the CPU baseline ("port": reference algorithm, C, fp64, all host cores).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* ---- mirrored declarations (see mpc_oracle.c / osqp_oracle.c) ---- */
typedef struct {
    int n_wp;
    int circular;
    const double *x, *y, *psi, *kappa, *v_ref;
    const double *seg_len, *ds_next, *cos_psi, *sin_psi, *cos_ub, *sin_ub, *cos_lb, *sin_lb, *border;
} orc_path;
typedef struct {
    int N;
    double Q[3], R[2], QN[3];
    double xmin[3], xmax[3], umin[2], umax[2];
    double ay_max, L, safety_margin;
} orc_mpc_cfg;
typedef struct {
    double rho, sigma, alpha;
    double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int max_iter, scaling, check_termination;
    int adaptive_rho, adaptive_rho_interval;
    double adaptive_rho_tolerance;
} orc_osqp_settings;

int orc_update_path_constraints(const int8_t *data, int H, int W, double ox, double oy, double res,
                                const orc_path *p, long wp_id, int N, double min_width,
                                double safety_margin, double *ub_out, double *lb_out,
                                double *cells_sm, long *n_cells_tested, long *n_sectors);
int orc_get_current_waypoint(const double *length_cum, int n_wp, double s);
void orc_t2s(double x, double y, double psi, double wx, double wy, double wpsi, double cos_wpsi,
             double sin_wpsi, double *out3);
void orc_drive(double *state4, double e_y, double e_psi, double kappa_wp, double v, double delta,
               double L, double Ts);
int orc_mpc_assemble(const orc_path *p, const orc_mpc_cfg *c, long wp_id, const double *x0,
                     const double *current_control, const double *ub, const double *lb, double *Pd,
                     double *q, int *Ap, int *Ai, double *Ax, double *l, double *u);
int orc_osqp_solve(int n, int m, const int *Pp, const int *Pi, const double *Px_in, const double *q_in,
                   const int *Ap, const int *Ai, const double *Ax_in, const double *l_in,
                   const double *u_in, const orc_osqp_settings *s, const int *perm, double *x_out,
                   double *y_out, int *iters_out, double *info);

typedef struct {
    orc_path path;
    const double *length_cum; /* np.cumsum(segment_lengths) */
    orc_mpc_cfg cfg;
    orc_osqp_settings osqp;
    int H, W;
    double ox, oy, res;
    double Ts;
} orc_world;

/* step status bits */
enum { ST_OK = 0, ST_QP_FALLBACK = 1, ST_DEAD = 2, ST_RAYCAST_FAIL = 4, ST_LOCALIZE_FAIL = 8 };

/* stage-interleaved ordering [x0 u0 x1 u1 ... xN] for the band Cholesky (bandwidth 7) */
static void stage_perm(int N, int *perm)
{
    int k = 0;
    for (int s = 0; s <= N; ++s) {
        for (int i = 0; i < 3; ++i) perm[k++] = 3 * s + i;
        if (s < N) for (int i = 0; i < 2; ++i) perm[k++] = 3 * (N + 1) + 2 * s + i;
    }
}

/*
 * One MPC.get_control() [+ drive()].
 * state4 = (x, y, psi, s) in/out; current_control[2N] in/out; infeas in/out.
 * Optional outputs (may be NULL): u2, xsol[5N+3] (dec.x), ub_lb[2N], spatial3, wp_id_out, iters, qp_status.
 */
int orc_mpc_step(const orc_world *w, const int8_t *grid, double *state4, double *current_control,
                 int *infeas, int do_drive, double *u2, double *xsol, double *ub_lb, double *spatial3,
                 int *wp_id_out, int *iters_out, int *qp_status_out)
{
    const int N = w->cfg.N, n = 5 * N + 3, m = 8 * N + 6, nnz = 16 * N + 6;
    int ret = ST_OK;
    int wp_id = orc_get_current_waypoint(w->length_cum, w->path.n_wp, state4[3]); /* MPC.py:172 */
    if (wp_id < 0) return ST_LOCALIZE_FAIL;
    double sp[3];
    orc_t2s(state4[0], state4[1], state4[2], w->path.x[wp_id], w->path.y[wp_id], w->path.psi[wp_id],
            w->path.cos_psi[wp_id], w->path.sin_psi[wp_id], sp);                /* MPC.py:175 */
    double *buf = (double *)malloc(sizeof(double) * (size_t)(2 * N + 2 * n + nnz + 2 * m + n));
    double *ub = buf, *lb = ub + N, *Pd = lb + N, *q = Pd + n, *Ax = q + n, *l = Ax + nnz, *u = l + m,
           *x = u + m;
    int *ibuf = (int *)malloc(sizeof(int) * (size_t)(n + 1 + nnz + 2 * n + 1 + n));
    int *Ap = ibuf, *Ai = Ap + n + 1, *Pp = Ai + nnz, *Pi = Pp + n + 1, *perm = Pi + n;
    int rs = orc_update_path_constraints(grid, w->H, w->W, w->ox, w->oy, w->res, &w->path, wp_id + 1, N,
                                         2 * w->cfg.safety_margin, w->cfg.safety_margin, ub, lb, NULL,
                                         NULL, NULL);                            /* MPC.py:116-118 */
    int qp_status = 0, iters = 0;
    double v = 0, delta = 0;
    if (rs) { ret |= ST_RAYCAST_FAIL; goto out; }
    if (orc_mpc_assemble(&w->path, &w->cfg, wp_id, sp, current_control, ub, lb, Pd, q, Ap, Ai, Ax, l, u)) {
        ret |= ST_RAYCAST_FAIL; goto out;
    }
    for (int j = 0; j < n; ++j) { Pp[j] = j; Pi[j] = j; }
    Pp[n] = n;
    stage_perm(N, perm);
    qp_status = orc_osqp_solve(n, m, Pp, Pi, Pd, q, Ap, Ai, Ax, l, u, &w->osqp, perm, x, NULL, &iters, NULL);
    if (xsol) memcpy(xsol, x, sizeof(double) * n);
    if (!isnan(x[0])) {
        /* MPC.py:187-206 */
        for (int k = 0; k < N; ++k) {
            current_control[2 * k] = x[3 * (N + 1) + 2 * k];
            current_control[2 * k + 1] = atan(x[3 * (N + 1) + 2 * k + 1] * w->cfg.L);
        }
        v = current_control[0]; delta = current_control[1];
        *infeas = 0;
    } else {
        /* MPC.py:208-216 */
        int id = 2 * (*infeas + 1);
        v = current_control[id]; delta = current_control[id + 1];
        *infeas += 1;
        ret |= ST_QP_FALLBACK;
    }
    if (*infeas == N - 1) ret |= ST_DEAD; /* MPC.py:218-220 exit(1) */
    if (do_drive && !(ret & ST_DEAD))
        orc_drive(state4, sp[0], sp[1], w->path.kappa[wp_id], v, delta, w->cfg.L, w->Ts);
out:
    if (u2) { u2[0] = v; u2[1] = delta; }
    if (ub_lb && !rs) { memcpy(ub_lb, ub, sizeof(double) * N); memcpy(ub_lb + N, lb, sizeof(double) * N); }
    if (spatial3) memcpy(spatial3, sp, sizeof(sp));
    if (wp_id_out) *wp_id_out = wp_id;
    if (iters_out) *iters_out = iters;
    if (qp_status_out) *qp_status_out = qp_status;
    free(buf); free(ibuf);
    return ret;
}

/*
 * B independent scenarios, `steps` closed-loop steps each, OpenMP over scenarios.
 * grids: either one shared grid (grid_stride = 0) or B grids of H*W int8.
 * states[B][4], controls[B][2N], infeas[B] in/out; stats[4] = {steps done, qp solves, admm iters, fallbacks}.
 */
int orc_batch_closed_loop(const orc_world *w, const int8_t *grids, long grid_stride, int B, int steps,
                          double *states, double *controls, int *infeas, int *alive, double *stats,
                          int n_threads)
{
    const int N = w->cfg.N;
    double t_steps = 0, t_iters = 0, t_fb = 0;
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : t_steps, t_iters, t_fb)
    for (int b = 0; b < B; ++b) {
        for (int k = 0; k < steps; ++k) {
            if (alive && !alive[b]) break;
            int it = 0, qs = 0;
            int r = orc_mpc_step(w, grids + grid_stride * b, states + 4 * b, controls + 2 * (size_t)N * b,
                                 infeas + b, 1, NULL, NULL, NULL, NULL, NULL, &it, &qs);
            t_steps += 1; t_iters += it;
            if (r & ST_QP_FALLBACK) t_fb += 1;
            if (r & (ST_DEAD | ST_RAYCAST_FAIL | ST_LOCALIZE_FAIL)) { if (alive) alive[b] = 0; break; }
        }
    }
    if (stats) { stats[0] = t_steps; stats[1] = t_steps; stats[2] = t_iters; stats[3] = t_fb; }
    return 0;
}

/* QP-only batch: B QPs sharing one sparsity pattern (the MPC pattern), values per instance. */
int orc_batch_qp_solve(int N, int B, const double *Pd, const double *q, const int *Ap, const int *Ai,
                       const double *Ax, const double *l, const double *u, const orc_osqp_settings *s,
                       double *x_out, int *iters, int *status, int n_threads)
{
    const int n = 5 * N + 3, m = 8 * N + 6, nnz = Ap[n];
    int *Pp = (int *)malloc(sizeof(int) * (2 * n + 1 + n));
    int *Pi = Pp + n + 1, *perm = Pi + n;
    for (int j = 0; j < n; ++j) { Pp[j] = j; Pi[j] = j; }
    Pp[n] = n;
    stage_perm(N, perm);
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int it = 0;
        int st = orc_osqp_solve(n, m, Pp, Pi, Pd + (size_t)n * b, q + (size_t)n * b, Ap, Ai,
                                Ax + (size_t)nnz * b, l + (size_t)m * b, u + (size_t)m * b, s, perm,
                                x_out + (size_t)n * b, NULL, &it, NULL);
        if (iters) iters[b] = it;
        if (status) status[b] = st;
    }
    free(Pp);
    return 0;
}

int orc_num_threads(void) { return omp_get_max_threads(); }
