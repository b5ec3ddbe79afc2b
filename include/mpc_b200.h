/*
 * mpc_b200.h -- C ABI of the B200 batched closed-loop MPC engine (libmpc_b200.so).
 *
 * The reference (matssteinweg/Multi-Purpose-MPC) has no FFI: its "interface" is four Python
 * classes.  Each entry point below therefore names the reference METHOD it replaces
 * (file:line in /root/reference/src) -- the Python host classes in multi-purpose-mpc_b200/ keep
 * those method names and call these functions through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every function returns 0 on success, a negative MPC_E_* code on error; mpc_last_error()
 *     returns the message of the last failure on the calling thread.
 *   - one engine handle per GPU (cudaSetDevice is the caller's business), not thread-safe.
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensor data_ptr());
 *     pointers named h_* are HOST pointers.  The library never frees caller memory.
 *   - all kernels are enqueued on the stream given to mpc_engine_set_stream (default stream 0);
 *     calls are asynchronous unless documented otherwise.
 *   - there is NO CPU fallback: if no CUDA device is usable every call fails with MPC_E_CUDA.
 *
 * Batch layouts (B = number of independent scenarios, N = horizon):
 *   state      double[4][B]   rows x, y, psi, s                     (TemporalState + model.s)
 *   spatial    double[2][B]   rows e_y, e_psi   (t is always 0, sbm.py:217)
 *   wp_id      int32[B]
 *   control    double[B][2N]  MPC.current_control (v0, delta0, v1, delta1, ...)   (MPC.py:56,194)
 *   ub, lb     double[B][N]   ReferencePath.update_path_constraints outputs       (rp.py:648)
 *   x_out      double[B][5N+3] dec.x in the reference's order [x0..xN | u0..uN-1] (MPC.py:187,197)
 *   u_out      double[B][2]   (v, delta) returned by MPC.get_control              (MPC.py:203)
 */
#ifndef MPC_B200_H
#define MPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_B200_ABI_VERSION 1

/* error codes */
#define MPC_OK 0
#define MPC_E_INVALID (-1)
#define MPC_E_CUDA (-2)
#define MPC_E_STATE (-3) /* call order: path / grid not set yet */
#define MPC_E_UNSUPPORTED (-4)

/* per-scenario QP status: the OSQP status values the reference would see in dec.info.status_val */
#define MPC_QP_SOLVED 1
#define MPC_QP_SOLVED_INACCURATE 2 /* max_iter reached, residuals within 10x tolerances */
#define MPC_QP_PRIMAL_INFEASIBLE_INACCURATE 3 /* max_iter reached, certificate within 10x tolerances: no solution (NaN) */
#define MPC_QP_DUAL_INFEASIBLE_INACCURATE 4
#define MPC_QP_MAX_ITER (-2)
#define MPC_QP_PRIMAL_INFEASIBLE (-3)
#define MPC_QP_DUAL_INFEASIBLE (-4)
#define MPC_QP_NON_CVX (-7)

/* per-scenario step flags (bitmask) -- H6 of SURVEY.md: the reference's print/exit/exception
 * paths become status bits */
#define MPC_ST_QP_FALLBACK 1   /* "Infeasible problem. Previously predicted control signal used!" (MPC.py:210) */
#define MPC_ST_DEAD 2          /* infeasibility_counter reached N-1: reference exit(1)           (MPC.py:218-220) */
#define MPC_ST_NO_SEGMENT 4    /* first horizon waypoint has no free segment: reference ValueError (rp.py:547) */
#define MPC_ST_END_OF_PATH 8   /* non-circular path exhausted: reference exit(1)                (rp.py:367-369) */
#define MPC_ST_INDEX_ERROR 16  /* a tested cell lies outside the grid: reference IndexError      (rp.py:496);
                                  also: a ray with more than 8 free segments wider than min_width (engine capacity) */
#define MPC_ST_FINISHED 32     /* s >= path length: the reference's while loop ends   (simulation.py:134) */

typedef struct mpc_engine mpc_engine;

/* Everything MPC.__init__ / BicycleModel.__init__ take (MPC.py:15-59, sbm.py:117-153,323-345) plus
 * the OSQP settings the reference leaves at their defaults (MPC.py:158-159). */
typedef struct mpc_config {
    int32_t N;                 /* horizon                                   MPC.py:30 */
    double Q[3], R[2], QN[3];  /* diagonals of the cost matrices            MPC.py:31-33 */
    double xmin[3], xmax[3];   /* StateConstraints (may be +-inf)           MPC.py:43 */
    double umin[2], umax[2];   /* InputConstraints on (v, kappa)            MPC.py:44 */
    double ay_max;             /*                                           MPC.py:47 */
    double car_length;         /* BicycleModel.length                       sbm.py:130 */
    double car_width;          /* BicycleModel.width; safety_margin = width / sqrt(2), sbm.py:252 */
    double Ts;                 /* sampling time                             sbm.py:141 */
    /* OSQP settings (defaults of osqp 0.6; see oracle/osqp_oracle.c) */
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int32_t max_iter, scaling, check_termination, adaptive_rho_interval;
    double adaptive_rho_tolerance;
    /* engine choices */
    int32_t precision;         /* 0 = fp32 ADMM (default; use with eps >= 1e-4), 1 = fp64 ADMM (validation path) */
} mpc_config;

/* Fills cfg with the reference's defaults (src/simulation.py:100-119 and OSQP 0.6 defaults). */
void mpc_config_default(mpc_config *cfg);

const char *mpc_last_error(void);
int mpc_abi_version(void);

int mpc_engine_create(const mpc_config *cfg, mpc_engine **out);
int mpc_engine_destroy(mpc_engine *h);
int mpc_engine_set_stream(mpc_engine *h, void *cuda_stream);
/* Re-reads the OSQP settings / weights / constraints from cfg (N must not change). */
int mpc_engine_update_config(mpc_engine *h, const mpc_config *cfg);
int mpc_engine_sync(mpc_engine *h);

/* ---- ReferencePath tables (HOST pointers, copied) ----------------------------------------
 * Replaces the attribute reads of Waypoint objects (rp.py:20-57) and ReferencePath.segment_lengths.
 * h_wp: double[12][n_wp], rows:
 *    0 x   1 y   2 psi   3 kappa   4 v_ref (may be NaN until mpc_set_vref)
 *    5 ds_next[k] = get_waypoint(k+1) - get_waypoint(k)  (Waypoint.__sub__, rp.py:57; wrap included)
 *    6 cos(psi)  7 sin(psi)
 *    8 cos(angle_ub)  9 sin(angle_ub)   angle_ub = mod( pi/2 + psi + pi, 2pi) - pi   (rp.py:622, 221)
 *   10 cos(angle_lb) 11 sin(angle_lb)   angle_lb = mod(-pi/2 + psi + pi, 2pi) - pi   (rp.py:624, 224)
 *   The trigonometric rows are computed by the Python host with numpy, i.e. with the very functions
 *   the reference calls, so they are bit-identical to what the reference evaluates per call.
 * h_length_cum: double[n_wp] = np.cumsum(segment_lengths)  (sbm.py:262)
 * h_border: double[n_wp][4] static_border_cells (ub_x, ub_y, lb_x, lb_y), or NULL when
 *           mpc_compute_width will produce them on the device. */
int mpc_set_path(mpc_engine *h, const double *h_wp, const double *h_length_cum, const double *h_border,
                 int32_t n_wp, int32_t circular);
int mpc_set_vref(mpc_engine *h, const double *h_vref, int32_t n_wp);

/* ---- Map (map.py:45-137) -------------------------------------------------------------------
 * h_data: int8[H][W], 1 = free, 0 = occupied (Map.data).  Stored bit-packed (bit = 1 free), row
 * pitch = 64-byte multiple.  Obstacles: per-scenario discs rasterised on the device with
 * Map.add_obstacles' rule (map.py:129-137) into per-scenario copies of the base grid.
 * h_obs: double[n_obs][3] (cx, cy, radius) world units; h_offsets: int32[B+1] CSR offsets.
 * B = 0 / h_obs = NULL: all scenarios share the base grid. */
int mpc_set_base_grid(mpc_engine *h, const int8_t *h_data, int32_t H, int32_t W, double origin_x,
                      double origin_y, double resolution);
int mpc_set_obstacles(mpc_engine *h, const double *h_obs, const int32_t *h_offsets, int32_t B);
/* Reads scenario b's grid back as int8[H][W] (tests). Synchronous. */
int mpc_get_grid(mpc_engine *h, int32_t b, int8_t *h_data_out);

/* ---- K3b: ReferencePath._compute_width (rp.py:206-287), once per base map -------------------
 * Writes h_ub[n_wp], h_lb[n_wp], h_border[n_wp][4] and keeps the border cells in the engine.
 * Synchronous. */
int mpc_compute_width(mpc_engine *h, double max_width, double *h_ub, double *h_lb, double *h_border);

/* ReferencePath._compute_width / _get_min_width for T tracks in one launch (reference_path.py:206-287; SURVEY 8f-2: scenarios
 * that randomise the TRACK, not only the obstacles).  Track t has its own base map h_maps[t] (H x W int8, 1 free / 0 occupied,
 * row = world-y cell, as Map.data) and its own waypoints: h_tables[t] is the double[12][n_wp_max] table mpc_set_path takes,
 * of which the first h_n_wp[t] columns are valid.  All maps share H, W, origin and resolution.
 * Outputs (host, any may be NULL): h_ub / h_lb [T][n_wp_max] = Waypoint.ub / .lb, h_border [T][n_wp_max][4] = the static
 * border cells (ub_x, ub_y, lb_x, lb_y), h_err [T] = MPC_ST_INDEX_ERROR where a ray left the map (the reference raises
 * IndexError, reference_path.py:279).  Does not touch the engine's own path / grid. */
int mpc_compute_width_batch(mpc_engine *h, int32_t T, const int8_t *h_maps, int32_t H, int32_t W, double origin_x,
                            double origin_y, double resolution, const double *h_tables, const int32_t *h_n_wp,
                            int32_t n_wp_max, double max_width, double *h_ub, double *h_lb, double *h_border,
                            int32_t *h_err);

/* ---- K4 front: get_current_waypoint + t2s (sbm.py:256-279, 183-219) ------------------------- */
int mpc_localize_t2s(mpc_engine *h, const double *d_state, int32_t *d_wp_id, double *d_spatial,
                     int32_t *d_flags, int32_t B);

/* ---- K3: ReferencePath.update_path_constraints(wp_id+1, N, 2*sm, sm) (rp.py:522-648) -------- *
 * d_wp_id: current waypoint per scenario (the +1 of MPC.py:117 is applied inside).
 * d_cells_sm: optional double[B][N][4] border_cells_hor_sm (may be NULL). */
int mpc_raycast(mpc_engine *h, const int32_t *d_wp_id, double *d_ub, double *d_lb, double *d_cells_sm,
                int32_t *d_flags, int32_t B);
/* Generic form used by ReferencePath.update_path_constraints(wp_id, N, min_width, safety_margin):
 * first waypoint = wp_id[b] + first_offset, explicit widths. */
int mpc_update_path_constraints(mpc_engine *h, const int32_t *d_wp_id, int32_t first_offset, int32_t N,
                                double min_width, double safety_margin, double *d_ub, double *d_lb,
                                double *d_cells_sm, int32_t *d_flags, int32_t B);

/* ---- K1+K2: MPC._init_problem + OSQP setup/solve + control extraction -----------------------
 * (MPC.py:61-159, 183-220; sbm.py:391-417).  d_control in/out, d_infeas (int32[B]) in/out.
 * Optional outputs (NULL to skip): d_x_out, d_iters, d_qp_status. */
int mpc_assemble_solve(mpc_engine *h, const double *d_spatial, const int32_t *d_wp_id, double *d_control,
                       const double *d_ub, const double *d_lb, int32_t *d_infeas, double *d_u_out,
                       double *d_x_out, int32_t *d_iters, int32_t *d_qp_status, int32_t *d_flags,
                       int32_t B);

/* ---- K2 alone: B QPs in the reference's layout (QP-only sweep, oracle parity) ----------------
 * d_Pd[B][n] diagonal of P, d_q[B][n], d_Ax[B][16N+6] values of A in the fixed CSC pattern of
 * MPC.py:128-135 (column-major walk, structural zeros kept), d_l[B][m], d_u[B][m]; n = 5N+3,
 * m = 8N+6.  x_out NaN-filled for infeasible problems, like OSQP. */
int mpc_solve_qp(mpc_engine *h, const double *d_Pd, const double *d_q, const double *d_Ax,
                 const double *d_l, const double *d_u, double *d_x_out, int32_t *d_iters,
                 int32_t *d_qp_status, int32_t B);

/* ---- MPC.update_prediction (MPC.py:224-248) + s2t (sbm.py:155-181), batched -------------------
 * d_x_sol[B][5N+3]: solver output in dec.x order (mpc_assemble_solve's d_x_out); d_wp_id[B]: the waypoint each
 * car is localised at; d_xy_out[B][N-2][2]: world x, y of the predicted stages 2 .. N-1 (NaN past the end of a
 * non-circular path, where the reference exits). */
int mpc_predict_xy(mpc_engine *h, const double *d_x_sol, const int32_t *d_wp_id, double *d_xy_out, int32_t B);

/* ---- K4 back: BicycleModel.drive (sbm.py:221-244) ------------------------------------------ */
int mpc_rollout(mpc_engine *h, double *d_state, const double *d_spatial, const int32_t *d_wp_id,
                const double *d_u, const int32_t *d_flags, int32_t B);

/* ---- fused closed loop on engine-owned scenario state --------------------------------------- *
 * mpc_scenarios_init: allocates B scenarios, copies h_state double[4][B] (x, y, psi, s);
 * controls and infeasibility counters start at zero (MPC.py:53-56).
 * mpc_step: one get_control() + drive(u) for every live scenario (simulation.py:137-140): two fused
 * kernels (localise+raycast, assemble+solve+rollout); mpc_run_closed_loop replays them as a CUDA graph.
 * mpc_run_closed_loop: max_steps steps; h_stats double[8] = {scenario-steps, QP solves, ADMM
 * iterations, QP fallbacks, dead scenarios, finished scenarios, sum |e_y|, max |e_y|}. Synchronous. */
int mpc_scenarios_init(mpc_engine *h, const double *h_state, int32_t B);
int mpc_scenarios_set_state(mpc_engine *h, const double *h_state, const double *h_control,
                            const int32_t *h_infeas);
/* Restores the per-scenario step flags (MPC_ST_* bitmask, int32[B]) after mpc_scenarios_set_state, which clears them: a
 * caller that continues a fleet keeps its DEAD / FINISHED scenarios out of the loop (the reference's exit(1) at
 * MPC.py:218-220 and the end of `while car.s < length`, simulation.py:134, are final). */
int mpc_scenarios_set_flags(mpc_engine *h, const int32_t *h_flags);
int mpc_step(mpc_engine *h);
int mpc_run_closed_loop(mpc_engine *h, int32_t max_steps, double *h_stats);
/* host-buffer variant of one step (the e2e path): reads h_state[4][B], steps, leaves h_u_out[B][2], the new h_state
 * and (optionally) h_flags[B] in the caller's memory.  Synchronous. */
int mpc_step_host(mpc_engine *h, double *h_state, double *h_u_out, int32_t *h_flags);
/* mpc_step_host copies through an internal page-locked block when the caller's buffers are pageable.  When
 * they are page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory, or the engine's own block
 * returned here: state[4][B] | u[B][2] | flags[B], valid until the next mpc_scenarios_init) the step is one
 * CUDA-graph launch with no copies at all: the first kernel reads the state out of the caller's memory and the
 * solve kernel stores its results there (environment MPC_HOST_IO=copy: H2D node + kernels + D2H node instead;
 * same bits either way). */
int mpc_host_io(mpc_engine *h, double **h_state, double **h_u, int32_t **h_flags);
/* device views of the engine-owned scenario arrays (for zero-copy inspection from torch/ctypes) */
int mpc_scenarios_ptrs(mpc_engine *h, double **d_state, double **d_spatial, int32_t **d_wp_id,
                       double **d_control, double **d_ub, double **d_lb, double **d_u, int32_t **d_iters,
                       int32_t **d_qp_status, int32_t **d_flags, int32_t **d_infeas);
int mpc_scenarios_read(mpc_engine *h, double *h_state, double *h_control, double *h_u, int32_t *h_iters,
                       int32_t *h_qp_status, int32_t *h_flags, int32_t *h_infeas, int32_t *h_wp_id,
                       double *h_ub, double *h_lb);

/* ---- ReferencePath.compute_speed_profile (rp.py:289-354), one-off, synchronous, no engine needed ---
 * n = n_waypoints - 1 variables; h_li[n-1] distances between consecutive waypoints, h_vmax[n] the
 * curvature-limited speed bound (rp.py:329-331); cfg supplies the OSQP settings (NULL = defaults).
 * fp64 on the device; the iterates follow the reference solver's so that its eps = 1e-3 answer -- which
 * becomes v_ref -- is reproduced, not merely the exact minimiser. */
int mpc_speed_profile(const double *h_li, const double *h_vmax, int32_t n, double v_min, double a_min,
                      double a_max, const mpc_config *cfg, double *h_v_out, int32_t *h_iters,
                      int32_t *h_status);
/* The same for T tracks at once, one CTA per track (scenario sets that randomise the track).
 * h_off[T+1]: waypoint offsets of each track in the concatenated h_li / h_vmax / h_v_out arrays (h_off[0] = 0;
 * track t has n_t = h_off[t+1] - h_off[t] variables; h_li carries n_t - 1 lengths per track, the last slot of each
 * track is ignored).  h_iters / h_status: [T].  No limit on n_t: tracks whose working set exceeds shared memory
 * keep it in a global workspace. */
int mpc_speed_profile_batch(const double *h_li, const double *h_vmax, const int32_t *h_off, int32_t T,
                            double v_min, double a_min, double a_max, const mpc_config *cfg,
                            double *h_v_out, int32_t *h_iters, int32_t *h_status);

/* number of kernel launches enqueued by this engine since creation (bench.py's gpu_launches) */
int64_t mpc_launch_count(mpc_engine *h);
/* name + average device time of the engine's kernels measured with CUDA events inside
 * mpc_run_closed_loop when profiling is on: h_ms double[4] = K4a, K3, K1K2, K4b totals (ms). */
int mpc_set_profiling(mpc_engine *h, int32_t on);
int mpc_get_profile(mpc_engine *h, double *h_ms, int64_t *h_launches);

#ifdef __cplusplus
}
#endif
#endif /* MPC_B200_H */
