/*
 * oracle/mpc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (fp64, scalar C) restatement of the reference's per-step hot path
 *   MPC.get_control() -> BicycleModel.drive(u)
 * (reference: /root/reference/src/MPC.py:161-222, spatial_bicycle_models.py:221-279,
 *  reference_path.py:206-287,466-648, map.py:77-137) plus restatements of the two
 * un-vendored third-party algorithms the path calls into:
 *   - skimage.draw.line_aa   (scikit-image, unpinned; skimage/draw/_draw.pyx::_line_aa)
 *   - osqp.OSQP.setup/solve  (OSQP, unpinned, ~0.6.x; Stellato et al., Math. Prog. Comp. 2020)
 *
 * PARITY UNPINNED for those two: neither package nor its source is installed in the build
 * container and the reference ships no tests / golden vectors, so the restatements are written
 * from the published algorithms and validated only through (a) the reference's own Python control
 * flow run on top of them (oracle/make_golden.py) and (b) solver-independent certificates (KKT
 * residuals, line invariants).  Everything that is pure numpy in the reference is checked against
 * the reference's own code imported under shims (tests/test_oracle_vs_reference_live.py re-generates the fixtures when /root/reference is present; tests/test_oracle_cpu.py checks this file against them).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call into this file.  The product (multi-purpose-mpc_b200/) never does.
 *
 * Note on `**2`: CPython / numpy scalar `x ** 2` calls libm pow(x, 2.0).  glibc 2.39's pow is not
 * correctly rounded (about 0.1 % of squares differ from x*x by one ulp), so the reference's
 * distances are libm-dependent at the last bit.  orc_set_pow_mode(1) reproduces that (libm pow),
 * mode 0 uses IEEE x*x.  Cells and every index decision are identical in both modes on all
 * fixtures; widths can differ by 1 ulp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_PI 3.141592653589793 /* math.pi */

static int g_pow_mode = 0; /* 0: x*x (IEEE), 1: libm pow(x,2.0) as CPython/numpy do */
void orc_set_pow_mode(int m) { g_pow_mode = m; }
/* through a volatile pointer: gcc folds a literal pow(x, 2.0) into x*x at -O2, which is exactly the
 * substitution this switch exists to expose */
static double (*volatile libm_pow)(double, double) = pow;
static inline double sq(double x) { return g_pow_mode ? libm_pow(x, 2.0) : x * x; }

/* numpy floor-mod for doubles (np.mod), b > 0 here */
static inline double np_mod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* Map.w2m / Map.m2w  (map.py:77-101)                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int H, W;
    const int8_t *data; /* [H][W], 1 free / 0 occupied, row = y cell */
    double ox, oy, res;
} orc_map;

static inline void w2m(const orc_map *mp, double x, double y, long *dx, long *dy)
{
    *dx = (long)floor((x - mp->ox) / mp->res); /* map.py:85 */
    *dy = (long)floor((y - mp->oy) / mp->res); /* map.py:86 */
}
static inline void m2w(const orc_map *mp, long dx, long dy, double *x, double *y)
{
    *x = ((double)dx + 0.5) * mp->res + mp->ox; /* map.py:98 */
    *y = ((double)dy + 0.5) * mp->res + mp->oy; /* map.py:99 */
}
void orc_w2m(double ox, double oy, double res, double x, double y, long *out)
{
    orc_map m = {0, 0, NULL, ox, oy, res};
    w2m(&m, x, y, &out[0], &out[1]);
}
void orc_m2w(double ox, double oy, double res, long dx, long dy, double *out)
{
    orc_map m = {0, 0, NULL, ox, oy, res};
    m2w(&m, dx, dy, &out[0], &out[1]);
}

/* numpy-style indexing data[y, x]: negative indices wrap once, otherwise out of range = error */
static inline int map_at(const orc_map *mp, long x, long y, int *err)
{
    if (x < 0) x += mp->W;
    if (y < 0) y += mp->H;
    if (x < 0 || y < 0 || x >= mp->W || y >= mp->H) { *err = 1; return 0; }
    return mp->data[y * (long)mp->W + x];
}

/* Map.add_obstacles for one disc (map.py:126-137).  data is modified in place.
 * Window [c-r, c+r) in both axes (np.ogrid[-r:r], quirk Q5), test dx^2+dy^2 <= r^2.
 * numpy slice semantics for the window are reproduced: the slice bounds are clipped the way
 * data[a:b] clips them (negative start wraps), and the boolean mask must match the window shape,
 * so the caller is expected to keep discs radius_px inside the map; otherwise returns -1. */
int orc_add_obstacle(int8_t *data, int H, int W, double ox, double oy, double res,
                     double cx, double cy, double radius)
{
    orc_map m = {H, W, data, ox, oy, res};
    long r = (long)ceil(radius / res); /* map.py:129 */
    long cxp, cyp;
    w2m(&m, cx, cy, &cxp, &cyp);
    if (cxp - r < 0 || cyp - r < 0 || cxp + r > W || cyp + r > H) return -1;
    for (long dy = -r; dy < r; ++dy)
        for (long dx = -r; dx < r; ++dx)
            if (dx * dx + dy * dy <= r * r) data[(cyp + dy) * (long)W + (cxp + dx)] = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* skimage.draw.line_aa cell sequence (skimage/draw/_draw.pyx::_line_aa; values not needed).   */
/* The reference calls line_aa(x0, y0, x1, y1): skimage's r is x and c is y (rp.py:268,484).   */
/* err and ed are C floats in skimage; kept as floats here.                                    */
/* ------------------------------------------------------------------------------------------ */
int orc_line_aa(long r0, long c0, long r1, long c1, long *rr, long *cc, int cap)
{
    int n = 0;
    int dc = (int)labs(c0 - c1);
    int dr = (int)labs(r0 - r1);
    float err = (float)(dc - dr);
    float err_prime;
    long c, r, c_prime;
    int sign_c = (c0 < c1) ? 1 : -1;
    int sign_r = (r0 < r1) ? 1 : -1;
    float ed;
    if (dc + dr == 0) ed = 1.0f;
    else ed = (float)sqrt((double)(dc * dc + dr * dr));
    c = c0; r = r0;
    for (;;) {
        if (n >= cap) return -1;
        cc[n] = c; rr[n] = r; ++n;
        err_prime = err;
        c_prime = c;
        if (2 * err_prime >= -dc) {
            if (c == c1) break;
            if (err_prime + dr < ed) {
                if (n >= cap) return -1;
                cc[n] = c; rr[n] = r + sign_r; ++n;
            }
            err -= dr;
            c += sign_c;
        }
        if (2 * err_prime <= dr) {
            if (r == r1) break;
            if (dc - err_prime < ed) {
                if (n >= cap) return -1;
                cc[n] = c_prime + sign_c; rr[n] = r; ++n;
            }
            err += dc;
            r += sign_r;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Path tables.  Everything transcendental is a per-waypoint table computed by the caller with   */
/* numpy, using the reference's own expressions, so this file never disagrees with numpy at the  */
/* ulp level on cos/sin.                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n_wp;
    int circular;
    const double *x, *y, *psi, *kappa, *v_ref;
    const double *seg_len;  /* reference_path.segment_lengths: [0.0, d(0,1), d(1,2), ...] (rp.py:201) */
    const double *ds_next;  /* ds_next[k] = get_waypoint(k+1) - get_waypoint(k), circular wrap included */
    const double *cos_psi, *sin_psi; /* np.cos(wp.psi), np.sin(wp.psi) */
    const double *cos_ub, *sin_ub;   /* of angle_ub = mod(pi/2 + psi + pi, 2pi) - pi (rp.py:622) */
    const double *cos_lb, *sin_lb;   /* of angle_lb = mod(-pi/2 + psi + pi, 2pi) - pi (rp.py:624) */
    const double *border;   /* [n_wp][4]: static_border_cells (ub_x, ub_y, lb_x, lb_y) world coords */
} orc_path;

#define ORC_MAX_LINE 4096
#define ORC_MAX_SEG 64

/* status codes shared with the product's header (include/mpc_b200.h) */
enum { ORC_OK = 0, ORC_NO_SEGMENT = 1, ORC_END_OF_PATH = 2, ORC_INDEX_ERROR = 3 };

/* ReferencePath._compute_free_segments (rp.py:466-520).
 * Returns number of segments; seg[i] = {ub_x, ub_y, lb_x, lb_y} in world coordinates.
 * If cells_out != NULL the visited cells (x,y) are appended (for the sector statistics). */
static int free_segments(const orc_map *mp, const double *border4, double min_width,
                         double seg[][4], int *err, long *cells_out, int *n_cells_out)
{
    long ubx, uby, lbx, lby;
    static __thread long rr[ORC_MAX_LINE], cc[ORC_MAX_LINE];
    w2m(mp, border4[0], border4[1], &ubx, &uby); /* rp.py:478 */
    w2m(mp, border4[2], border4[3], &lbx, &lby); /* rp.py:480 */
    int n = orc_line_aa(ubx, uby, lbx, lby, rr, cc, ORC_MAX_LINE); /* rp.py:484: rr->x, cc->y */
    if (n < 0) { *err = 1; return 0; }
    long ub_o[2] = {ubx, uby}, lb_o[2] = {ubx, uby};
    (void)lb_o;
    int free_cells = 0, nseg = 0;
    for (int i = 1; i < n; ++i) { /* rp.py:494 skips the first cell (Q4) */
        long x = rr[i], y = cc[i];
        if (cells_out) { cells_out[2 * (*n_cells_out)] = x; cells_out[2 * (*n_cells_out) + 1] = y; ++*n_cells_out; }
        int v = map_at(mp, x, y, err);
        if (*err) return 0;
        if (v == 1) { free_cells = 1; lb_o[0] = x; lb_o[1] = y; }
        if ((v == 0 || (x == lbx && y == lby)) && free_cells) {
            double ux, uy, lx, ly;
            m2w(mp, ub_o[0], ub_o[1], &ux, &uy);
            m2w(mp, x, y, &lx, &ly);
            if (sqrt(sq(ux - lx) + sq(uy - ly)) > min_width) { /* rp.py:510 */
                if (nseg >= ORC_MAX_SEG) { *err = 1; return 0; }
                seg[nseg][0] = ux; seg[nseg][1] = uy; seg[nseg][2] = lx; seg[nseg][3] = ly;
                ++nseg;
            }
            ub_o[0] = x; ub_o[1] = y;
            free_cells = 0;
        } else if (v == 0 && !free_cells) {
            ub_o[0] = x; ub_o[1] = y; lb_o[0] = x; lb_o[1] = y;
        }
    }
    return nseg;
}

static inline int wrap_wp(const orc_path *p, long id, int *status)
{
    if (id >= p->n_wp) {
        if (p->circular) id = id % p->n_wp; /* rp.py:364-365 */
        else { *status = ORC_END_OF_PATH; return 0; } /* rp.py:367-369 exit(1) */
    }
    return (int)id;
}

static inline double sign_of(double a) { return (a > 0) - (a < 0); }

/* ReferencePath.update_path_constraints (rp.py:522-648).
 * ub[N], lb[N]; cells_sm[N][4] = border_cells_hor_sm; sectors (optional) = number of distinct
 * 32-byte sectors of a bit-packed grid with 64-byte row pitch that contain a tested cell. */
int orc_update_path_constraints(const int8_t *data, int H, int W, double ox, double oy, double res,
                                const orc_path *p, long wp_id, int N, double min_width,
                                double safety_margin, double *ub_out, double *lb_out,
                                double *cells_sm, long *n_cells_tested, long *n_sectors)
{
    orc_map mp = {H, W, data, ox, oy, res};
    int status = ORC_OK, err = 0;
    double prev_cells[4] = {0, 0, 0, 0};
    double seg[ORC_MAX_SEG][4];
    long *cells = NULL;
    int ncells = 0;
    if (n_sectors || n_cells_tested) cells = (long *)malloc(sizeof(long) * 2 * ORC_MAX_LINE * (size_t)N);
    for (int n = 0; n < N; ++n) {
        int k = wrap_wp(p, wp_id + n, &status);
        if (status) break;
        int nseg = free_segments(&mp, p->border + 4 * k, min_width, seg, &err, cells, &ncells);
        if (err) { status = ORC_INDEX_ERROR; break; }
        double ubx, uby, lbx, lby;
        if (n == 0) {
            if (nseg == 0) { status = ORC_NO_SEGMENT; break; } /* rp.py:547 max([]) -> ValueError */
            int best = 0; double bestl = -1.0;
            for (int i = 0; i < nseg; ++i) {
                double l = sqrt(sq(seg[i][0] - seg[i][2]) + sq(seg[i][1] - seg[i][3]));
                if (l > bestl) { bestl = l; best = i; } /* list.index(max()) = first maximum */
            }
            ubx = seg[best][0]; uby = seg[best][1]; lbx = seg[best][2]; lby = seg[best][3];
        } else {
            int kp = wrap_wp(p, wp_id + n - 1, &status);
            double ds = p->ds_next[kp]; /* wp_prev - wp (rp.py:558); same value as wp - wp_prev */
            double upx = prev_cells[0] + ds * p->cos_psi[kp]; /* rp.py:559 */
            double upy = prev_cells[1] + ds * p->cos_psi[kp]; /* rp.py:560 (quirk Q2) */
            double lpx = prev_cells[2] + ds * p->sin_psi[kp]; /* rp.py:561 */
            double lpy = prev_cells[3] + ds * p->sin_psi[kp]; /* rp.py:562 */
            if (nseg >= 2) {
                int best = 0; double bestd = INFINITY;
                for (int i = 0; i < nseg; ++i) {
                    double d_ub = sqrt(sq(seg[i][0] - upx) + sq(seg[i][1] - upy));
                    double d_lb = sqrt(sq(seg[i][2] - lpx) + sq(seg[i][3] - lpy));
                    double md = (d_ub + d_lb) / 2;
                    if (md < bestd) { bestd = md; best = i; }
                }
                ubx = seg[best][0]; uby = seg[best][1]; lbx = seg[best][2]; lby = seg[best][3];
            } else if (nseg == 1) {
                ubx = seg[0][0]; uby = seg[0][1]; lbx = seg[0][2]; lby = seg[0][3];
            } else {
                ubx = p->x[k]; uby = p->y[k]; lbx = p->x[k]; lby = p->y[k]; /* rp.py:595 */
            }
        }
        double wx = p->x[k], wy = p->y[k], psi = p->psi[k];
        double angle_ub = np_mod(atan2(uby - wy, ubx - wx) - psi + ORC_PI, 2 * ORC_PI) - ORC_PI;
        double angle_lb = np_mod(atan2(lby - wy, lbx - wx) - psi + ORC_PI, 2 * ORC_PI) - ORC_PI;
        double ub = sign_of(angle_ub) * sqrt(sq(ubx - wx) + sq(uby - wy)); /* rp.py:606 */
        double lb = sign_of(angle_lb) * sqrt(sq(lbx - wx) + sq(lby - wy)); /* rp.py:608 */
        ub -= safety_margin;
        lb += safety_margin;
        if (ub < lb) { ub = 0.0; lb = 0.0; }
        double cu = p->cos_ub[k], su = p->sin_ub[k], cl = p->cos_lb[k], sl = p->sin_lb[k];
        if (cells_sm) {
            cells_sm[4 * n + 0] = wx + ub * cu; cells_sm[4 * n + 1] = wy + ub * su; /* rp.py:627 */
            cells_sm[4 * n + 2] = wx - lb * cl; cells_sm[4 * n + 3] = wy - lb * sl; /* rp.py:629 */
        }
        prev_cells[0] = wx + (ub + safety_margin) * cu; /* rp.py:633 */
        prev_cells[1] = wy + (ub + safety_margin) * su;
        prev_cells[2] = wx - (lb - safety_margin) * cl; /* rp.py:635 */
        prev_cells[3] = wy - (lb - safety_margin) * sl;
        ub_out[n] = ub;
        lb_out[n] = lb;
    }
    if (cells) {
        if (n_cells_tested) *n_cells_tested = ncells;
        if (n_sectors) {
            /* distinct (row, 32-byte sector) pairs: sector = x / 256, two per 64-byte row */
            long cnt = 0;
            uint8_t *seen = (uint8_t *)calloc((size_t)H * 2 + 2, 1);
            for (int i = 0; i < ncells; ++i) {
                long x = cells[2 * i], y = cells[2 * i + 1];
                if (x < 0 || y < 0 || y >= H || x >= 512) continue;
                long key = y * 2 + (x >> 8);
                if (!seen[key]) { seen[key] = 1; ++cnt; }
            }
            free(seen);
            *n_sectors = cnt;
        }
        free(cells);
    }
    return status;
}

/* ReferencePath._compute_width + _get_min_width (rp.py:206-287) for all waypoints.
 * Needs per-waypoint cos/sin of the left/right angles (tables by the caller, numpy):
 *   left : mod(psi + pi/2 + pi, 2pi) - pi ; right: mod(psi - pi/2 + pi, 2pi) - pi  (= cos_ub/.. tables)
 * out_ub[n_wp], out_lb[n_wp], out_border[n_wp][4]. */
int orc_compute_width(const int8_t *data, int H, int W, double ox, double oy, double res,
                      const orc_path *p, double max_width, double *out_ub, double *out_lb,
                      double *out_border)
{
    orc_map mp = {H, W, data, ox, oy, res};
    static __thread long rr[ORC_MAX_LINE], cc[ORC_MAX_LINE];
    int err = 0;
    for (int k = 0; k < p->n_wp; ++k) {
        double wx = p->x[k], wy = p->y[k];
        double b_value[2], b_cell[2][2];
        for (int side = 0; side < 2; ++side) {
            double ca = side == 0 ? p->cos_ub[k] : p->cos_lb[k];
            double sa = side == 0 ? p->sin_ub[k] : p->sin_lb[k];
            long tx, ty, wpx, wpy;
            w2m(&mp, wx + max_width * ca, wy + max_width * sa, &tx, &ty); /* rp.py:227 */
            w2m(&mp, wx, wy, &wpx, &wpy);                                  /* rp.py:263 */
            double min_w = max_width;
            double mcx, mcy;
            m2w(&mp, tx + 1, ty + 1, &mcx, &mcy); /* rp.py:274 with the leaked loop variables (Q3) */
            for (int i = -1; i <= 1; ++i)
                for (int j = -1; j <= 1; ++j) {
                    int n = orc_line_aa(wpx, wpy, tx + i, ty + j, rr, cc, ORC_MAX_LINE);
                    if (n < 0) return ORC_INDEX_ERROR;
                    for (int q = 0; q < n; ++q) {
                        int v = map_at(&mp, rr[q], cc[q], &err);
                        if (err) return ORC_INDEX_ERROR;
                        if (v == 0) {
                            double cx, cy;
                            m2w(&mp, rr[q], cc[q], &cx, &cy);
                            double d = sqrt(sq(wx - cx) + sq(wy - cy)); /* rp.py:282 */
                            if (d < min_w) { min_w = d; mcx = cx; mcy = cy; }
                        }
                    }
                }
            b_value[side] = min_w; b_cell[side][0] = mcx; b_cell[side][1] = mcy;
        }
        out_ub[k] = b_value[0];
        out_lb[k] = -1 * b_value[1];
        out_border[4 * k + 0] = b_cell[0][0]; out_border[4 * k + 1] = b_cell[0][1];
        out_border[4 * k + 2] = b_cell[1][0]; out_border[4 * k + 3] = b_cell[1][1];
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* SpatialBicycleModel.get_current_waypoint / t2s / drive (sbm.py:183-279)                      */
/* ------------------------------------------------------------------------------------------ */
/* length_cum = np.cumsum(segment_lengths) is passed in (numpy's cumsum order of additions). */
int orc_get_current_waypoint(const double *length_cum, int n_wp, double s)
{
    /* greater_than_threshold.searchsorted(True): first index with length_cum > s; if none the
     * array is all False and searchsorted returns n_wp -> IndexError in the reference. */
    int next = n_wp;
    for (int i = 0; i < n_wp; ++i) if (length_cum[i] > s) { next = i; break; }
    if (next >= n_wp) return -1;
    int prev = next - 1; /* -1 wraps to the last element in numpy */
    double s_next = length_cum[next];
    double s_prev = length_cum[prev < 0 ? n_wp - 1 : prev];
    if (fabs(s - s_next) < fabs(s - s_prev)) return next;
    return prev < 0 ? n_wp - 1 : prev; /* waypoints[-1] */
}

void orc_t2s(double x, double y, double psi, double wx, double wy, double wpsi, double cos_wpsi,
             double sin_wpsi, double *out3)
{
    double e_y = cos_wpsi * (y - wy) - sin_wpsi * (x - wx);        /* sbm.py:202-205 */
    double e_psi = psi - wpsi;
    e_psi = np_mod(e_psi + ORC_PI, 2 * ORC_PI) - ORC_PI;           /* sbm.py:209 */
    out3[0] = e_y; out3[1] = e_psi; out3[2] = 0.0;                 /* sbm.py:217 */
}

/* state4 = (x, y, psi, s) updated in place. */
void orc_drive(double *state4, double e_y, double e_psi, double kappa_wp, double v, double delta,
               double L, double Ts)
{
    double psi = state4[2];
    double x_dot = v * cos(psi);            /* sbm.py:231 */
    double y_dot = v * sin(psi);            /* sbm.py:232 */
    double psi_dot = v / L * tan(delta);    /* sbm.py:233 */
    state4[0] += x_dot * Ts;
    state4[1] += y_dot * Ts;
    state4[2] += psi_dot * Ts;
    double s_dot = 1 / (1 - e_y * kappa_wp) * v * cos(e_psi); /* sbm.py:240 */
    state4[3] += s_dot * Ts;
}

/* ------------------------------------------------------------------------------------------ */
/* MPC._init_problem (MPC.py:61-155): the QP in the reference's layout.                          */
/* Variables [x0..xN (3 each) | u0..uN-1 (2 each)], n = 5N+3.  Rows: 3(N+1) equalities then n    */
/* identity rows, m = 8N+6.  A is emitted in CSC with the FIXED structural pattern (zeros kept,  */
/* SURVEY H8): nnz = 16N+6.  P is diagonal (Pd).                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int N;
    double Q[3], R[2], QN[3];
    double xmin[3], xmax[3], umin[2], umax[2];
    double ay_max, L, safety_margin;
} orc_mpc_cfg;

int orc_mpc_nnz(int N) { return 16 * N + 6; }

int orc_mpc_assemble(const orc_path *p, const orc_mpc_cfg *c, long wp_id, const double *x0,
                     const double *current_control /*2N*/, const double *ub, const double *lb,
                     double *Pd, double *q, int *Ap, int *Ai, double *Ax, double *l, double *u)
{
    const int N = c->N, nx = 3, nu = 2;
    const int n = nx * (N + 1) + nu * N, neq = nx * (N + 1);
    int status = ORC_OK;
    double *Alin = (double *)calloc((size_t)N * 9, sizeof(double));
    double *Blin = (double *)calloc((size_t)N * 6, sizeof(double));
    double *ur = (double *)calloc((size_t)nu * N, sizeof(double));
    double *xr = (double *)calloc((size_t)nx * (N + 1), sizeof(double));
    double *uq = (double *)calloc((size_t)nx * N, sizeof(double));
    double *xmin_dyn = (double *)malloc(sizeof(double) * nx * (N + 1));
    double *xmax_dyn = (double *)malloc(sizeof(double) * nx * (N + 1));
    double *umax_dyn = (double *)malloc(sizeof(double) * nu * N);
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nx; ++i) { xmin_dyn[nx * k + i] = c->xmin[i]; xmax_dyn[nx * k + i] = c->xmax[i]; }
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nu; ++i) umax_dyn[nu * k + i] = c->umax[i];
    for (int k = 0; k < N; ++k) {
        int w0 = wrap_wp(p, wp_id + k, &status);
        if (status) goto done;
        double delta_s = p->ds_next[w0]; /* next_waypoint - current_waypoint (MPC.py:95) */
        double kappa_ref = p->kappa[w0], v_ref = p->v_ref[w0];
        /* BicycleModel.linearize (sbm.py:404-412) */
        double *A = Alin + 9 * k, *B = Blin + 6 * k;
        A[0] = 1; A[1] = delta_s; A[2] = 0;
        A[3] = -sq(kappa_ref) * delta_s; A[4] = 1; A[5] = 0;
        A[6] = -kappa_ref / v_ref * delta_s; A[7] = 0; A[8] = 1;
        B[0] = 0; B[1] = 0; B[2] = 0; B[3] = delta_s;
        B[4] = -1 / sq(v_ref) * delta_s; B[5] = 0;
        double f[3] = {0.0, 0.0, 1 / v_ref * delta_s};
        ur[nu * k] = v_ref; ur[nu * k + 1] = kappa_ref;
        for (int i = 0; i < 3; ++i) /* B_lin.dot([v_ref, kappa_ref]) - f (MPC.py:107) */
            uq[nx * k + i] = (B[2 * i] * v_ref + B[2 * i + 1] * kappa_ref) - f[i];
        /* kappa_pred[n] = tan(cc[3+n] + cc[2N-1]) / L  (MPC.py:86, quirk Q1) */
        double kp = tan(current_control[3 + k] + current_control[2 * N - 1]) / c->L;
        /* only the first 2N-3 entries exist; for k >= 2N-3 the reference would raise IndexError,
         * which cannot happen because k < N <= 2N-3 for N >= 3 */
        double vmax_dyn = sqrt(c->ay_max / (fabs(kp) + 1e-12));
        if (vmax_dyn < umax_dyn[nu * k]) umax_dyn[nu * k] = vmax_dyn;
    }
    xmin_dyn[0] = x0[0]; xmax_dyn[0] = x0[0]; /* MPC.py:119-120 (quirk Q6) */
    for (int k = 0; k < N; ++k) {
        xmin_dyn[nx * (k + 1)] = lb[k];
        xmax_dyn[nx * (k + 1)] = ub[k];
        xr[nx * (k + 1)] = (lb[k] + ub[k]) / 2; /* MPC.py:125 */
    }
    /* CSC of A = [[Ax_blocks, Bu],[I]] with Ax = kron(I, -I) + A_lin sub-diagonal blocks */
    {
        int nz = 0;
        for (int col = 0; col < n; ++col) {
            Ap[col] = nz;
            if (col < neq) {
                int k = col / nx, j = col % nx;
                Ai[nz] = col; Ax[nz] = -1.0; ++nz;           /* -I */
                if (k < N) {
                    /* column j of A_lin[k] lands in rows nx*(k+1) .. ; fixed pattern:
                     * col0: rows 0,1,2 ; col1: rows 0,1 ; col2: row 2 */
                    const double *A = Alin + 9 * k;
                    if (j == 0) {
                        Ai[nz] = nx * (k + 1) + 0; Ax[nz] = A[0]; ++nz;
                        Ai[nz] = nx * (k + 1) + 1; Ax[nz] = A[3]; ++nz;
                        Ai[nz] = nx * (k + 1) + 2; Ax[nz] = A[6]; ++nz;
                    } else if (j == 1) {
                        Ai[nz] = nx * (k + 1) + 0; Ax[nz] = A[1]; ++nz;
                        Ai[nz] = nx * (k + 1) + 1; Ax[nz] = A[4]; ++nz;
                    } else {
                        Ai[nz] = nx * (k + 1) + 2; Ax[nz] = A[8]; ++nz;
                    }
                }
            } else {
                int k = (col - neq) / nu, j = (col - neq) % nu;
                const double *B = Blin + 6 * k;
                if (j == 0) { Ai[nz] = nx * (k + 1) + 2; Ax[nz] = B[4]; ++nz; }
                else        { Ai[nz] = nx * (k + 1) + 1; Ax[nz] = B[3]; ++nz; }
            }
            Ai[nz] = neq + col; Ax[nz] = 1.0; ++nz;          /* identity (bounds) row */
        }
        Ap[n] = nz;
    }
    for (int i = 0; i < nx; ++i) { l[i] = -x0[i]; u[i] = -x0[i]; }          /* MPC.py:142-144 */
    for (int i = 0; i < nx * N; ++i) { l[nx + i] = uq[i]; u[nx + i] = uq[i]; }
    for (int i = 0; i < nx * (N + 1); ++i) { l[neq + i] = xmin_dyn[i]; u[neq + i] = xmax_dyn[i]; }
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nu; ++i) {
            l[neq + nx * (N + 1) + nu * k + i] = c->umin[i];
            u[neq + nx * (N + 1) + nu * k + i] = umax_dyn[nu * k + i];
        }
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nx; ++i) { Pd[nx * k + i] = c->Q[i]; q[nx * k + i] = -c->Q[i] * xr[nx * k + i]; }
    for (int i = 0; i < nx; ++i) { Pd[nx * N + i] = c->QN[i]; q[nx * N + i] = -(c->QN[i] * xr[nx * N + i]); }
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nu; ++i) {
            Pd[neq + nu * k + i] = c->R[i];
            q[neq + nu * k + i] = -c->R[i] * ur[nu * k + i];
        }
done:
    free(Alin); free(Blin); free(ur); free(xr); free(uq); free(xmin_dyn); free(xmax_dyn); free(umax_dyn);
    return status;
}
