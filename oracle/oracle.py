"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the CPU oracle (oracle/liborc.so, built by oracle/Makefile) and the numpy
table builders the oracle's C functions expect.  Every table is computed with the reference's own
numpy expression (file:line cited) so that nothing here can disagree with numpy at the ulp level.

PARITY UNPINNED for the two third-party algorithms restated in C (skimage.draw.line_aa, OSQP):
see the headers of mpc_oracle.c / osqp_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)
c_i8_p = C.POINTER(C.c_int8)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("mpc_oracle.c", "osqp_oracle.c", "orc_batch.c")]
    if force or not os.path.exists(so) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class OrcPath(C.Structure):
    _fields_ = [("n_wp", C.c_int), ("circular", C.c_int)] + [
        (k, c_double_p) for k in ("x", "y", "psi", "kappa", "v_ref", "seg_len", "ds_next", "cos_psi",
                                  "sin_psi", "cos_ub", "sin_ub", "cos_lb", "sin_lb", "border")]


class OrcMpcCfg(C.Structure):
    _fields_ = [("N", C.c_int), ("Q", C.c_double * 3), ("R", C.c_double * 2), ("QN", C.c_double * 3),
                ("xmin", C.c_double * 3), ("xmax", C.c_double * 3), ("umin", C.c_double * 2),
                ("umax", C.c_double * 2), ("ay_max", C.c_double), ("L", C.c_double),
                ("safety_margin", C.c_double)]


class OrcOsqpSettings(C.Structure):
    _fields_ = [("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double),
                ("eps_dual_inf", C.c_double), ("max_iter", C.c_int), ("scaling", C.c_int),
                ("check_termination", C.c_int), ("adaptive_rho", C.c_int),
                ("adaptive_rho_interval", C.c_int), ("adaptive_rho_tolerance", C.c_double)]


class OrcWorld(C.Structure):
    _fields_ = [("path", OrcPath), ("length_cum", c_double_p), ("cfg", OrcMpcCfg),
                ("osqp", OrcOsqpSettings), ("H", C.c_int), ("W", C.c_int), ("ox", C.c_double),
                ("oy", C.c_double), ("res", C.c_double), ("Ts", C.c_double)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_line_aa.restype = C.c_int
        L.orc_line_aa.argtypes = [C.c_long] * 4 + [c_long_p, c_long_p, C.c_int]
        L.orc_w2m.argtypes = [C.c_double] * 5 + [c_long_p]
        L.orc_m2w.argtypes = [C.c_double] * 3 + [C.c_long, C.c_long, c_double_p]
        L.orc_add_obstacle.restype = C.c_int
        L.orc_add_obstacle.argtypes = [c_i8_p, C.c_int, C.c_int] + [C.c_double] * 6
        L.orc_update_path_constraints.restype = C.c_int
        L.orc_update_path_constraints.argtypes = [
            c_i8_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(OrcPath), C.c_long,
            C.c_int, C.c_double, C.c_double, c_double_p, c_double_p, c_double_p, c_long_p, c_long_p]
        L.orc_compute_width.restype = C.c_int
        L.orc_compute_width.argtypes = [c_i8_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.POINTER(OrcPath), C.c_double, c_double_p, c_double_p, c_double_p]
        L.orc_get_current_waypoint.restype = C.c_int
        L.orc_get_current_waypoint.argtypes = [c_double_p, C.c_int, C.c_double]
        L.orc_t2s.argtypes = [C.c_double] * 8 + [c_double_p]
        L.orc_drive.argtypes = [c_double_p] + [C.c_double] * 7
        L.orc_mpc_assemble.restype = C.c_int
        L.orc_mpc_assemble.argtypes = [C.POINTER(OrcPath), C.POINTER(OrcMpcCfg), C.c_long, c_double_p,
                                       c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                       c_int_p, c_int_p, c_double_p, c_double_p, c_double_p]
        L.orc_osqp_default_settings.argtypes = [C.POINTER(OrcOsqpSettings)]
        L.orc_osqp_solve.restype = C.c_int
        L.orc_osqp_solve.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p, c_int_p,
                                     c_int_p, c_double_p, c_double_p, c_double_p,
                                     C.POINTER(OrcOsqpSettings), c_int_p, c_double_p, c_double_p,
                                     c_int_p, c_double_p]
        L.orc_mpc_step.restype = C.c_int
        L.orc_mpc_step.argtypes = [C.POINTER(OrcWorld), c_i8_p, c_double_p, c_double_p, c_int_p, C.c_int,
                                   c_double_p, c_double_p, c_double_p, c_double_p, c_int_p, c_int_p, c_int_p]
        L.orc_batch_closed_loop.restype = C.c_int
        L.orc_batch_closed_loop.argtypes = [C.POINTER(OrcWorld), c_i8_p, C.c_long, C.c_int, C.c_int,
                                            c_double_p, c_double_p, c_int_p, c_int_p, c_double_p, C.c_int]
        L.orc_batch_qp_solve.restype = C.c_int
        L.orc_batch_qp_solve.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_int_p, c_int_p,
                                         c_double_p, c_double_p, c_double_p, C.POINTER(OrcOsqpSettings),
                                         c_double_p, c_int_p, c_int_p, C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_pow_mode.argtypes = [C.c_int]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def set_pow_mode(libm_pow: bool):
    """True: `**2` = libm pow(x, 2.0) exactly as CPython/numpy evaluate it; False: IEEE x*x."""
    lib().orc_set_pow_mode(1 if libm_pow else 0)


# --------------------------------------------------------------------------------------------
# third-party restatements exposed with the third-party call signatures
# --------------------------------------------------------------------------------------------
def line_aa(r0, c0, r1, c1):
    """skimage.draw.line_aa(r0, c0, r1, c1) -> (rr, cc, val); val is not restated (unused by the
    reference, rp.py:268,484; map.py:153) and returned as ones."""
    cap = 4 * (abs(int(r1) - int(r0)) + abs(int(c1) - int(c0))) + 8
    rr = np.empty(cap, dtype=np.int64)
    cc = np.empty(cap, dtype=np.int64)
    n = lib().orc_line_aa(int(r0), int(c0), int(r1), int(c1), rr.ctypes.data_as(c_long_p),
                          cc.ctypes.data_as(c_long_p), cap)
    assert n > 0
    return rr[:n].astype(np.intp), cc[:n].astype(np.intp), np.ones(n)


def remove_small_holes(ar, area_threshold=64, connectivity=1):
    """skimage.morphology.remove_small_holes for 2-D input: fill background components smaller
    than area_threshold.  connectivity >= ndim means the full (8-connected) structuring element."""
    from scipy import ndimage as ndi
    a = np.asarray(ar).astype(bool)
    conn = min(int(connectivity), a.ndim)
    footprint = ndi.generate_binary_structure(a.ndim, conn)
    lab, _ = ndi.label(~a, structure=footprint)
    sizes = np.bincount(lab.ravel())
    too_small = sizes < area_threshold
    too_small[0] = False
    out = a.copy()
    out[too_small[lab]] = True
    return out


OSQP_STATUS = {1: "solved", 2: "solved inaccurate", -2: "maximum iterations reached",
               -3: "primal infeasible", 3: "primal infeasible inaccurate", -4: "dual infeasible",
               4: "dual infeasible inaccurate", -7: "problem non convex", -100: "oracle factor failed"}


def default_settings(**kw):
    s = OrcOsqpSettings()
    lib().orc_osqp_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


def osqp_solve(P, q, A, l, u, perm=None, **settings):
    """Solve with the OSQP restatement.  P, A: scipy sparse (any format).  Returns dict."""
    from scipy import sparse
    P = sparse.triu(sparse.csc_matrix(P), format="csc")
    A = sparse.csc_matrix(A)
    P.sort_indices()
    A.sort_indices()
    n, m = P.shape[0], A.shape[0]
    q = np.ascontiguousarray(q, dtype=np.float64)
    l = np.ascontiguousarray(l, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    Pp, Pi, Px = P.indptr.astype(np.int32), P.indices.astype(np.int32), P.data.astype(np.float64)
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    s = default_settings(**settings)
    x = np.empty(n)
    y = np.empty(m)
    it = C.c_int(0)
    info = np.zeros(8)
    permp = None
    if perm is not None:
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        permp = _ip(perm)
    st = lib().orc_osqp_solve(n, m, _ip(Pp), _ip(Pi), _dp(Px), _dp(q), _ip(Ap), _ip(Ai), _dp(Ax), _dp(l),
                              _dp(u), C.byref(s), permp, _dp(x), _dp(y), C.byref(it), _dp(info))
    return dict(x=x, y=y, status=st, status_str=OSQP_STATUS.get(st, "?"), iter=it.value, pri_res=info[0],
                dua_res=info[1], obj=info[2], rho=info[3], rho_updates=int(info[4]),
                n_factor=int(info[5]), bandwidth=int(info[7]))


def stage_perm(N):
    """perm[new] = old for the ordering [x0 u0 x1 u1 ... xN] of the reference's [x.. | u..] layout."""
    p = []
    for s in range(N + 1):
        p += [3 * s, 3 * s + 1, 3 * s + 2]
        if s < N:
            p += [3 * (N + 1) + 2 * s, 3 * (N + 1) + 2 * s + 1]
    return np.array(p, dtype=np.int32)


def kkt_residuals(P, q, A, l, u, x, y):
    """Solver-independent certificate (fp64): primal residual, dual residual, complementarity."""
    Ax = A @ x
    pri = max(np.max(np.maximum(l - Ax, 0), initial=0.0), np.max(np.maximum(Ax - u, 0), initial=0.0))
    dua = np.max(np.abs(P @ x + q + A.T @ y), initial=0.0)
    yp, ym = np.maximum(y, 0), np.minimum(y, 0)
    with np.errstate(invalid="ignore"):
        gap_u = np.where(np.isfinite(u), yp * (u - Ax), np.where(yp > 0, np.inf, 0.0))
        gap_l = np.where(np.isfinite(l), ym * (l - Ax), np.where(ym < 0, np.inf, 0.0))
    comp = max(np.max(np.abs(gap_u), initial=0.0), np.max(np.abs(gap_l), initial=0.0))
    return pri, dua, comp


# --------------------------------------------------------------------------------------------
# path tables
# --------------------------------------------------------------------------------------------
class PathTables:
    """Plain arrays of a reference ReferencePath (or of the product's) + the cos/sin tables."""

    def __init__(self, x, y, psi, kappa, v_ref, seg_len, border, circular=True):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.x, self.y, self.psi, self.kappa = f(x), f(y), f(psi), f(kappa)
        self.v_ref = f(v_ref)
        self.seg_len = f(seg_len)
        self.border = f(border).reshape(-1, 4)
        self.n_wp = len(self.x)
        self.circular = bool(circular)
        n = self.n_wp
        # ds_next[k] = get_waypoint(k+1) - get_waypoint(k) (Waypoint.__sub__, rp.py:57)
        ds = np.empty(n)
        for k in range(n):
            k1 = (k + 1) % n
            ds[k] = ((self.x[k1] - self.x[k]) ** 2 + (self.y[k1] - self.y[k]) ** 2) ** 0.5
        self.ds_next = ds
        self.cos_psi = f([np.cos(p) for p in self.psi])
        self.sin_psi = f([np.sin(p) for p in self.psi])
        a_ub = [np.mod(math.pi / 2 + p + math.pi, 2 * math.pi) - math.pi for p in self.psi]   # rp.py:622
        a_lb = [np.mod(-math.pi / 2 + p + math.pi, 2 * math.pi) - math.pi for p in self.psi]  # rp.py:624
        self.cos_ub, self.sin_ub = f([np.cos(a) for a in a_ub]), f([np.sin(a) for a in a_ub])
        self.cos_lb, self.sin_lb = f([np.cos(a) for a in a_lb]), f([np.sin(a) for a in a_lb])
        self.length_cum = np.cumsum(self.seg_len)  # sbm.py:262

    @classmethod
    def from_reference(cls, rp):
        w = rp.waypoints
        border = [[wp.static_border_cells[0][0], wp.static_border_cells[0][1],
                   wp.static_border_cells[1][0], wp.static_border_cells[1][1]] for wp in w]
        v = [wp.v_ref if wp.v_ref is not None else np.nan for wp in w]
        return cls([wp.x for wp in w], [wp.y for wp in w], [wp.psi for wp in w], [wp.kappa for wp in w],
                   v, rp.segment_lengths, border, rp.circular)

    def c_struct(self):
        s = OrcPath()
        s.n_wp, s.circular = self.n_wp, int(self.circular)
        for k in ("x", "y", "psi", "kappa", "v_ref", "seg_len", "ds_next", "cos_psi", "sin_psi", "cos_ub",
                  "sin_ub", "cos_lb", "sin_lb", "border"):
            setattr(s, k, _dp(getattr(self, k)))
        return s


def mpc_cfg(N, Q, R, QN, xmin, xmax, umin, umax, ay_max, L, safety_margin):
    c = OrcMpcCfg()
    c.N = N
    c.Q[:] = list(Q)
    c.R[:] = list(R)
    c.QN[:] = list(QN)
    c.xmin[:] = list(xmin)
    c.xmax[:] = list(xmax)
    c.umin[:] = list(umin)
    c.umax[:] = list(umax)
    c.ay_max, c.L, c.safety_margin = ay_max, L, safety_margin
    return c


def update_path_constraints(grid, origin, res, pt: PathTables, wp_id, N, min_width, safety_margin,
                            want_stats=False):
    grid = np.ascontiguousarray(grid, dtype=np.int8)
    H, W = grid.shape
    ub, lb, cells = np.empty(N), np.empty(N), np.empty((N, 4))
    ps = pt.c_struct()
    nc, ns = C.c_long(0), C.c_long(0)
    st = lib().orc_update_path_constraints(
        grid.ctypes.data_as(c_i8_p), H, W, float(origin[0]), float(origin[1]), float(res), C.byref(ps),
        int(wp_id), int(N), float(min_width), float(safety_margin), _dp(ub), _dp(lb), _dp(cells),
        C.byref(nc) if want_stats else None, C.byref(ns) if want_stats else None)
    if want_stats:
        return st, ub, lb, cells, nc.value, ns.value
    return st, ub, lb, cells


def compute_width(grid, origin, res, pt: PathTables, max_width):
    grid = np.ascontiguousarray(grid, dtype=np.int8)
    H, W = grid.shape
    ub, lb, border = np.empty(pt.n_wp), np.empty(pt.n_wp), np.empty((pt.n_wp, 4))
    ps = pt.c_struct()
    st = lib().orc_compute_width(grid.ctypes.data_as(c_i8_p), H, W, float(origin[0]), float(origin[1]),
                                 float(res), C.byref(ps), float(max_width), _dp(ub), _dp(lb), _dp(border))
    return st, ub, lb, border


def add_obstacle(grid, origin, res, cx, cy, radius):
    assert grid.dtype == np.int8 and grid.flags.c_contiguous
    H, W = grid.shape
    return lib().orc_add_obstacle(grid.ctypes.data_as(c_i8_p), H, W, float(origin[0]), float(origin[1]),
                                  float(res), float(cx), float(cy), float(radius))


def mpc_assemble(pt: PathTables, cfg: OrcMpcCfg, wp_id, x0, current_control, ub, lb):
    """Returns (Pd, q, A (scipy csc, fixed pattern), l, u) in the reference's layout (MPC.py:128-155)."""
    from scipy import sparse
    N = cfg.N
    n, m, nnz = 5 * N + 3, 8 * N + 6, 16 * N + 6
    Pd, q, l, u = np.empty(n), np.empty(n), np.empty(m), np.empty(m)
    Ap, Ai, Ax = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    cc = np.ascontiguousarray(current_control, dtype=np.float64)
    ub = np.ascontiguousarray(ub, dtype=np.float64)
    lb = np.ascontiguousarray(lb, dtype=np.float64)
    ps = pt.c_struct()
    st = lib().orc_mpc_assemble(C.byref(ps), C.byref(cfg), int(wp_id), _dp(x0), _dp(cc), _dp(ub), _dp(lb),
                                _dp(Pd), _dp(q), _ip(Ap), _ip(Ai), _dp(Ax), _dp(l), _dp(u))
    assert st == 0
    A = sparse.csc_matrix((Ax, Ai, Ap), shape=(m, n))
    return Pd, q, A, l, u


class World:
    """Everything orc_mpc_step needs; keeps the numpy arrays alive."""

    def __init__(self, pt: PathTables, cfg: OrcMpcCfg, grid_shape, origin, res, Ts, **osqp_settings):
        self.pt, self.cfg = pt, cfg
        self.w = OrcWorld()
        self.w.path = pt.c_struct()
        self.w.length_cum = _dp(pt.length_cum)
        self.w.cfg = cfg
        self.w.osqp = default_settings(**osqp_settings)
        self.w.H, self.w.W = int(grid_shape[0]), int(grid_shape[1])
        self.w.ox, self.w.oy, self.w.res, self.w.Ts = float(origin[0]), float(origin[1]), float(res), float(Ts)

    def step(self, grid, state4, current_control, infeas, drive=True):
        N = self.cfg.N
        grid = np.ascontiguousarray(grid, dtype=np.int8)
        st = np.ascontiguousarray(state4, dtype=np.float64).copy()
        cc = np.ascontiguousarray(current_control, dtype=np.float64).copy()
        inf = C.c_int(int(infeas))
        u2, xs, ublb, sp = np.zeros(2), np.zeros(5 * N + 3), np.zeros(2 * N), np.zeros(3)
        wp, it, qs = C.c_int(0), C.c_int(0), C.c_int(0)
        r = lib().orc_mpc_step(C.byref(self.w), grid.ctypes.data_as(c_i8_p), _dp(st), _dp(cc), C.byref(inf),
                               1 if drive else 0, _dp(u2), _dp(xs), _dp(ublb), _dp(sp), C.byref(wp),
                               C.byref(it), C.byref(qs))
        return dict(ret=r, state=st, current_control=cc, infeas=inf.value, u=u2, x=xs, ub=ublb[:N],
                    lb=ublb[N:], spatial=sp, wp_id=wp.value, iters=it.value, qp_status=qs.value)

    def batch_closed_loop(self, grids, states, controls, infeas, alive, steps, n_threads=0):
        grids = np.ascontiguousarray(grids, dtype=np.int8)
        B = states.shape[0]
        stride = 0 if grids.ndim == 2 else grids.shape[1] * grids.shape[2]
        stats = np.zeros(4)
        lib().orc_batch_closed_loop(C.byref(self.w), grids.ctypes.data_as(c_i8_p), stride, B, int(steps),
                                    _dp(states), _dp(controls), _ip(infeas), _ip(alive), _dp(stats),
                                    int(n_threads))
        return stats


def batch_qp_solve(N, Pd, q, Ap, Ai, Ax, l, u, n_threads=0, **settings):
    B = Pd.shape[0]
    n = 5 * N + 3
    s = default_settings(**settings)
    x = np.empty((B, n))
    it, st = np.zeros(B, np.int32), np.zeros(B, np.int32)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    Pd, q, Ax, l, u = f(Pd), f(q), f(Ax), f(l), f(u)
    Ap, Ai = np.ascontiguousarray(Ap, np.int32), np.ascontiguousarray(Ai, np.int32)
    lib().orc_batch_qp_solve(N, B, _dp(Pd), _dp(q), _ip(Ap), _ip(Ai), _dp(Ax), _dp(l), _dp(u), C.byref(s),
                             _dp(x), _ip(it), _ip(st), int(n_threads))
    return x, it, st


def num_threads():
    return lib().orc_num_threads()
