"""ctypes binding of libmpc_b200.so (include/mpc_b200.h) and a thin Engine wrapper.

The library is the product: if it is missing or no CUDA device is usable, everything here raises.
There is no CPU fallback and nothing in this package imports oracle/.
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPC_B200_LIB: developer switch for A/B runs of differently-built libraries (tools/build_variants.sh); product = in-tree .so
LIB_PATH = os.environ.get("MPC_B200_LIB") or os.path.join(_HERE, "libmpc_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)
c_i8_p = C.POINTER(C.c_int8)
c_i64_p = C.POINTER(C.c_int64)

# status bits / codes (mirrors of include/mpc_b200.h)
ST_QP_FALLBACK, ST_DEAD, ST_NO_SEGMENT, ST_END_OF_PATH, ST_INDEX_ERROR, ST_FINISHED = 1, 2, 4, 8, 16, 32
QP_SOLVED, QP_MAX_ITER, QP_PRIMAL_INFEASIBLE, QP_DUAL_INFEASIBLE, QP_NON_CVX = 1, -2, -3, -4, -7


class MpcConfig(C.Structure):
    _fields_ = [("N", C.c_int32), ("Q", C.c_double * 3), ("R", C.c_double * 2), ("QN", C.c_double * 3),
                ("xmin", C.c_double * 3), ("xmax", C.c_double * 3), ("umin", C.c_double * 2),
                ("umax", C.c_double * 2), ("ay_max", C.c_double), ("car_length", C.c_double),
                ("car_width", C.c_double), ("Ts", C.c_double), ("rho", C.c_double), ("sigma", C.c_double),
                ("alpha", C.c_double), ("eps_abs", C.c_double), ("eps_rel", C.c_double),
                ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double), ("max_iter", C.c_int32),
                ("scaling", C.c_int32), ("check_termination", C.c_int32), ("adaptive_rho_interval", C.c_int32),
                ("adaptive_rho_tolerance", C.c_double), ("precision", C.c_int32)]


EXPORTS = [
    "mpc_config_default", "mpc_last_error", "mpc_abi_version", "mpc_engine_create", "mpc_engine_destroy",
    "mpc_engine_set_stream", "mpc_engine_update_config", "mpc_engine_sync", "mpc_set_path", "mpc_set_vref",
    "mpc_set_base_grid", "mpc_set_obstacles", "mpc_get_grid", "mpc_compute_width", "mpc_compute_width_batch", "mpc_localize_t2s",
    "mpc_raycast", "mpc_update_path_constraints", "mpc_assemble_solve", "mpc_solve_qp", "mpc_rollout",
    "mpc_scenarios_init", "mpc_scenarios_set_state", "mpc_scenarios_set_flags", "mpc_step", "mpc_run_closed_loop", "mpc_step_host",
    "mpc_scenarios_ptrs", "mpc_scenarios_read", "mpc_launch_count", "mpc_set_profiling", "mpc_get_profile",
    "mpc_speed_profile", "mpc_speed_profile_batch", "mpc_predict_xy", "mpc_host_io",
]

_lib = None


class MpcError(RuntimeError):
    pass


def load():
    """Loads the shared library (building is __graft_entry__.build()'s job). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpcError("libmpc_b200.so is not built: run `python multi-purpose-mpc_b200/build_ext.py` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.mpc_config_default.argtypes = [C.POINTER(MpcConfig)]
    L.mpc_config_default.restype = None
    L.mpc_last_error.restype = C.c_char_p
    L.mpc_abi_version.restype = C.c_int
    L.mpc_engine_create.argtypes = [C.POINTER(MpcConfig), C.POINTER(vp)]
    L.mpc_engine_destroy.argtypes = [vp]
    L.mpc_engine_set_stream.argtypes = [vp, vp]
    L.mpc_engine_update_config.argtypes = [vp, C.POINTER(MpcConfig)]
    L.mpc_engine_sync.argtypes = [vp]
    L.mpc_set_path.argtypes = [vp, c_double_p, c_double_p, c_double_p, C.c_int32, C.c_int32]
    L.mpc_set_vref.argtypes = [vp, c_double_p, C.c_int32]
    L.mpc_set_base_grid.argtypes = [vp, c_i8_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double]
    L.mpc_set_obstacles.argtypes = [vp, c_double_p, c_int_p, C.c_int32]
    L.mpc_get_grid.argtypes = [vp, C.c_int32, c_i8_p]
    L.mpc_compute_width.argtypes = [vp, C.c_double, c_double_p, c_double_p, c_double_p]
    L.mpc_compute_width_batch.argtypes = [vp, C.c_int32, C.POINTER(C.c_int8), C.c_int32, C.c_int32, C.c_double, C.c_double,
                                          C.c_double, c_double_p, c_int_p, C.c_int32, C.c_double, c_double_p, c_double_p,
                                          c_double_p, c_int_p]
    L.mpc_localize_t2s.argtypes = [vp, vp, vp, vp, vp, C.c_int32]
    L.mpc_raycast.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int32]
    L.mpc_update_path_constraints.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_double, C.c_double, vp, vp, vp,
                                              vp, C.c_int32]
    L.mpc_assemble_solve.argtypes = [vp] + [vp] * 11 + [C.c_int32]
    L.mpc_solve_qp.argtypes = [vp] + [vp] * 8 + [C.c_int32]
    L.mpc_rollout.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int32]
    L.mpc_scenarios_init.argtypes = [vp, c_double_p, C.c_int32]
    L.mpc_scenarios_set_state.argtypes = [vp, c_double_p, c_double_p, c_int_p]
    L.mpc_scenarios_set_flags.argtypes = [vp, c_int_p]
    L.mpc_step.argtypes = [vp]
    L.mpc_run_closed_loop.argtypes = [vp, C.c_int32, c_double_p]
    L.mpc_step_host.argtypes = [vp, c_double_p, c_double_p, c_int_p]
    L.mpc_host_io.argtypes = [vp, C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_int_p)]
    L.mpc_scenarios_ptrs.argtypes = [vp] + [C.POINTER(vp)] * 11
    L.mpc_scenarios_read.argtypes = [vp, c_double_p, c_double_p, c_double_p, c_int_p, c_int_p, c_int_p, c_int_p,
                                     c_int_p, c_double_p, c_double_p]
    L.mpc_launch_count.argtypes = [vp]
    L.mpc_launch_count.restype = C.c_int64
    L.mpc_set_profiling.argtypes = [vp, C.c_int32]
    L.mpc_get_profile.argtypes = [vp, c_double_p, c_i64_p]
    L.mpc_speed_profile.argtypes = [c_double_p, c_double_p, C.c_int32, C.c_double, C.c_double, C.c_double,
                                    C.POINTER(MpcConfig), c_double_p, c_int_p, c_int_p]
    L.mpc_speed_profile_batch.argtypes = [c_double_p, c_double_p, c_int_p, C.c_int32, C.c_double, C.c_double, C.c_double,
                                          C.POINTER(MpcConfig), c_double_p, c_int_p, c_int_p]
    for name in EXPORTS:
        f = getattr(L, name)
        if name not in ("mpc_config_default", "mpc_last_error", "mpc_launch_count"):
            f.restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise MpcError("libmpc_b200 error %d: %s" % (rc, load().mpc_last_error().decode()))


def default_config(**kw):
    cfg = MpcConfig()
    load().mpc_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        if isinstance(getattr(cfg, k), C.Array):
            getattr(cfg, k)[:] = [float(x) for x in v]
        else:
            setattr(cfg, k, v)
    return cfg


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _ptr(t):
    """device pointer of a torch CUDA tensor (or None)"""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def path_table(x, y, psi, kappa, v_ref):
    """The double[12][n_wp] table mpc_set_path expects, computed with the reference's own numpy
    expressions (Waypoint.__sub__ rp.py:57; angles rp.py:221-225 / 622-625)."""
    x, y, psi, kappa = (np.asarray(a, dtype=np.float64) for a in (x, y, psi, kappa))
    n = len(x)
    v = np.full(n, np.nan) if v_ref is None else np.asarray(
        [np.nan if q is None else q for q in v_ref], dtype=np.float64)
    t = np.empty((12, n))
    t[0], t[1], t[2], t[3], t[4] = x, y, psi, kappa, v
    for k in range(n):
        k1 = (k + 1) % n
        t[5, k] = ((x[k1] - x[k]) ** 2 + (y[k1] - y[k]) ** 2) ** 0.5
        p = psi[k]
        t[6, k], t[7, k] = np.cos(p), np.sin(p)
        a_ub = np.mod(math.pi / 2 + p + math.pi, 2 * math.pi) - math.pi
        a_lb = np.mod(-math.pi / 2 + p + math.pi, 2 * math.pi) - math.pi
        t[8, k], t[9, k] = np.cos(a_ub), np.sin(a_ub)
        t[10, k], t[11, k] = np.cos(a_lb), np.sin(a_lb)
    return np.ascontiguousarray(t)


class Engine:
    """Owns one mpc_engine handle on the current CUDA device."""

    def __init__(self, cfg=None, **kw):
        import torch
        if not torch.cuda.is_available():
            raise MpcError("no CUDA device: the B200 engine has no CPU fallback")
        self.L = load()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self.h = C.c_void_p()
        _check(self.L.mpc_engine_create(C.byref(self.cfg), C.byref(self.h)))
        self.N = self.cfg.N
        self.n_wp = 0
        self.B = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.mpc_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration --------------------------------------------------------------------
    def update_config(self, **kw):
        for k, v in kw.items():
            if isinstance(getattr(self.cfg, k), C.Array):
                getattr(self.cfg, k)[:] = [float(x) for x in v]
            else:
                setattr(self.cfg, k, v)
        _check(self.L.mpc_engine_update_config(self.h, C.byref(self.cfg)))

    def set_stream(self, stream_ptr):
        _check(self.L.mpc_engine_set_stream(self.h, C.c_void_p(stream_ptr)))

    def sync(self):
        _check(self.L.mpc_engine_sync(self.h))

    def set_path(self, table12, length_cum, border=None, circular=True):
        table12 = np.ascontiguousarray(table12, dtype=np.float64)
        length_cum = np.ascontiguousarray(length_cum, dtype=np.float64)
        n = table12.shape[1]
        b = None if border is None else np.ascontiguousarray(border, dtype=np.float64)
        _check(self.L.mpc_set_path(self.h, _dp(table12), _dp(length_cum), None if b is None else _dp(b), n,
                                   1 if circular else 0))
        self.n_wp = n

    def set_vref(self, v_ref):
        v = np.ascontiguousarray(v_ref, dtype=np.float64)
        _check(self.L.mpc_set_vref(self.h, _dp(v), len(v)))

    def set_base_grid(self, data, origin, resolution):
        d = np.ascontiguousarray(data, dtype=np.int8)
        _check(self.L.mpc_set_base_grid(self.h, d.ctypes.data_as(c_i8_p), d.shape[0], d.shape[1], float(origin[0]),
                                        float(origin[1]), float(resolution)))
        self.grid_shape = d.shape

    def set_obstacles(self, obs, offsets):
        if obs is None:
            _check(self.L.mpc_set_obstacles(self.h, None, None, 0))
            return
        o = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        _check(self.L.mpc_set_obstacles(self.h, _dp(o), off.ctypes.data_as(c_int_p), len(off) - 1))

    def get_grid(self, b=0):
        out = np.empty(self.grid_shape, dtype=np.int8)
        _check(self.L.mpc_get_grid(self.h, int(b), out.ctypes.data_as(c_i8_p)))
        return out

    def compute_width(self, max_width):
        ub, lb, border = np.empty(self.n_wp), np.empty(self.n_wp), np.empty((self.n_wp, 4))
        _check(self.L.mpc_compute_width(self.h, float(max_width), _dp(ub), _dp(lb), _dp(border)))
        return ub, lb, border

    def compute_width_batch(self, maps, origin, resolution, tables, max_width):
        """ReferencePath._compute_width for T tracks at once.  maps: [T, H, W] int8 (Map.data of every track), tables: list
        of T double[12][n_wp_t] path tables (path_table()).  Returns (ub, lb, border, err): [T, n_max], [T, n_max],
        [T, n_max, 4], [T]; columns beyond a track's own n_wp are zero."""
        maps = np.ascontiguousarray(maps, dtype=np.int8)
        T, H, W = maps.shape
        assert len(tables) == T
        n_wp = np.array([t.shape[1] for t in tables], dtype=np.int32)
        n_max = int(n_wp.max())
        tab = np.zeros((T, 12, n_max))
        for t, a in enumerate(tables):
            tab[t, :, :a.shape[1]] = a
        ub, lb, border = np.empty((T, n_max)), np.empty((T, n_max)), np.empty((T, n_max, 4))
        err = np.zeros(T, np.int32)
        _check(self.L.mpc_compute_width_batch(self.h, T, maps.ctypes.data_as(c_i8_p), H, W, float(origin[0]), float(origin[1]),
                                              float(resolution), _dp(tab), _ip(n_wp), n_max, float(max_width), _dp(ub), _dp(lb),
                                              _dp(border), _ip(err)))
        return ub, lb, border, err

    # ---- per-step kernels on caller-owned torch CUDA tensors ------------------------------------
    def localize_t2s(self, state, wp_id, spatial, flags=None):
        _check(self.L.mpc_localize_t2s(self.h, _ptr(state), _ptr(wp_id), _ptr(spatial), _ptr(flags), state.shape[1]))

    def raycast(self, wp_id, ub, lb, cells_sm=None, flags=None):
        _check(self.L.mpc_raycast(self.h, _ptr(wp_id), _ptr(ub), _ptr(lb), _ptr(cells_sm), _ptr(flags),
                                  wp_id.shape[0]))

    def update_path_constraints(self, wp_id, first_offset, N, min_width, safety_margin, ub, lb, cells_sm=None,
                                flags=None):
        _check(self.L.mpc_update_path_constraints(self.h, _ptr(wp_id), int(first_offset), int(N), float(min_width),
                                                  float(safety_margin), _ptr(ub), _ptr(lb), _ptr(cells_sm),
                                                  _ptr(flags), wp_id.shape[0]))

    def assemble_solve(self, spatial, wp_id, control, ub, lb, infeas, u_out, x_out=None, iters=None, qp_status=None,
                       flags=None):
        _check(self.L.mpc_assemble_solve(self.h, _ptr(spatial), _ptr(wp_id), _ptr(control), _ptr(ub), _ptr(lb),
                                         _ptr(infeas), _ptr(u_out), _ptr(x_out), _ptr(iters), _ptr(qp_status),
                                         _ptr(flags), wp_id.shape[0]))

    def solve_qp(self, Pd, q, Ax, l, u, x_out=None, iters=None, qp_status=None):
        _check(self.L.mpc_solve_qp(self.h, _ptr(Pd), _ptr(q), _ptr(Ax), _ptr(l), _ptr(u), _ptr(x_out), _ptr(iters),
                                   _ptr(qp_status), Pd.shape[0]))

    def predict_xy(self, x_sol, wp_id, xy_out):
        """MPC.update_prediction for every scenario: xy_out[B, N-2, 2] from the solver output x_sol[B, 5N+3]."""
        _check(self.L.mpc_predict_xy(self.h, _ptr(x_sol), _ptr(wp_id), _ptr(xy_out), x_sol.shape[0]))

    def rollout(self, state, spatial, wp_id, u, flags=None):
        _check(self.L.mpc_rollout(self.h, _ptr(state), _ptr(spatial), _ptr(wp_id), _ptr(u), _ptr(flags),
                                  state.shape[1]))

    # ---- engine-owned scenarios ---------------------------------------------------------------
    def scenarios_init(self, state4xB):
        s = np.ascontiguousarray(state4xB, dtype=np.float64)
        assert s.ndim == 2 and s.shape[0] == 4
        _check(self.L.mpc_scenarios_init(self.h, _dp(s), s.shape[1]))
        self.B = s.shape[1]

    def scenarios_set_state(self, state4xB=None, control=None, infeas=None):
        s = None if state4xB is None else np.ascontiguousarray(state4xB, dtype=np.float64)
        c = None if control is None else np.ascontiguousarray(control, dtype=np.float64)
        i = None if infeas is None else np.ascontiguousarray(infeas, dtype=np.int32)
        _check(self.L.mpc_scenarios_set_state(self.h, None if s is None else _dp(s), None if c is None else _dp(c),
                                              None if i is None else i.ctypes.data_as(c_int_p)))

    def scenarios_set_flags(self, flags):
        f = np.ascontiguousarray(flags, dtype=np.int32)
        assert f.shape == (self.B,)
        _check(self.L.mpc_scenarios_set_flags(self.h, _ip(f)))

    def step(self):
        _check(self.L.mpc_step(self.h))

    def run_closed_loop(self, max_steps):
        stats = np.zeros(8)
        _check(self.L.mpc_run_closed_loop(self.h, int(max_steps), _dp(stats)))
        return dict(zip(("scenario_steps", "qp_solves", "admm_iters", "qp_fallbacks", "dead", "finished",
                         "sum_abs_ey", "max_abs_ey"), stats.tolist()))

    def host_io(self):
        """The engine's page-locked I/O block as numpy views (state[4, B], u[B, 2], flags[B]); step_host on these
        skips the staging copies and runs as a single graph launch.  Valid until the next scenarios_init."""
        ps, pu, pf = c_double_p(), c_double_p(), c_int_p()
        _check(self.L.mpc_host_io(self.h, C.byref(ps), C.byref(pu), C.byref(pf)))
        B = self.B
        return (np.ctypeslib.as_array(ps, shape=(4, B)), np.ctypeslib.as_array(pu, shape=(B, 2)),
                np.ctypeslib.as_array(pf, shape=(B,)))

    def step_host(self, state4xB, u_out, flags=None):
        _check(self.L.mpc_step_host(self.h, _dp(state4xB), _dp(u_out),
                                    None if flags is None else flags.ctypes.data_as(c_int_p)))

    def scenarios_read(self):
        B, N = self.B, self.N
        out = dict(state=np.empty((4, B)), control=np.empty((B, 2 * N)), u=np.empty((B, 2)),
                   iters=np.empty(B, np.int32), qp_status=np.empty(B, np.int32), flags=np.empty(B, np.int32),
                   infeas=np.empty(B, np.int32), wp_id=np.empty(B, np.int32), ub=np.empty((B, N)),
                   lb=np.empty((B, N)))
        ip = lambda a: a.ctypes.data_as(c_int_p)
        _check(self.L.mpc_scenarios_read(self.h, _dp(out["state"]), _dp(out["control"]), _dp(out["u"]),
                                         ip(out["iters"]), ip(out["qp_status"]), ip(out["flags"]), ip(out["infeas"]),
                                         ip(out["wp_id"]), _dp(out["ub"]), _dp(out["lb"])))
        return out

    def launch_count(self):
        return int(self.L.mpc_launch_count(self.h))

    def set_profiling(self, on):
        _check(self.L.mpc_set_profiling(self.h, 1 if on else 0))

    def get_profile(self):
        ms = np.zeros(4)
        n = np.zeros(4, np.int64)
        _check(self.L.mpc_get_profile(self.h, _dp(ms), n.ctypes.data_as(c_i64_p)))
        return dict(zip(("localize", "raycast", "assemble_solve", "rollout"), ms.tolist())), n.tolist()
