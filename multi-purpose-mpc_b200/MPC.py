"""MPC / BatchedMPC -- host mirror of the reference's src/MPC.py.

MPC(model, N, Q, R, QN, StateConstraints, InputConstraints, ay_max).get_control() keeps the reference's
signature and return value (np.array([v, delta])).  Per call the GPU runs
    K4 localise + t2s  ->  K3 raycast  ->  K1+K2 LTV assembly + OSQP-equivalent ADMM
(MPC.py:161-222, rp.py:522-648, sbm.py:183-279,391-417); BatchedMPC does the same for B scenarios at
once and adds run_closed_loop() (one CUDA graph per step).
"""
import numpy as np

from . import _lib
from .spatial_bicycle_models import SimpleSpatialState, _Batch


def _diag(M):
    """diagonal of a scipy sparse / numpy matrix / 1-D sequence"""
    if hasattr(M, "diagonal"):
        return np.asarray(M.diagonal(), dtype=np.float64).ravel()
    M = np.asarray(M, dtype=np.float64)
    return M if M.ndim == 1 else np.diag(M)


class _MPCBase:
    def _make_engine(self, model, N, Q, R, QN, StateConstraints, InputConstraints, ay_max, **engine_opts):
        self.N = N
        self.Q, self.R, self.QN = Q, R, QN
        self.model = model
        self.nx = model.n_states
        self.nu = 2
        self.state_constraints = StateConstraints
        self.input_constraints = InputConstraints
        self.ay_max = ay_max
        cfg = _lib.default_config(
            N=int(N), Q=_diag(Q), R=_diag(R), QN=_diag(QN), xmin=StateConstraints['xmin'],
            xmax=StateConstraints['xmax'], umin=InputConstraints['umin'], umax=InputConstraints['umax'],
            ay_max=float(ay_max), car_length=float(model.length), car_width=float(model.width), Ts=float(model.Ts),
            **engine_opts)
        self.engine = _lib.Engine(cfg)
        self._path_version = None
        self._map_version = None
        self._sync_tables()

    def _sync_tables(self):
        rp = self.model.reference_path
        if self._path_version != rp.version:
            t, lc, border = rp.tables()
            self.engine.set_path(t, lc, border, rp.circular)
            self._path_version = rp.version
        if self._map_version != rp.map.version:
            self.engine.set_base_grid(rp.map.data, rp.map.origin, rp.map.resolution)
            self._map_version = rp.map.version

    def set_solver_options(self, **kw):
        """OSQP settings / engine precision (eps_abs, eps_rel, max_iter, precision, refine, ...)."""
        self.engine.update_config(**kw)


class MPC(_MPCBase):
    def __init__(self, model, N, Q, R, QN, StateConstraints, InputConstraints, ay_max, **engine_opts):
        self._make_engine(model, N, Q, R, QN, StateConstraints, InputConstraints, ay_max, **engine_opts)
        self.current_prediction = None
        self.infeasibility_counter = 0
        self.current_control = np.zeros((self.nu * self.N))
        self._b = _Batch(1, N)
        model._eng, model._batch = self.engine, self._b
        model._eng_path_version = self._path_version
        self.last_status = None
        self.last_iters = None
        self.last_x = None

    def get_control(self):
        """One MPC step for the single car (MPC.py:161-222)."""
        m, b, eng, t = self.model, self._b, self.engine, self._b.torch
        self._sync_tables()
        m._eng_path_version = self._path_version
        ts = m.temporal_state
        b.state[:, 0] = t.tensor([ts.x, ts.y, ts.psi, m.s], dtype=t.float64)
        b.control[0] = t.tensor(np.asarray(self.current_control, dtype=np.float64))
        b.infeas[0] = int(self.infeasibility_counter)
        b.flags.zero_()
        eng.localize_t2s(b.state, b.wp_id, b.spatial, b.flags)
        eng.raycast(b.wp_id, b.ub, b.lb, None, b.flags)
        eng.assemble_solve(b.spatial, b.wp_id, b.control, b.ub, b.lb, b.infeas, b.u, b.x_sol, b.iters, b.qp_status,
                           b.flags)
        flags = int(b.flags[0].item())
        if flags & _lib.ST_FINISHED:
            raise IndexError("index out of bounds: the car is past the end of the path (sbm.py:271)")
        if flags & _lib.ST_NO_SEGMENT:
            raise ValueError("max() arg is an empty sequence")  # rp.py:547
        if flags & _lib.ST_END_OF_PATH:
            print('Reached end of path!')
            exit(1)
        if flags & _lib.ST_INDEX_ERROR:
            raise IndexError("ray left the occupancy grid")
        m.wp_id = int(b.wp_id[0].item())
        m.current_waypoint = m.reference_path.waypoints[m.wp_id]
        sp = b.spatial[:, 0].cpu().numpy()
        m.spatial_state = SimpleSpatialState(sp[0], sp[1], 0.0)
        self.last_status = int(b.qp_status[0].item())
        self.last_iters = int(b.iters[0].item())
        u = b.u[0].cpu().numpy().copy()
        if flags & _lib.ST_QP_FALLBACK:
            print('Infeasible problem. Previously predicted control signal used!')
        else:
            self.current_control = b.control[0].cpu().numpy().copy()
            self.last_x = b.x_sol[0].cpu().numpy().copy()
            x = np.reshape(self.last_x[:(self.N + 1) * self.nx], (self.N + 1, self.nx))
            self.current_prediction = self.update_prediction(x)
        self.infeasibility_counter = int(b.infeas[0].item())
        if self.infeasibility_counter == (self.N - 1):
            print('No control signal computed!')
            exit(1)
        return u

    def update_prediction(self, spatial_state_prediction):
        """Predicted x / y coordinates for stages 2..N-1 (MPC.py:224-248); visualisation only."""
        x_pred, y_pred = [], []
        for n in range(2, self.N):
            wp = self.model.reference_path.get_waypoint(self.model.wp_id + n)
            st = self.model.s2t(wp, spatial_state_prediction[n, :])
            x_pred.append(st.x)
            y_pred.append(st.y)
        return x_pred, y_pred

    def show_prediction(self):
        raise NotImplementedError("plotting is not part of the B200 engine")


class BatchedMPC(_MPCBase):
    """B scenarios stepped at once.  `model` is a BatchedBicycleModel.  Per-scenario obstacle sets are
    optional: set_scenario_obstacles(list of lists of Obstacle / (cx, cy, r))."""

    def __init__(self, model, N, Q, R, QN, StateConstraints, InputConstraints, ay_max, **engine_opts):
        self._make_engine(model, N, Q, R, QN, StateConstraints, InputConstraints, ay_max, **engine_opts)
        self.B = model.B
        self._b = _Batch(self.B, N)
        model._attach(self.engine, self._b)

    # -- views ------------------------------------------------------------------------------------
    @property
    def current_control(self):
        return self._b.control

    @property
    def infeasibility_counter(self):
        return self._b.infeas

    @property
    def status(self):
        return self._b.qp_status

    @property
    def iters(self):
        return self._b.iters

    @property
    def flags(self):
        return self._b.flags

    @property
    def solution(self):
        return self._b.x_sol

    @property
    def bounds(self):
        return self._b.ub, self._b.lb

    def set_scenario_obstacles(self, obstacle_lists):
        obs, off = [], [0]
        for lst in obstacle_lists:
            for o in lst:
                obs.append((o.cx, o.cy, o.radius) if hasattr(o, "cx") else tuple(o))
            off.append(len(obs))
        assert len(off) - 1 == self.B
        self._sync_tables()
        self.engine.set_obstacles(np.array(obs, dtype=np.float64).reshape(-1, 3), np.array(off, dtype=np.int32))

    def get_control(self, want_solution=True):
        """get_control() for every scenario; returns the (B, 2) torch tensor of (v, delta)."""
        b, eng = self._b, self.engine
        self._sync_tables()
        eng.localize_t2s(b.state, b.wp_id, b.spatial, b.flags)
        eng.raycast(b.wp_id, b.ub, b.lb, None, b.flags)
        eng.assemble_solve(b.spatial, b.wp_id, b.control, b.ub, b.lb, b.infeas, b.u, b.x_sol if want_solution else None,
                           b.iters, b.qp_status, b.flags)
        return b.u

    def update_prediction(self):
        """MPC.update_prediction (MPC.py:224-248) for every scenario, on the device: (B, N-2, 2) tensor of the world
        x / y of the predicted stages 2 .. N-1 of the last get_control(want_solution=True)."""
        b, t = self._b, self._b.torch
        if getattr(b, "xy_pred", None) is None:
            b.xy_pred = t.zeros((self.B, max(self.N - 2, 0), 2), dtype=t.float64, device=b.x_sol.device)
        self.engine.predict_xy(b.x_sol, b.wp_id, b.xy_pred)
        return b.xy_pred

    @property
    def current_prediction(self):
        return self.update_prediction()

    def run_closed_loop(self, max_steps):
        """max_steps x (get_control + drive) on engine-owned state (CUDA graph); returns the statistics
        dict and leaves the final state in the model's tensors."""
        b, eng = self._b, self.engine
        self._sync_tables()
        eng.sync()
        st = b.state.cpu().numpy()
        if getattr(self, "_loop_B", None) != self.B:   # (re)allocate the engine-owned fleet only when its size changes:
            eng.scenarios_init(st)                     # scenarios_init drops the captured graphs
            self._loop_B = self.B
        eng.scenarios_set_state(st, b.control.cpu().numpy(), b.infeas.cpu().numpy())
        # scenarios that died (N-1 infeasible QPs in a row: the reference exits, MPC.py:218-220) or finished their lap in an
        # earlier call stay out of the loop; the per-step bits (fallback, ...) are recomputed by the next step
        eng.scenarios_set_flags(b.flags.cpu().numpy() & (2 | 32))
        stats = eng.run_closed_loop(max_steps)
        out = eng.scenarios_read()
        t = b.torch
        b.state.copy_(t.tensor(out["state"]))
        b.control.copy_(t.tensor(out["control"]))
        b.u.copy_(t.tensor(out["u"]))
        b.iters.copy_(t.tensor(out["iters"]))
        b.qp_status.copy_(t.tensor(out["qp_status"]))
        b.flags.copy_(t.tensor(out["flags"]))
        b.infeas.copy_(t.tensor(out["infeas"]))
        b.wp_id.copy_(t.tensor(out["wp_id"]))
        return stats
