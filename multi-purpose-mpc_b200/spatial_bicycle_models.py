"""TemporalState / SimpleSpatialState / BicycleModel -- host mirror of the reference's
src/spatial_bicycle_models.py, plus BatchedBicycleModel (B independent cars on one path).

The single-car classes keep the reference's attribute bag semantics (temporal_state.x, spatial_state[:],
s, wp_id, current_waypoint) and run localise / t2s / drive through the same CUDA kernels as the batched
model with B = 1 (K4: csrc/geometry.cu::localize_t2s_kernel, rollout_kernel).
"""
import numpy as np

from . import _lib


class TemporalState:
    def __init__(self, x, y, psi):
        self.x = x
        self.y = y
        self.psi = psi
        self.members = ['x', 'y', 'psi']

    def __iadd__(self, other):
        for i, name in enumerate(self.members):
            vars(self)[name] += other[i]
        return self


class SpatialState:
    """Indexable / sliceable attribute bag (sbm.py:53-91)."""

    def __init__(self):
        self.members = None
        self.e_y = None
        self.e_psi = None

    def __getitem__(self, item):
        members = [self.members[item]] if isinstance(item, int) else self.members[item]
        return [vars(self)[key] for key in members]

    def __setitem__(self, key, value):
        vars(self)[self.members[key]] = value

    def __len__(self):
        return len(self.members)

    def __iadd__(self, other):
        for i, name in enumerate(self.members):
            vars(self)[name] += other[i]
        return self

    def list_states(self):
        return self.members


class SimpleSpatialState(SpatialState):
    def __init__(self, e_y=0.0, e_psi=0.0, t=0.0):
        super(SimpleSpatialState, self).__init__()
        self.e_y = e_y
        self.e_psi = e_psi
        self.t = t
        self.members = ['e_y', 'e_psi', 't']


class _Batch:
    """B scenarios' device tensors + the engine calls that move them (shared by the single-car and the
    batched front ends).  Layouts are those of include/mpc_b200.h."""

    def __init__(self, B, N):
        import torch
        self.torch = torch
        self.B, self.N = B, N
        dev = torch.device("cuda", torch.cuda.current_device())
        f64, i32 = torch.float64, torch.int32
        z = lambda *shape, dt=f64: torch.zeros(shape, dtype=dt, device=dev)
        self.state = z(4, B)
        self.spatial = z(2, B)
        self.wp_id = z(B, dt=i32)
        self.control = z(B, 2 * N)
        self.ub = z(B, N)
        self.lb = z(B, N)
        self.infeas = z(B, dt=i32)
        self.u = z(B, 2)
        self.x_sol = z(B, 5 * N + 3)
        self.iters = z(B, dt=i32)
        self.qp_status = z(B, dt=i32)
        self.flags = z(B, dt=i32)


class SpatialBicycleModel:
    def __init__(self, reference_path, length, width, Ts):
        """(sbm.py:116-153)"""
        self.eps = 1e-12
        self.length = length
        self.width = width
        self.safety_margin = self._compute_safety_margin()
        self.reference_path = reference_path
        self.s = 0.0
        self.Ts = Ts
        self.wp_id = 0
        self.current_waypoint = self.reference_path.waypoints[self.wp_id]
        self.spatial_state = None
        self.temporal_state = None
        self._eng = None  # set by MPC (shares its engine) or created lazily
        self._batch = None

    # -- device plumbing ------------------------------------------------------------------------
    def _engine(self):
        if self._eng is None:
            self._eng = _lib.Engine(car_length=float(self.length), car_width=float(self.width), Ts=float(self.Ts))
            self._eng_path_version = None
        rp = self.reference_path
        if getattr(self, "_eng_path_version", None) != rp.version:
            t, lc, border = rp.tables()
            self._eng.set_path(t, lc, border, rp.circular)
            self._eng_path_version = rp.version
        if self._batch is None:
            self._batch = _Batch(1, self._eng.N)
        return self._eng, self._batch

    def _push_state(self, b):
        ts = self.temporal_state
        b.state[:, 0] = b.torch.tensor([ts.x, ts.y, ts.psi, self.s], dtype=b.torch.float64)

    # -- reference API --------------------------------------------------------------------------
    def s2t(self, reference_waypoint, reference_state):
        """Spatial -> temporal state about a waypoint (sbm.py:155-181)."""
        if isinstance(reference_state, np.ndarray):
            e_y, e_psi = reference_state[0], reference_state[1]
        elif isinstance(reference_state, SpatialState):
            e_y, e_psi = reference_state.e_y, reference_state.e_psi
        else:
            print('Reference State type not supported!')
            exit(1)
        x = reference_waypoint.x - e_y * np.sin(reference_waypoint.psi)
        y = reference_waypoint.y + e_y * np.cos(reference_waypoint.psi)
        psi = reference_waypoint.psi + e_psi
        return TemporalState(x, y, psi)

    def t2s(self, reference_waypoint, reference_state):
        """Temporal -> spatial state about a waypoint (sbm.py:183-219); t is reset to 0."""
        import math
        if isinstance(reference_state, np.ndarray):
            x, y, psi = reference_state[0], reference_state[1], reference_state[2]
        elif isinstance(reference_state, TemporalState):
            x, y, psi = reference_state.x, reference_state.y, reference_state.psi
        else:
            print('Reference State type not supported!')
            exit(1)
        e_y = np.cos(reference_waypoint.psi) * (y - reference_waypoint.y) - \
            np.sin(reference_waypoint.psi) * (x - reference_waypoint.x)
        e_psi = psi - reference_waypoint.psi
        e_psi = np.mod(e_psi + math.pi, 2 * math.pi) - math.pi
        return SimpleSpatialState(e_y, e_psi, 0.0)

    def drive(self, u):
        """Explicit-Euler step of the kinematic bicycle + arc-length update (sbm.py:221-244) -- K4."""
        eng, b = self._engine()
        self._push_state(b)
        b.spatial[:, 0] = b.torch.tensor([self.spatial_state.e_y, self.spatial_state.e_psi], dtype=b.torch.float64)
        b.wp_id[0] = int(self.wp_id) % self.reference_path.n_waypoints
        b.u[0] = b.torch.tensor([float(u[0]), float(u[1])], dtype=b.torch.float64)
        eng.rollout(b.state, b.spatial, b.wp_id, b.u, None)
        st = b.state[:, 0].cpu().numpy()
        self.temporal_state.x, self.temporal_state.y, self.temporal_state.psi = st[0], st[1], st[2]
        self.s = st[3]

    def _compute_safety_margin(self):
        return self.width / np.sqrt(2)  # sbm.py:246-254

    def get_current_waypoint(self):
        """Nearest waypoint by travelled arc length (sbm.py:256-279) -- K4."""
        eng, b = self._engine()
        self._push_state(b)
        b.flags.zero_()
        eng.localize_t2s(b.state, b.wp_id, b.spatial, b.flags)
        if int(b.flags[0].item()) & _lib.ST_FINISHED:
            raise IndexError("index %d is out of bounds" % self.reference_path.n_waypoints)  # sbm.py:271
        self.wp_id = int(b.wp_id[0].item())
        self.current_waypoint = self.reference_path.waypoints[self.wp_id]
        sp = b.spatial[:, 0].cpu().numpy()
        self._last_spatial = SimpleSpatialState(sp[0], sp[1], 0.0)

    def show(self):
        raise NotImplementedError("plotting is not part of the B200 engine")


class BicycleModel(SpatialBicycleModel):
    def __init__(self, reference_path, length, width, Ts):
        super(BicycleModel, self).__init__(reference_path, length=length, width=width, Ts=Ts)
        self.spatial_state = SimpleSpatialState()
        self.n_states = len(self.spatial_state)
        self.temporal_state = self.s2t(reference_state=self.spatial_state, reference_waypoint=self.current_waypoint)

    def get_temporal_derivatives(self, state, input, kappa):
        """(sbm.py:347-366)"""
        e_y, e_psi, t = state
        v, delta = input
        s_dot = 1 / (1 - (e_y * kappa)) * v * np.cos(e_psi)
        psi_dot = v / self.length * np.tan(delta)
        return s_dot, psi_dot

    def get_spatial_derivatives(self, state, input, kappa):
        """(sbm.py:368-389)"""
        e_y, e_psi, t = state
        v, delta = input
        s_dot, psi_dot = self.get_temporal_derivatives(state, input, kappa)
        return np.array([v * np.sin(e_psi) / s_dot, psi_dot / s_dot - kappa, 1 / s_dot])

    def linearize(self, v_ref, kappa_ref, delta_s):
        """LTV Jacobians about (v_ref, kappa_ref) over one path segment (sbm.py:391-417).  The engine
        builds these inside the assemble+solve kernel (csrc/admm.cuh::assemble_stage); this host
        version exists for API compatibility."""
        A = np.array([[1, delta_s, 0],
                      [-kappa_ref ** 2 * delta_s, 1, 0],
                      [-kappa_ref / v_ref * delta_s, 0, 1]])
        B = np.array([[0, 0],
                      [0, delta_s],
                      [-1 / (v_ref ** 2) * delta_s, 0]])
        f = np.array([0.0, 0.0, 1 / v_ref * delta_s])
        return f, A, B


class BatchedBicycleModel:
    """B independent cars on one reference path; all state lives on the GPU (fp64, SoA).

    temporal_state: torch view (3, B) rows x, y, psi;  s: (B,);  spatial_state: (2, B) rows e_y, e_psi;
    wp_id: (B,) int32.  drive(u) advances every car in place (sbm.py:221-244 per car)."""

    def __init__(self, reference_path, length, width, Ts, B, start_wp=None, e_y=None, e_psi=None):
        self.reference_path = reference_path
        self.length, self.width, self.Ts, self.B = length, width, Ts, int(B)
        self.safety_margin = width / np.sqrt(2)
        self.n_states = 3
        self._eng = None
        self._batch = None
        rp = reference_path
        w = rp.waypoints
        start_wp = np.zeros(self.B, dtype=np.int64) if start_wp is None else np.asarray(start_wp, dtype=np.int64)
        e_y = np.zeros(self.B) if e_y is None else np.asarray(e_y, dtype=np.float64)
        e_psi = np.zeros(self.B) if e_psi is None else np.asarray(e_psi, dtype=np.float64)
        wx = np.array([p.x for p in w])[start_wp]
        wy = np.array([p.y for p in w])[start_wp]
        wpsi = np.array([p.psi for p in w])[start_wp]
        # s2t about the start waypoint (sbm.py:171-175); s = arc length of that waypoint
        self.initial_state = np.stack([wx - e_y * np.sin(wpsi), wy + e_y * np.cos(wpsi), wpsi + e_psi,
                                       np.cumsum(rp.segment_lengths)[start_wp]])

    def _attach(self, eng, batch):
        self._eng, self._batch = eng, batch
        batch.state.copy_(batch.torch.tensor(self.initial_state, dtype=batch.torch.float64))

    @property
    def temporal_state(self):
        return self._batch.state[:3]

    @property
    def s(self):
        return self._batch.state[3]

    @property
    def spatial_state(self):
        return self._batch.spatial

    @property
    def wp_id(self):
        return self._batch.wp_id

    def get_current_waypoint(self):
        self._eng.localize_t2s(self._batch.state, self._batch.wp_id, self._batch.spatial, self._batch.flags)

    def drive(self, u=None):
        b = self._batch
        if u is not None and u.data_ptr() != b.u.data_ptr():
            b.u.copy_(u)
        self._eng.rollout(b.state, b.spatial, b.wp_id, b.u, b.flags)
