"""Waypoint / ReferencePath -- host mirror of the reference's src/reference_path.py.

Same constructor, attributes and method names.  One-off geometry (waypoint construction, rp.py:110-204)
is numpy on the host, written so that every value is bit-identical to the reference's; everything that
touches the occupancy grid runs on the GPU:
  _compute_width           -> K3b  (mpc_compute_width,           rp.py:206-287)
  update_path_constraints  -> K3   (mpc_update_path_constraints, rp.py:466-648)
  compute_speed_profile    -> the speed-profile QP on the device (rp.py:289-354)
"""
import math

import numpy as np

from . import _lib


class Waypoint:
    def __init__(self, x, y, psi, kappa):
        """(rp.py:20-48) position, heading, curvature + v_ref / width attributes filled in later."""
        self.x = x
        self.y = y
        self.psi = psi
        self.kappa = kappa
        self.v_ref = None
        self.lb = None
        self.ub = None
        self.static_border_cells = None
        self.dynamic_border_cells = None

    def __sub__(self, other):
        """Euclidean distance between two waypoints (rp.py:50-57)."""
        return ((self.x - other.x) ** 2 + (self.y - other.y) ** 2) ** 0.5


class ReferencePath:
    def __init__(self, map, wp_x, wp_y, resolution, smoothing_distance, max_width, circular):
        self.eps = 1e-12
        self.map = map
        self.resolution = resolution
        self.smoothing_distance = smoothing_distance
        self.circular = circular
        self.waypoints = self._construct_path(wp_x, wp_y)
        self.n_waypoints = len(self.waypoints)
        self.length, self.segment_lengths = self._compute_length()
        self._engine = None
        self._engine_map_version = None
        self.version = 0  # bumped when v_ref / widths change so that MPC engines re-upload
        self._compute_width(max_width=max_width)

    # ------------------------------------------------------------------ construction (host) --
    def _construct_path(self, wp_x, wp_y):
        """Densify the corner polyline, smooth with a centred moving average (rp.py:110-146)."""
        n_seg = len(wp_x) - 1
        counts = [int(np.sqrt((wp_x[i + 1] - wp_x[i]) ** 2 + (wp_y[i + 1] - wp_y[i]) ** 2) / self.resolution)
                  for i in range(n_seg)]
        xs = np.concatenate([np.linspace(wp_x[i], wp_x[i + 1], counts[i], endpoint=False) for i in range(n_seg)]
                            + [np.array([wp_x[-1]], dtype=float)])
        ys = np.concatenate([np.linspace(wp_y[i], wp_y[i + 1], counts[i], endpoint=False) for i in range(n_seg)]
                            + [np.array([wp_y[-1]], dtype=float)])
        sd = self.smoothing_distance
        # np.mean over the same 2*sd+1 window the reference averages (same summation order)
        sx = [np.mean(xs[i - sd:i + sd + 1]) for i in range(sd, len(xs) - sd)]
        sy = [np.mean(ys[i - sd:i + sd + 1]) for i in range(sd, len(ys) - sd)]
        return self._construct_waypoints(list(zip(sx, sy)))

    def _construct_waypoints(self, waypoint_coordinates):
        """(x, y) list -> Waypoint objects with heading and curvature (rp.py:148-193)."""
        pts = [np.array(c) for c in waypoint_coordinates]
        out = []
        heading_prev = None
        for i in range(len(pts) - 1):
            ahead = pts[i + 1] - pts[i]
            psi = np.arctan2(ahead[1], ahead[0])
            dist_ahead = np.linalg.norm(ahead, 2)
            if i == 0:
                kappa = 0
            else:
                turn = np.mod(psi - heading_prev + math.pi, 2 * math.pi) - math.pi
                kappa = turn / (dist_ahead + self.eps)
            heading_prev = psi
            out.append(Waypoint(pts[i][0], pts[i][1], psi, kappa))
        return out

    def _compute_length(self):
        """(rp.py:195-204)"""
        seg = [0.0] + [self.waypoints[i + 1] - self.waypoints[i] for i in range(len(self.waypoints) - 1)]
        return sum(seg), seg

    # ------------------------------------------------------------------ device plumbing --------
    def tables(self):
        """(table12, length_cum, border or None) for Engine.set_path."""
        w = self.waypoints
        t = _lib.path_table([p.x for p in w], [p.y for p in w], [p.psi for p in w], [p.kappa for p in w],
                            [p.v_ref for p in w])
        border = None
        if w[0].static_border_cells is not None:
            border = np.array([[p.static_border_cells[0][0], p.static_border_cells[0][1],
                                p.static_border_cells[1][0], p.static_border_cells[1][1]] for p in w])
        return t, np.cumsum(self.segment_lengths), border

    def _get_engine(self):
        """Engine used for the path's own grid queries (static width, update_path_constraints)."""
        if self._engine is None:
            self._engine = _lib.Engine()
            t, lc, border = self.tables()
            self._engine.set_path(t, lc, border, self.circular)
            self._engine_version = self.version
        if self._engine_map_version != self.map.version:
            self._engine.set_base_grid(self.map.data, self.map.origin, self.map.resolution)
            self._engine_map_version = self.map.version
        if getattr(self, "_engine_version", None) != self.version:
            t, lc, border = self.tables()
            self._engine.set_path(t, lc, border, self.circular)
            self._engine_version = self.version
        return self._engine

    def _compute_width(self, max_width):
        """Static drivable width left / right of every waypoint (rp.py:206-287) -- kernel K3b."""
        eng = self._get_engine()
        ub, lb, border = eng.compute_width(max_width)
        for k, wp in enumerate(self.waypoints):
            wp.ub = ub[k]
            wp.lb = lb[k]
            cells = ((border[k, 0], border[k, 1]), (border[k, 2], border[k, 3]))
            wp.static_border_cells = cells
            wp.dynamic_border_cells = cells
        self.version += 1

    # ------------------------------------------------------------------ speed profile ----------
    def set_speed_profile(self, v_ref):
        """Assign precomputed reference velocities (one per waypoint)."""
        assert len(v_ref) == self.n_waypoints
        for wp, v in zip(self.waypoints, v_ref):
            wp.v_ref = float(v)
        self.version += 1

    def compute_speed_profile(self, Constraints):
        """Reference velocity per waypoint from curvature and acceleration limits (rp.py:289-354):
        min 1/2 ||v||^2 - v_max' v  s.t. a_min <= D1 v <= a_max, v_min <= v <= v_max_dyn, solved by the
        device ADMM (speed_profile kernel)."""
        from .speed_profile import solve_speed_profile
        N = self.n_waypoints - 1
        li = np.array([self.get_waypoint(i + 1) - self.get_waypoint(i) for i in range(N)])
        ki = np.array([self.get_waypoint(i).kappa for i in range(N)])
        v_max = np.ones(N) * Constraints['v_max']
        v_dyn = np.sqrt(Constraints['ay_max'] / (np.abs(ki) + self.eps))
        v_max = np.where(v_dyn < v_max, v_dyn, v_max)
        v = solve_speed_profile(li, v_max, Constraints['v_min'], Constraints['a_min'], Constraints['a_max'])
        for i, wp in enumerate(self.waypoints[:-1]):
            wp.v_ref = v[i]
        self.waypoints[-1].v_ref = self.waypoints[-2].v_ref
        self.version += 1

    # ------------------------------------------------------------------ queries ----------------
    def get_waypoint(self, wp_id):
        """Circular indexing (rp.py:356-371)."""
        if wp_id >= self.n_waypoints and self.circular:
            wp_id = np.mod(wp_id, self.n_waypoints)
        elif wp_id >= self.n_waypoints and not self.circular:
            print('Reached end of path!')
            exit(1)
        return self.waypoints[wp_id]

    def update_path_constraints(self, wp_id, N, min_width, safety_margin):
        """Dynamic drivable corridor over N waypoints starting at wp_id (rp.py:522-648) -- kernel K3.
        Returns (ub, lb, border_cells_hor_sm) like the reference and updates dynamic_border_cells."""
        import torch
        eng = self._get_engine()
        dev = torch.device("cuda", torch.cuda.current_device())
        wid = torch.tensor([int(wp_id)], dtype=torch.int32, device=dev)
        ub = torch.empty((1, N), dtype=torch.float64, device=dev)
        lb = torch.empty((1, N), dtype=torch.float64, device=dev)
        cells = torch.empty((1, N, 4), dtype=torch.float64, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        eng.update_path_constraints(wid, 0, N, min_width, safety_margin, ub, lb, cells, flags)
        eng.sync()
        f = int(flags.item())
        if f & _lib.ST_NO_SEGMENT:
            raise ValueError("max() arg is an empty sequence")  # what the reference raises (rp.py:547)
        if f & _lib.ST_END_OF_PATH:
            print('Reached end of path!')
            exit(1)
        if f & _lib.ST_INDEX_ERROR:
            raise IndexError("ray left the occupancy grid")
        c = cells[0].cpu().numpy()
        border = [[(c[n, 0], c[n, 1]), (c[n, 2], c[n, 3])] for n in range(N)]
        for n in range(N):
            self.get_waypoint(wp_id + n).dynamic_border_cells = (border[n][0], border[n][1])
        return ub[0].cpu().numpy(), lb[0].cpu().numpy(), border

    def show(self, display_drivable_area=True):
        raise NotImplementedError("plotting is not part of the B200 engine")
