"""pytest configuration: registers the `gpu` marker and provides fixtures shared by both tiers.

`-m "not gpu"` runs here without a GPU: the oracle against the golden vectors and (when
/root/reference exists) against the reference's own code, the host logic, the C-ABI export check.
`-m gpu` runs on the B200: the CUDA path through the C ABI against the oracle and the golden vectors.
Nothing in the gpu tier reads /root/reference.
"""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built library (it is git-ignored): build it once, in-tree, exactly as
    # __graft_entry__.build() does (nvcc cross-compiles for sm_100a without a GPU; ~1 min).  Building is not using:
    # the product still refuses to run without a CUDA device.
    lib = os.path.join(REPO, "multi-purpose-mpc_b200", "libmpc_b200.so")
    if not os.path.exists(lib):
        import importlib.util
        spec = importlib.util.spec_from_file_location("mpc_b200_build_ext", os.path.join(REPO, "multi-purpose-mpc_b200", "build_ext.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        try:
            mod.build()
        except Exception as e:  # no nvcc: the ABI tests will say so
            print("could not build libmpc_b200.so:", e, file=sys.stderr)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class Track:
    """The sim track of src/simulation.py as plain arrays (tests/golden/sim_track.npz)."""

    def __init__(self):
        T = np.load(os.path.join(GOLDEN, "sim_track.npz"))
        self.raw = T
        W = int(T["grid_shape"][1])
        self.grid = np.unpackbits(T["grid_bits"], axis=1)[:, :W].astype(np.int8)
        self.grid_obs = np.unpackbits(T["grid_obstacles_bits"], axis=1)[:, :W].astype(np.int8)
        self.origin = T["origin"]
        self.res = float(T["resolution"])
        self.n_wp = len(T["wp_x"])
        self.length_cum = np.cumsum(T["segment_lengths"])
        for k in ("wp_x", "wp_y", "wp_psi", "wp_kappa", "wp_vref", "wp_ub", "wp_lb", "border", "segment_lengths",
                  "obstacles", "corner_x", "corner_y"):
            setattr(self, k, T[k])
        self.length = float(T["length"])


@pytest.fixture(scope="session")
def track():
    return Track()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def orc_path(track, orc):
    return orc.PathTables(track.wp_x, track.wp_y, track.wp_psi, track.wp_kappa, track.wp_vref,
                          track.segment_lengths, track.border, True)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def ulps(a, b):
    """distance in units in the last place between two float64 arrays (-0.0 == +0.0)"""
    def ordered(v):
        i = np.ascontiguousarray(v, np.float64).view(np.int64)
        return np.where(i >= 0, i, -(i & np.int64(0x7FFFFFFFFFFFFFFF)))
    return np.abs(ordered(a) - ordered(b))


def fixed_pattern(N):
    """(Ap, Ai) of the structural CSC pattern of the reference's constraint matrix (MPC.py:128-135)."""
    nx, nu = 3, 2
    neq = nx * (N + 1)
    n = neq + nu * N
    rows, cols = [], []
    for col in range(neq):
        k, j = divmod(col, nx)
        rows.append(col); cols.append(col)
        if k < N:
            rr = {0: [0, 1, 2], 1: [0, 1], 2: [2]}[j]
            rows += [nx * (k + 1) + r for r in rr]; cols += [col] * len(rr)
        rows.append(neq + col); cols.append(col)
    for col in range(neq, n):
        k, j = divmod(col - neq, nu)
        rows.append(nx * (k + 1) + (2 if j == 0 else 1)); cols.append(col)
        rows.append(neq + col); cols.append(col)
    Ap = np.zeros(n + 1, np.int32)
    for c in cols:
        Ap[c + 1] += 1
    return np.cumsum(Ap).astype(np.int32), np.array(rows, np.int32)


def h1_null_direction(N, Pd, Ax):
    """with R[1] = 0 and QN[1] = 0 (the reference's weights) the QP of MPC._init_problem is not strictly
    convex -- (u_{N-1}.kappa : 1, x_N.e_psi : B[1,1]) costs nothing and violates no equality, so the minimiser is a
    segment and a solver's position on it depends on scaling, rho schedule and start point.  Returns the unit vector of that
    direction in the reference's x ordering for a batch of QPs ([B, 5N+3]; a zero row where the direction does not exist,
    e.g. R[1] > 0), read off the constraint values: the e_psi_N dynamics row is  a.x_{N-1} + b.kappa_{N-1} - e_psi_N = ...
    The direction is verified to lie in the null space of [diag(P); A_eq] before it is returned."""
    Ap, Ai = fixed_pattern(N)
    n, neq = 5 * N + 3, 3 * (N + 1)
    Pd = np.atleast_2d(Pd); Ax = np.atleast_2d(Ax)
    ik, ie = n - 1, 3 * N + 1                       # u_{N-1}.kappa, x_N.e_psi
    def entry(row, col):
        k = [j for j in range(Ap[col], Ap[col + 1]) if Ai[j] == row]
        assert len(k) == 1
        return Ax[:, k[0]]
    b = entry(ie, ik)                               # coefficient of kappa_{N-1} in the e_psi_N dynamics row
    c = entry(ie, ie)                               # -1: coefficient of e_psi_N in its own row
    d = np.zeros((Pd.shape[0], n))
    d[:, ik] = 1.0
    d[:, ie] = -b / c
    # any other equality row touching the two variables, or a cost on either, removes the direction
    free = (Pd[:, ik] == 0) & (Pd[:, ie] == 0)
    for col in (ik, ie):
        for j in range(Ap[col], Ap[col + 1]):
            if Ai[j] < neq and Ai[j] != ie:
                free &= Ax[:, j] == 0
    d[~free] = 0.0
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    return np.divide(d, nrm, out=np.zeros_like(d), where=nrm > 0)


def h1_split(N, Pd, Ax, err):
    """(remainder, null coordinate): err with the H1 direction projected out, and its component along it."""
    d = h1_null_direction(N, Pd, Ax)
    c = np.einsum("bi,bi->b", np.nan_to_num(err), d)
    return err - c[:, None] * d, c


def sim_cfg(orc, N=30):
    """orc_mpc_cfg of src/simulation.py:100-111"""
    kmax = np.tan(0.66) / 0.12
    return orc.mpc_cfg(N, [1.0, 0.0, 0.0], [0.5, 0.0], [1.0, 0.0, 0.0], [-np.inf] * 3, [np.inf] * 3, [0.0, -kmax],
                       [1.0, kmax], 4.0, 0.12, 0.06 / np.sqrt(2))


@pytest.fixture(scope="session")
def engine_factory(track):
    """Engine on the sim track (path + static border cells from the golden file, grid with the
    reference's nine obstacles unless told otherwise)."""
    import mpc_b200
    from mpc_b200 import _lib
    made = []

    def make(grid="obstacles", border=True, **cfg):
        eng = mpc_b200.Engine(**cfg)
        tab = _lib.path_table(track.wp_x, track.wp_y, track.wp_psi, track.wp_kappa, track.wp_vref)
        eng.set_path(tab, track.length_cum, track.border if border else None, True)
        eng.set_base_grid(track.grid_obs if grid == "obstacles" else track.grid, track.origin, track.res)
        made.append(eng)
        return eng

    yield make
    for e in made:
        e.close()
