"""oracle/ref_harness.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imports the UNMODIFIED reference modules from /root/reference/src in this container so that the
reference's own Python control flow produces golden data.  Three things stop the reference from
importing as-is here and are shimmed, nothing else is touched:
  * matplotlib missing            -> oracle/ref_shims/matplotlib (no-op; visualisation only)
  * skimage / osqp missing        -> oracle/ref_shims/{skimage,osqp}: the oracle's RESTATEMENTS
                                     (if the real packages import, they are preferred and recorded)
  * scipy >= 1.14 removed `.A`    -> property alias to .toarray() on scipy sparse matrices
/root/reference exists only in the build container: nothing on the GPU box imports this module.
"""
import importlib
import os
import sys

REF_SRC = "/root/reference/src"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROVENANCE = {}


def available():
    return os.path.isdir(REF_SRC)


def _real(modname):
    try:
        saved = list(sys.path)
        sys.path = [p for p in sys.path if os.path.abspath(p) != _SHIMS]
        try:
            m = importlib.import_module(modname)
        finally:
            sys.path = saved
        return m
    except Exception:
        return None


def load():
    """Returns a namespace with the reference's Map, Obstacle, ReferencePath, BicycleModel, MPC."""
    if not available():
        raise RuntimeError("reference sources not present (GPU box?)")
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    for name in ("osqp", "skimage", "matplotlib"):
        real = _real(name) if name not in sys.modules else sys.modules[name]
        if real is not None and "ref_shims" not in (getattr(real, "__file__", "") or ""):
            PROVENANCE[name] = "real %s %s" % (name, getattr(real, "__version__", "?"))
        else:
            PROVENANCE[name] = "oracle restatement (NOT %s)" % name
    if _SHIMS not in sys.path:
        sys.path.append(_SHIMS)  # appended: real packages win if they exist
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    # scipy `.A` alias (MPC.py:153,155 use self.Q.A)
    from scipy import sparse
    for cls_name in ("spmatrix", "dia_matrix", "csc_matrix", "csr_matrix", "coo_matrix"):
        cls = getattr(sparse, cls_name, None)
        if cls is not None and not hasattr(cls, "A"):
            try:
                setattr(cls, "A", property(lambda self: self.toarray()))
            except Exception:
                pass

    class NS:
        pass

    ns = NS()
    ns.map = importlib.import_module("map")
    ns.reference_path = importlib.import_module("reference_path")
    ns.sbm = importlib.import_module("spatial_bicycle_models")
    ns.MPC_mod = importlib.import_module("MPC")
    ns.Map, ns.Obstacle = ns.map.Map, ns.map.Obstacle
    ns.ReferencePath, ns.Waypoint = ns.reference_path.ReferencePath, ns.reference_path.Waypoint
    ns.BicycleModel = ns.sbm.BicycleModel
    ns.MPC = ns.MPC_mod.MPC
    ns.osqp = sys.modules["osqp"]
    ns.provenance = dict(PROVENANCE)
    return ns


# ---- the reference's default configuration (src/simulation.py:20-54,100-119), as data ----
SIM = dict(
    map_file=os.path.join(REF_SRC, "maps", "sim_map.png"), origin=[-1, -2], resolution=0.005,
    wp_x=[-0.75, -0.25, -0.25, 0.25, 0.25, 1.25, 1.25, 0.75, 0.75, 1.25, 1.25, -0.75, -0.75, -0.25],
    wp_y=[-1.5, -1.5, -0.5, -0.5, -1.5, -1.5, -1, -1, -0.5, -0.5, 0, 0, -1.5, -1.5],
    path_resolution=0.05, smoothing_distance=5, max_width=0.23, circular=True,
    obstacles=[(0.0, 0.0, 0.05), (-0.8, -0.5, 0.08), (-0.7, -1.5, 0.05), (-0.3, -1.0, 0.08),
               (0.27, -1.0, 0.05), (0.78, -1.47, 0.05), (0.73, -0.9, 0.07), (1.2, 0.0, 0.08),
               (0.67, -0.05, 0.06)],
    car_length=0.12, car_width=0.06, Ts=0.05, N=30, Q=[1.0, 0.0, 0.0], R=[0.5, 0.0], QN=[1.0, 0.0, 0.0],
    v_max=1.0, delta_max=0.66, ay_max=4.0, a_min=-0.1, a_max=0.5)


def build_sim(ns, use_obstacles=True, N=None, obstacles=None):
    """Construct the reference objects exactly as src/simulation.py:20-119 does."""
    import numpy as np
    from scipy import sparse
    c = SIM
    mp = ns.Map(file_path=c["map_file"], origin=c["origin"], resolution=c["resolution"])
    rp = ns.ReferencePath(mp, c["wp_x"], c["wp_y"], c["path_resolution"],
                          smoothing_distance=c["smoothing_distance"], max_width=c["max_width"],
                          circular=c["circular"])
    if use_obstacles:
        obs = c["obstacles"] if obstacles is None else obstacles
        mp.add_obstacles([ns.Obstacle(cx=o[0], cy=o[1], radius=o[2]) for o in obs])
    car = ns.BicycleModel(length=c["car_length"], width=c["car_width"], reference_path=rp, Ts=c["Ts"])
    N = c["N"] if N is None else N
    Q, R, QN = sparse.diags(c["Q"]), sparse.diags(c["R"]), sparse.diags(c["QN"])
    ic = {"umin": np.array([0.0, -np.tan(c["delta_max"]) / car.length]),
          "umax": np.array([c["v_max"], np.tan(c["delta_max"]) / car.length])}
    sc = {"xmin": np.array([-np.inf, -np.inf, -np.inf]), "xmax": np.array([np.inf, np.inf, np.inf])}
    mpc = ns.MPC(car, N, Q, R, QN, sc, ic, c["ay_max"])
    spc = {"a_min": c["a_min"], "a_max": c["a_max"], "v_min": 0.0, "v_max": c["v_max"],
           "ay_max": c["ay_max"]}
    car.reference_path.compute_speed_profile(spc)
    return mp, rp, car, mpc
