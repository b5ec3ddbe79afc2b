"""tools/warm_start_study.py -- SURVEY 8f-4: would warm-starting the ADMM from the shifted previous plan (a deliberate,
opt-in departure: the reference builds a new OSQP object every step and therefore cold-starts, MPC.py:161-222) pay?  CPU only:
the fp64 stage-form OSQP iteration of tools/admm_pcr_model.py with an optional start point (x0, y0, rho0), driven along 120
closed-loop steps of the reference lap (oracle stepping, QPs assembled by the oracle = MPC._init_problem).
Result (profiles/r2_warm_start_study.json): NO gain.  With OSQP's check interval of 25 a solve takes 25 or 50 passes cold
(mean 48.3) and 50 warm (x only, or x + y + rho carried over); with a check every 5 passes 30.9 cold against 31.2 / 31.9 warm.
The sanity leg shows the start point is honoured (a QP restarted from its own solution stops at the first check): the shifted
plan is simply no closer to the next QP's ADMM fixed point than zero is -- consecutive solutions differ by 0.1 .. 0.7 in the
curvature inputs, the directions the cost barely sees (tools/precision_study.py) -- so the kernels keep the cold start and no
warm-start switch is shipped."""
import json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "tools"))
from conftest import Track, sim_cfg
from oracle import oracle as orc
from admm_pcr_model import *
T = Track(); N = 30
pt = orc.PathTables(T.wp_x, T.wp_y, T.wp_psi, T.wp_kappa, T.wp_vref, T.segment_lengths, T.border, True)
cfg = sim_cfg(orc, N)
world = orc.World(pt, cfg, T.grid.shape, T.origin, T.res, 0.05)
grid = T.grid

def admm_ws(N, Pd, q, Ax, l, u, x0=None, y0=None, rho0=None, chk=25, **kw):
    """fp64 OSQP iteration in stage form with optional warm start; returns x, y (unscaled), iters, status"""
    dt = np.dtype(np.float64); Tt = dt.type
    rho=0.1 if rho0 is None else rho0; sigma=1e-6; alpha=1.6; eps_abs=eps_rel=1e-3; eps_prim_inf=1e-4; max_iter=4000
    s = from_reference_layout(N, Pd, q, Ax, l, u, dt)
    L = N + 1
    mask = np.ones((L, 5), bool); mask[N, 3:] = False; s["mask"] = mask
    s["e"][N, 3:] = 0; s["lo"][N, 3:] = 0; s["hi"][N, 3:] = 0
    ruiz(s, 10, 5*N+3)
    a, c, e = s["a"], s["c"], s["e"]; D, Ed, Eb, cs = s["D"], s["Ed"], s["Eb"], s["cs"]
    lo, hi, dd, qq, P = s["lo"], s["hi"], s["d"], s["q"], s["P"]
    thr = Tt(OSQP_INFTY * MIN_SCALING)
    ctype = np.where((lo < -thr) & (hi > thr), -1, np.where(hi - lo < Tt(RHO_TOL), 1, 0)); ctype[N, 3:] = 1
    rho_vec = lambda r: np.where(ctype == -1, RHO_MIN, np.where(ctype == 1, RHO_EQ_OVER_RHO_INEQ * r, r))
    rb = rho_vec(rho)
    s["P"][N, 3:] = 1.0; factorize(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    rd = RHO_EQ_OVER_RHO_INEQ * rho
    x = np.zeros((L, 5)); zd = np.zeros((L, 3)); zb = np.zeros((L, 5)); yd = np.zeros((L, 3)); yb = np.zeros((L, 5))
    if x0 is not None:
        x = np.where(mask, x0 / D, 0)
        zd, zb = A_apply(s, x)
    if y0 is not None:
        yd = y0[0] / Ed * cs; yb = np.where(mask, y0[1] / Eb * cs, 0)
    status = 0
    for it in range(1, max_iter + 1):
        rhs = np.where(mask, sigma * x - qq + At_apply(s, rd * zd - yd, rb * zb - yb), 0)
        xt = solve(s, rhs)
        ztd, ztb = A_apply(s, xt)
        xn = alpha * xt + (1 - alpha) * x
        vd = alpha * ztd + (1 - alpha) * zd; vb = alpha * ztb + (1 - alpha) * zb
        zdn = dd.copy(); zbn = np.minimum(np.maximum(vb + yb / rb, lo), hi)
        dyd = rd * (vd + yd / rd - zdn) - yd  # y+ = y + rho(v - z+)  (eq rows: v includes y/rho?)  -> use OSQP form
        yd_n = yd + rd * (vd - zdn); yb_n = yb + rb * (vb - zbn)
        dyd = yd_n - yd; dyb = yb_n - yb
        x, zd, zb, yd, yb = xn, zdn, zbn, yd_n, yb_n
        if it % chk == 0:
            Axd, Axb = A_apply(s, x)
            rpd, rpb = Axd - zd, Axb - zb
            Px = P * x; Aty = At_apply(s, yd, yb)
            rdual = np.where(mask, Px + qq + Aty, 0)
            pri = max(np.abs(rpd / Ed).max(), np.abs(rpb / Eb).max()); dua = np.abs(rdual / D).max() / cs
            eps_prim = eps_abs + eps_rel * max(np.abs(zd / Ed).max(), np.abs(zb / Eb).max(), np.abs(Axd / Ed).max(), np.abs(Axb / Eb).max())
            eps_dual = eps_abs + eps_rel * max(np.abs(qq / D).max(), np.abs(Aty / D).max(), np.abs(Px / D).max()) / cs
            if pri < eps_prim and dua < eps_dual: status = 1; break
            if it >= 1000: break
            pn = max(np.abs(rpd).max(), np.abs(rpb).max()) / (max(np.abs(zd).max(), np.abs(zb).max(), np.abs(Axd).max(), np.abs(Axb).max()) + 1e-10)
            dn = np.abs(rdual).max() / (max(np.abs(qq).max(), np.abs(Aty).max(), np.abs(Px).max()) + 1e-10)
            rnew = min(max(rho * np.sqrt(pn / (dn + 1e-10)), RHO_MIN), RHO_MAX)
            if rnew > rho * 5 or rnew < rho / 5:
                rho = rnew; rb = rho_vec(rho); rd = RHO_EQ_OVER_RHO_INEQ * rho
                s["P"][N, 3:] = 1.0; factorize(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    return dict(x=D * x, yd=Ed * yd / cs, yb=Eb * yb / cs, iter=it, status=status, rho=rho)

def shift(arr, k):
    if k <= 0: return arr
    out = arr.copy(); out[:-k] = arr[k:]; out[-k:] = arr[-1]; return out


def lap(chk, steps=120):
    st = np.array([T.wp_x[0], T.wp_y[0] + 0.03, T.wp_psi[0] + 0.05, 0.0]); ctrl = np.zeros(2 * N); inf = 0
    prev = prev_wp = None
    tot = {"cold": 0, "warm_x": 0, "warm_x_y_rho": 0}; n = 0; dk = []
    for k in range(steps):
        r = world.step(grid, st, ctrl, inf, drive=False)
        Pd, q, A, l, u = orc.mpc_assemble(pt, cfg, r["wp_id"], r["spatial"], ctrl, r["ub"], r["lb"])
        Ax = np.asarray(A.data)
        cold = admm_ws(N, Pd, q, Ax, l, u, chk=chk)
        if prev is not None:
            sh = r["wp_id"] - prev_wp
            x0 = shift(prev["x"], sh)
            wx = admm_ws(N, Pd, q, Ax, l, u, x0=x0, chk=chk)
            wxy = admm_ws(N, Pd, q, Ax, l, u, x0=x0, y0=(shift(prev["yd"], sh), shift(prev["yb"], sh)), rho0=prev["rho"], chk=chk)
            tot["cold"] += cold["iter"]; tot["warm_x"] += wx["iter"]; tot["warm_x_y_rho"] += wxy["iter"]; n += 1
            dk.append(float(np.abs(x0[:, 4] - cold["x"][:, 4]).max()))
        prev, prev_wp = cold, r["wp_id"]
        rr = world.step(grid, st, ctrl, inf, drive=True)
        st, ctrl, inf = rr["state"], rr["current_control"], rr["infeas"]
    out = {k: v / n for k, v in tot.items()}
    out["max_kappa_distance_shifted_plan_to_solution_median"] = float(np.median(dk))
    return out


def main():
    out = {}
    st = np.array([T.wp_x[0], T.wp_y[0] + 0.03, T.wp_psi[0] + 0.05, 0.0]); ctrl = np.zeros(2 * N)
    r = world.step(grid, st, ctrl, 0, drive=False)
    Pd, q, A, l, u = orc.mpc_assemble(pt, cfg, r["wp_id"], r["spatial"], ctrl, r["ub"], r["lb"]); Ax = np.asarray(A.data)
    cold = admm_ws(N, Pd, q, Ax, l, u)
    own = admm_ws(N, Pd, q, Ax, l, u, x0=cold["x"], y0=(cold["yd"], cold["yb"]), rho0=cold["rho"])
    out["sanity_restart_from_own_solution"] = {"cold_passes": cold["iter"], "restarted_passes": own["iter"]}
    for chk in (25, 5):
        out["mean_passes_check_every_%d" % chk] = lap(chk)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
