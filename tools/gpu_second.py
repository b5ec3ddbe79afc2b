"""tools/gpu_second.py -- development aid: every kernel against the golden fixtures, one summary."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mpc_b200  # noqa: E402
from mpc_b200 import _lib  # noqa: E402

G = os.path.join(REPO, "tests", "golden")
T = np.load(os.path.join(G, "sim_track.npz"))
dev = torch.device("cuda:0")
grid = np.unpackbits(T["grid_bits"], axis=1)[:, :T["grid_shape"][1]].astype(np.int8)
grid_obs = np.unpackbits(T["grid_obstacles_bits"], axis=1)[:, :T["grid_shape"][1]].astype(np.int8)
origin, res = T["origin"], float(T["resolution"])
tab = _lib.path_table(T["wp_x"], T["wp_y"], T["wp_psi"], T["wp_kappa"], T["wp_vref"])
lc = np.cumsum(T["segment_lengths"])


def ulps(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a.view(np.int64) - b.view(np.int64))


# ---- static width (K3b) on the obstacle-free map
eng = mpc_b200.Engine(precision=1)
eng.set_path(tab, lc, None, True)
eng.set_base_grid(grid, origin, res)
ub, lb, border = eng.compute_width(0.23)
print("K3b ub exact", int((ub == T["wp_ub"]).sum()), "/", len(ub), "max ulp", int(ulps(ub, T["wp_ub"]).max()),
      "| lb max ulp", int(ulps(lb, T["wp_lb"]).max()), "| border exact", bool((border == T["border"]).all()))

# ---- speed profile
from mpc_b200.speed_profile import solve_speed_profile  # noqa: E402
n = len(T["wp_x"]) - 1
li = tab[5][:n]
vmax = np.minimum(1.0, np.sqrt(4.0 / (np.abs(T["wp_kappa"][:n]) + 1e-12)))
v, it, st = solve_speed_profile(li, vmax, 0.0, -0.1, 0.5, return_info=True)
print("speed profile iters", it, "status", st, "max |v - golden|", float(np.abs(v - T["wp_vref"][:n]).max()))

# ---- rasteriser + raycast on random obstacle sets
R = np.load(os.path.join(G, "raycast_random.npz"))
eng.set_path(tab, lc, T["border"], True)
nsc = len(R["obs_off"]) - 1
eng.set_obstacles(R["obs"], R["obs_off"])
gb = np.unpackbits(R["grid_bits"], axis=2)[:, :, :grid.shape[1]].astype(np.int8)
print("rasteriser grids exact:", all(bool((eng.get_grid(s) == gb[s]).all()) for s in range(nsc)))
cases = R["wp_id"]
# one launch per case group: scenario s, waypoint w  -> emulate with B = nsc, wp_id per scenario
tot, exact, maxulp = 0, 0, 0
for c in range(0, len(cases), 1):
    s, w, ok = cases[c]
    wid = torch.zeros(nsc, dtype=torch.int32, device=dev)
    wid[s] = int(w)
    ubt = torch.zeros((nsc, 30), dtype=torch.float64, device=dev)
    lbt = torch.zeros((nsc, 30), dtype=torch.float64, device=dev)
    cel = torch.zeros((nsc, 30, 4), dtype=torch.float64, device=dev)
    fl = torch.zeros(nsc, dtype=torch.int32, device=dev)
    eng.raycast(wid, ubt, lbt, cel, fl)
    torch.cuda.synchronize()
    u_, l_ = ubt[s].cpu().numpy(), lbt[s].cpu().numpy()
    tot += 60
    exact += int((u_ == R["ub"][c]).sum() + (l_ == R["lb"][c]).sum())
    maxulp = max(maxulp, int(ulps(u_, R["ub"][c]).max()), int(ulps(l_, R["lb"][c]).max()))
    if ulps(u_, R["ub"][c]).max() > 4 or ulps(l_, R["lb"][c]).max() > 4:
        print("  raycast MISMATCH case", c, s, w, "flags", int(fl[s].item()))
        print("   ub", u_[:8], R["ub"][c][:8])
print("K3 raycast: exact", exact, "/", tot, "max ulp", maxulp)
eng.close()

# ---- teacher-forced steps (all four kernels) fp64
TF = np.load(os.path.join(G, "teacher_forced.npz"))
for prec, refine in ((1, 0), (0, 1)):
    eng = mpc_b200.Engine(precision=prec)
    eng.set_path(tab, lc, T["border"], True)
    eng.set_base_grid(grid_obs, origin, res)
    B = TF["state"].shape[0]
    eng.scenarios_init(np.ascontiguousarray(TF["state"].T))
    eng.scenarios_set_state(np.ascontiguousarray(TF["state"].T), np.ascontiguousarray(TF["control"]), None)
    eng.step()
    o = eng.scenarios_read()
    print("teacher-forced prec", prec, "refine", refine)
    print("  wp_id equal", bool((o["wp_id"] == TF["wp_id"]).all()), "| ub exact", int((o["ub"] == TF["ub"]).sum()), "/",
          TF["ub"].size, "max ulp", int(ulps(o["ub"], TF["ub"]).max()), "| lb max ulp", int(ulps(o["lb"], TF["lb"]).max()))
    print("  status equal", int((o["qp_status"] == TF["status"]).sum()), "/", B, "iters equal",
          int((o["iters"] == TF["iters"]).sum()), "| max |u - ref|", float(np.abs(o["u"] - TF["u"]).max()))
    rel = np.abs(o["state"].T - TF["state_after"]) / np.maximum(np.abs(TF["state_after"]), 1e-9)
    print("  state_after max rel diff", float(rel.max()), "| control_after max diff",
          float(np.abs(o["control"] - TF["control_after"]).max()), "flags", sorted(set(o["flags"].tolist())))
    eng.close()

# ---- C1 closed loop, free running, fp64 vs the reference lap
C1 = np.load(os.path.join(G, "c1_lap.npz"))
for prec, refine in ((1, 0), (0, 1)):
    eng = mpc_b200.Engine(precision=prec)
    eng.set_path(tab, lc, T["border"], True)
    eng.set_base_grid(grid_obs, origin, res)
    st0 = np.ascontiguousarray(C1["state"][0].reshape(4, 1))
    eng.scenarios_init(st0)
    nst = C1["state"].shape[0]
    maxd, firstbad = 0.0, None
    for k in range(nst):
        eng.step()
        o = eng.scenarios_read()
        d = np.abs(o["state"][:, 0] - C1["state_after"][k]).max()
        maxd = max(maxd, d)
        if firstbad is None and (d > 1e-6 or o["iters"][0] != C1["iters"][k]):
            firstbad = (k, float(d), int(o["iters"][0]), int(C1["iters"][k]), int(o["qp_status"][0]), int(C1["status"][k]))
    print("C1 lap prec", prec, "refine", refine, "steps", nst, "max |state - ref|", maxd, "first divergence", firstbad,
          "final s", float(o["state"][3, 0]), "ref", float(C1["state_after"][-1][3]), "flags", int(o["flags"][0]))
    eng.close()
