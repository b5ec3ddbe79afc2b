"""tools/c4_lap.py [B] [N] -- BASELINE config 4 style: time-optimal weights (build-defined), horizon N (default
50), B scenarios with randomised start offsets, closed loop until every car has finished its lap (or the step cap);
reports steps, wall time, QP solves/s and the statistics vector.  fp32 production path."""
import json, os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_b200
from mpc_b200 import _lib
from conftest import Track
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
T = Track()
rng = np.random.default_rng(4)
tab = _lib.path_table(T.wp_x, T.wp_y, T.wp_psi, T.wp_kappa, T.wp_vref)
lc = np.cumsum(T.segment_lengths)
eng = mpc_b200.Engine(N=N, precision=0, Q=[0.1, 0.0, 0.0], R=[0.01, 0.0], QN=[0.1, 0.0, 5.0])
eng.set_path(tab, lc, T.border, True)
eng.set_base_grid(T.grid, T.origin, float(T.res))
ey = rng.uniform(-0.03, 0.03, B); ep = rng.uniform(-0.05, 0.05, B)
st = np.stack([T.wp_x[0] - ey * np.sin(T.wp_psi[0]), T.wp_y[0] + ey * np.cos(T.wp_psi[0]), T.wp_psi[0] + ep, np.zeros(B)])
eng.scenarios_init(np.ascontiguousarray(st))
torch.cuda.synchronize()
t0 = time.perf_counter()
total = None
steps = 0
while steps < 1200:
    s = eng.run_closed_loop(50)
    steps += 50
    if s["finished"] + s["dead"] >= B:
        total = s
        break
    total = s
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out = eng.scenarios_read()
print(json.dumps({"workload": "time-optimal weights, N=%d, %d scenarios, one lap closed loop (BASELINE configs[3] style, 1 GPU share)" % (N, B),
                  "steps_run": steps, "wall_s": dt, "qp_solves": total["qp_solves"], "qp_solves_per_s": total["qp_solves"] / dt,
                  "finished": total["finished"], "dead": total["dead"], "qp_fallbacks": total["qp_fallbacks"],
                  "mean_admm_iters": total["admm_iters"] / max(total["qp_solves"], 1), "max_abs_ey": total["max_abs_ey"],
                  "mean_lap_s": float(np.mean(out["state"][3] > 0))}))
