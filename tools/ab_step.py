"""tools/ab_step.py [--batch B] [--steps K] -- quick A/B timing of the closed-loop step for the library named by
MPC_B200_LIB (default: the in-tree .so): device ms/step with the L2 flushed between steps (as bench.py) and the
per-kernel times of the profiling arm.  One JSON line."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
import torch, mpc_b200
from mpc_b200 import _lib
T, grid = bench.load_track()
st = bench.scenario_states(T, a.batch, 0, a.batch)
e = mpc_b200.Engine(precision=0)
e.set_path(_lib.path_table(T["wp_x"], T["wp_y"], T["wp_psi"], T["wp_kappa"], T["wp_vref"]), np.cumsum(T["segment_lengths"]), T["border"], True)
e.set_base_grid(grid, T["origin"], float(T["resolution"]))
e.scenarios_init(st)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
for _ in range(5):
    e.step()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for k in range(a.steps):
    flush.fill_(float(k)); ev[k][0].record(); e.step(); ev[k][1].record()
torch.cuda.synchronize()
ms = sorted(x.elapsed_time(y) for x, y in ev)
e.set_profiling(True); e.run_closed_loop(a.steps); prof, nl = e.get_profile(); e.set_profiling(False)
out = e.scenarios_read()
print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "B": a.batch, "ms_mean": float(np.mean(ms)), "ms_median": float(np.median(ms)),
                  "ms_min": ms[0], "solve_ms_warm": prof["assemble_solve"] / max(nl[2], 1), "raycast_ms_warm": prof["raycast"] / max(nl[1], 1),
                  "mean_iters": float(out["iters"].mean()), "sum_u": float(np.abs(out["u"]).sum())}))
