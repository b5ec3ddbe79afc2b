"""tools/iters_hist.py -- histogram of ADMM iteration counts per closed-loop step on the bench workload (4096 cars)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, mpc_b200
from mpc_b200 import _lib
T, grid = bench.load_track()
st = bench.scenario_states(T, 4096, 0, 4096)
e = mpc_b200.Engine(precision=0)
e.set_path(_lib.path_table(T["wp_x"], T["wp_y"], T["wp_psi"], T["wp_kappa"], T["wp_vref"]), np.cumsum(T["segment_lengths"]), T["border"], True)
e.set_base_grid(grid, T["origin"], float(T["resolution"]))
e.scenarios_init(st)
for k in range(12):
    e.step()
    it = e.scenarios_read()["iters"]
    print(k, dict(zip(*[x.tolist() for x in np.unique(it, return_counts=True)])))
