import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_b200
from mpc_b200 import _lib, distributed as D
from conftest import Track
track = Track()
B = 4096
sc = D.make_scenarios(track.n_wp, B, seed=2)
w = sc["start_wp"]
st = np.stack([track.wp_x[w] - sc["e_y"] * np.sin(track.wp_psi[w]), track.wp_y[w] + sc["e_y"] * np.cos(track.wp_psi[w]),
               track.wp_psi[w] + sc["e_psi"], track.length_cum[w]])
def run(st, prec):
    eng = mpc_b200.Engine(precision=prec)
    tab = _lib.path_table(track.wp_x, track.wp_y, track.wp_psi, track.wp_kappa, track.wp_vref)
    eng.set_path(tab, track.length_cum, track.border, True)
    eng.set_base_grid(track.grid, track.origin, track.res)
    eng.scenarios_init(np.ascontiguousarray(st)); eng.step(); o = eng.scenarios_read(); eng.close(); return o
perm = np.random.default_rng(0).permutation(B)
for prec in (0, 1):
    o1 = run(st, prec); o1b = run(st, prec); o2 = run(st[:, perm], prec)
    for k in ("u", "iters", "qp_status", "wp_id", "ub", "lb", "flags"):
        a, b = o1[k][perm], o2[k]
        same_rerun = np.array_equal(o1[k], o1b[k], equal_nan=True) if o1[k].dtype.kind == 'f' else np.array_equal(o1[k], o1b[k])
        neq = (a != b) & ~((a != a) & (b != b)) if a.dtype.kind == 'f' else (a != b)
        print("prec", prec, k, "rerun-equal", same_rerun, "perm-mismatch rows", int(np.any(neq.reshape(B, -1), axis=1).sum()))
    bad = np.nonzero(np.any((o1["u"][perm] != o2["u"]).reshape(B, -1), axis=1))[0][:5]
    for i in bad:
        print("  row", i, "orig idx", perm[i], o1["u"][perm][i], o2["u"][i], o1["iters"][perm][i], o2["iters"][i], o1["flags"][perm][i], o2["flags"][i], "wp", o2["wp_id"][i])
o = run(st, 0)
import collections
print("status hist", collections.Counter(o["qp_status"].tolist()), "flags", collections.Counter(o["flags"].tolist()))
s1 = o["qp_status"] == 1
print("u0 min", o["u"][s1, 0].min(), "max |delta|", np.abs(o["u"][s1, 1]).max(), "iters mean", o["iters"].mean())
