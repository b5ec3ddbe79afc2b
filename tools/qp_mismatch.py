"""tools/qp_mismatch.py -- fp32 ADMM kernel vs the oracle on the golden QPs: which instances differ in status / iterations."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_b200
from oracle import oracle as orc
from conftest import load_golden, fixed_pattern
TF, C1 = load_golden("teacher_forced.npz"), load_golden("c1_lap.npz")
Pd, q, Ax, l, u = (np.concatenate([TF["qp_" + k], C1["qp_" + k]]) for k in ("Pd", "q", "Ax", "l", "u"))
B, n = Pd.shape[0], 153
Ap, Ai = fixed_pattern(30)
eps = 1e-3
xo, ito, sto = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u, eps_abs=eps, eps_rel=eps)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
eng = mpc_b200.Engine(precision=0, eps_abs=eps, eps_rel=eps)
x = torch.zeros((B, n), dtype=torch.float64, device=dev)
it = torch.zeros(B, dtype=torch.int32, device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
eng.solve_qp(t(Pd), t(q), t(Ax), t(l), t(u), x, it, st); eng.sync()
x, it, st = x.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
bad = np.nonzero((it != ito) | (st != sto))[0]
print("mismatches:", len(bad), "of", B)
for b in bad:
    print("  qp %d: oracle (st %d, it %d)  gpu (st %d, it %d)" % (b, sto[b], ito[b], st[b], it[b]))
ok = ~np.isin(sto, (-3, -4, -7, 3, 4)) & ~np.isin(st, (-3, -4, -7, 3, 4))
d = np.abs(x[ok] - xo[ok]).max(axis=1)
print("max |x - oracle| over solved: %.3e  (mean %.3e)" % (d.max(), d.mean()))
