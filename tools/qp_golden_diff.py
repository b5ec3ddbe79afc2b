"""tools/qp_golden_diff.py -- the 76 golden QPs through mpc_solve_qp (fp32) against the oracle: which QPs differ in status /
iteration count, and the error statistics on identical traces."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch, mpc_b200
from conftest import fixed_pattern, h1_split, load_golden
from oracle import oracle as orc
TF, C1 = load_golden("teacher_forced.npz"), load_golden("c1_lap.npz")
Pd, q, Ax, l, u = (np.concatenate([TF["qp_" + k], C1["qp_" + k]]) for k in ("Pd", "q", "Ax", "l", "u"))
Ap, Ai = fixed_pattern(30)
xo, ito, sto = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
eng = mpc_b200.Engine(precision=0)
B = Pd.shape[0]
x = torch.zeros((B, 153), dtype=torch.float64, device=dev)
it = torch.zeros(B, dtype=torch.int32, device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
eng.solve_qp(t(Pd), t(q), t(Ax), t(l), t(u), x, it, st); eng.sync()
x, it, st = x.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
for b in np.nonzero((it != ito) | (st != sto))[0]:
    print("QP %d: gpu (status %d, %d passes)  oracle (status %d, %d passes)" % (b, st[b], it[b], sto[b], ito[b]))
same = (st == sto) & (it == ito) & (sto == 1)
rem, null = h1_split(30, Pd[same], Ax[same], x[same] - xo[same])
e = np.abs(rem).max(axis=1)
print("identical traces: %d of %d solved; max |x - oracle| %.3e, median %.3e, null coordinate %.3e" % (same.sum(), (sto == 1).sum(), e.max(), np.median(e), np.abs(null).max()))
