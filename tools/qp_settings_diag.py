"""tools/qp_settings_diag.py -- fp32 kernels vs the oracle under non-default OSQP settings (diagnostic)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_b200
from oracle import oracle as orc
from conftest import load_golden, fixed_pattern
TF = load_golden("teacher_forced.npz")
ks = list(range(0, 48))
Pd, q, Ax, l, u = (TF["qp_" + k][ks] for k in ("Pd", "q", "Ax", "l", "u"))
Ap, Ai = fixed_pattern(30)
dev = torch.device("cuda:0")
t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
for settings in (dict(scaling=0), dict(alpha=1.0, rho=1.0), dict(rho=1.0), dict(alpha=1.0), dict(scaling=2)):
    xo, ito, sto = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u, **settings)
    eng = mpc_b200.Engine(precision=0, **settings)
    x = torch.zeros((len(ks), 153), dtype=torch.float64, device=dev)
    it = torch.zeros(len(ks), dtype=torch.int32, device=dev); st = torch.zeros(len(ks), dtype=torch.int32, device=dev)
    eng.solve_qp(t(Pd), t(q), t(Ax), t(l), t(u), x, it, st); eng.sync()
    it, st, x = it.cpu().numpy(), st.cpu().numpy(), x.cpu().numpy()
    eng.close()
    solved = sto == 1
    bad = np.nonzero(solved & ((it != ito) | (st != sto)))[0]
    both = solved & (st == 1)
    print(os.environ.get("MPC_ADMM_KERNEL", "pair"), settings, "solved", int(solved.sum()), "mismatching", len(bad),
          [(int(ito[b]), int(it[b]), int(st[b])) for b in bad[:8]], "max|x-xo| %.2e" % np.abs(x[both] - xo[both]).max())
