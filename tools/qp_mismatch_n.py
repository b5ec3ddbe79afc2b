"""tools/qp_mismatch_n.py N [N ...] -- fp32 ADMM kernels vs the oracle on 64 sweep-style QPs per horizon (the same
generator as tools/qp_sweep.py): lists the instances whose status / iteration count differ."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_b200
from oracle import oracle as orc
from conftest import Track, fixed_pattern
track = Track()
pt = orc.PathTables(track.wp_x, track.wp_y, track.wp_psi, track.wp_kappa, track.wp_vref, track.segment_lengths, track.border, True)
dev = torch.device("cuda:0")
kmax = np.tan(0.66) / 0.12; sm = 0.06 / np.sqrt(2)
want = [int(a) for a in sys.argv[1:]] or [10, 30, 50]
rng = np.random.default_rng(5)
for N in (10, 30, 50, 100):
    cfg = orc.mpc_cfg(N, [1.0, 0, 0], [0.5, 0], [1.0, 0, 0], [-np.inf] * 3, [np.inf] * 3, [0.0, -kmax], [1.0, kmax], 4.0, 0.12, sm)
    base = {k: [] for k in ("Pd", "q", "Ax", "l", "u")}
    while len(base["Pd"]) < 64:
        w = int(rng.integers(0, 200)); ey, eps_ = rng.uniform(-0.05, 0.05), rng.uniform(-0.1, 0.1)
        st, ub, lb, _ = orc.update_path_constraints(track.grid_obs, track.origin, track.res, pt, w + 1, N, 2 * sm, sm)
        if st:
            continue
        Pd, q, A, l, u = orc.mpc_assemble(pt, cfg, w, [ey, eps_, 0.0], np.zeros(2 * N), ub, lb)
        for k, v in zip(("Pd", "q", "Ax", "l", "u"), (Pd, q, A.data, l, u)):
            base[k].append(v)
    if N not in want:
        continue
    base = {k: np.array(v) for k, v in base.items()}
    Ap, Ai = fixed_pattern(N)
    xo, ito, sto = orc.batch_qp_solve(N, base["Pd"], base["q"], Ap, Ai, base["Ax"], base["l"], base["u"], eps_abs=1e-3, eps_rel=1e-3)
    eng = mpc_b200.Engine(N=N, precision=0, eps_abs=1e-3, eps_rel=1e-3)
    args = [torch.tensor(base[k], dtype=torch.float64, device=dev) for k in ("Pd", "q", "Ax", "l", "u")]
    x = torch.zeros((64, 5 * N + 3), dtype=torch.float64, device=dev)
    it = torch.zeros(64, dtype=torch.int32, device=dev); stt = torch.zeros(64, dtype=torch.int32, device=dev)
    eng.solve_qp(*args, x, it, stt); eng.sync()
    it, stt, x = it.cpu().numpy(), stt.cpu().numpy(), x.cpu().numpy()
    bad = np.nonzero((it != ito) | (stt != sto))[0]
    print("N=%d kernel=%s: %d of 64 differ" % (N, os.environ.get("MPC_ADMM_KERNEL", "pair"), len(bad)))
    for b in bad:
        print("   qp %2d: oracle (st %d, it %d)   gpu (st %d, it %d)" % (b, sto[b], ito[b], stt[b], it[b]))
    ok = (sto == 1) & (stt == 1)
    if ok.any():
        print("   max |x - oracle| over commonly solved: %.3e" % np.abs(x[ok] - xo[ok]).max())
    eng.close()
