"""tools/qp_random_parity.py [n] -- K2 parity on a RANDOM sample of real MPC QPs (not the golden fixture): random waypoint,
lateral / heading offsets, previous plans and drivable widths from the oracle's raycast on the obstacle map, assembled by
the oracle (MPC._init_problem restatement), solved by the oracle's OSQP restatement on all host cores and by mpc_solve_qp in
fp64 and fp32.  Prints one JSON line of agreement statistics."""
import json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch, mpc_b200
from conftest import Track, fixed_pattern, h1_split, sim_cfg
from oracle import oracle as orc


def make_qps(n, seed=5, N=30):
    T = Track()
    pt = orc.PathTables(T.wp_x, T.wp_y, T.wp_psi, T.wp_kappa, T.wp_vref, T.segment_lengths, T.border, True)
    cfg = sim_cfg(orc, N)
    rng = np.random.default_rng(seed)
    sm = 0.06 / np.sqrt(2)
    rows = []
    while len(rows) < n:
        wp = int(rng.integers(0, T.n_wp))
        x0 = np.array([rng.uniform(-0.06, 0.06), rng.uniform(-0.15, 0.15), 0.0])
        cc = np.zeros(2 * N)
        if rng.random() < 0.7:  # a previous plan: speeds and steering angles (MPC.current_control)
            cc[0::2] = rng.uniform(0.2, 1.0, N); cc[1::2] = rng.uniform(-0.5, 0.5, N)
        st, ub, lb, _ = orc.update_path_constraints(T.grid_obs, T.origin, T.res, pt, wp + 1, N, 2 * sm, sm)
        if st != 0:
            continue
        Pd, q, A, l, u = orc.mpc_assemble(pt, cfg, wp, x0, cc, ub, lb)
        rows.append((Pd, q, np.asarray(A.data, dtype=np.float64), l, u))
    return [np.ascontiguousarray(np.stack([r[i] for r in rows])) for i in range(5)]


def solve_gpu(precision, eps, Pd, q, Ax, l, u):
    dev = torch.device("cuda:0")
    t = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
    B = Pd.shape[0]
    eng = mpc_b200.Engine(precision=precision, eps_abs=eps, eps_rel=eps)
    x = torch.zeros((B, Pd.shape[1]), dtype=torch.float64, device=dev)
    it = torch.zeros(B, dtype=torch.int32, device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
    eng.solve_qp(t(Pd), t(q), t(Ax), t(l), t(u), x, it, st)
    eng.sync()
    out = x.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
    eng.close()
    return out


def compare(n=1024, seed=5):
    Pd, q, Ax, l, u = make_qps(n, seed)
    Ap, Ai = fixed_pattern(30)
    res = {"n": n, "seed": seed}
    is_kappa = np.zeros(153, bool); is_kappa[3 * 31 + 1::2] = True
    for precision, eps in ((1, 1e-5), (1, 1e-3), (0, 1e-3)):
        xo, ito, sto = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u, eps_abs=eps, eps_rel=eps)
        x, it, st = solve_gpu(precision, eps, Pd, q, Ax, l, u)
        ok = ~np.isin(sto, (-3, -4, -7, 3, 4)) & (st == sto)
        err = np.abs(x[ok] - xo[ok])
        rel_k = (err[:, is_kappa] / np.maximum(1.0, np.abs(xo[ok][:, is_kappa]))).max() if ok.any() else 0.0
        same = (st == sto) & (it == ito) & (sto == 1)
        e2 = np.abs(x[same] - xo[same])
        diff = (st == sto) & (it != ito) & (sto == 1)
        e3 = np.abs(x[diff] - xo[diff]) if diff.any() else np.zeros((1, 153))
        relk = e2[:, is_kappa] / np.maximum(1.0, np.abs(xo[same][:, is_kappa]))
        rem, null = h1_split(30, Pd[same], Ax[same], x[same] - xo[same])
        extra = dict(same_trace_max_err_h1_projected=float(np.abs(rem).max()),
                     same_trace_p999_err_h1_projected=float(np.quantile(np.abs(rem).max(axis=1), 0.999)),
                     same_trace_max_h1_null_coordinate=float(np.abs(null).max()),
                     same_trace=float(same.mean()), same_trace_max_err_states_v=float(e2[:, ~is_kappa].max()),
                     same_trace_max_err_kappa=float(e2[:, is_kappa].max()), same_trace_max_rel_err_kappa=float(relk.max()),
                     same_trace_p999_rel_err_kappa=float(np.quantile(relk.max(axis=1), 0.999)),
                     n_solved_other_iters=int(diff.sum()), other_iters_max_err=float(e3.max()),
                     other_iters=[[int(a), int(b)] for a, b in zip(it[diff][:8], ito[diff][:8])],
                     status_pairs=[[int(a), int(b)] for a, b in zip(st[st != sto][:8], sto[st != sto][:8])])
        res["%s_eps%g" % ("fp64" if precision else "fp32", eps)] = dict(**extra, **dict(
            status_equal=float((st == sto).mean()), iters_equal=float((it == ito).mean()),
            iters_equal_where_solved=float((it == ito)[sto == 1].mean()), solved=float((sto == 1).mean()),
            infeasible=float(np.isin(sto, (-3, -4)).mean()), max_err_states_v=float(err[:, ~is_kappa].max()),
            max_err_kappa=float(err[:, is_kappa].max()), max_rel_err_kappa=float(rel_k), mean_iters=float(ito.mean())))
    return res


if __name__ == "__main__":
    print(json.dumps(compare(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)))
