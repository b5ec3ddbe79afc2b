"""tools/qp_sweep.py -- BASELINE config 5: QP-only sweep, batch size x horizon, GPU ADMM (K2 alone, mpc_solve_qp)
beside the CPU oracle port on all host cores.  QPs are real MPC QPs (oracle assembly on the sim track for random
states), cold start, eps = 1e-3 (fp32 path) and 1e-5 (fp64 path).  `python bench.py --workload sweep` runs sweep() and
prints it as one JSON line; run directly it writes gpurun_out/qp_sweep.json."""
import json, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))


def sweep(dev, horizons=(10, 30, 50, 100), batches=(1, 10, 100, 1000, 10000, 100000, 1000000), settings=((1e-3, 0), (1e-5, 1)),
          log=None):
    import torch
    import mpc_b200
    from oracle import oracle as orc
    from conftest import Track, fixed_pattern
    track = Track()
    pt = orc.PathTables(track.wp_x, track.wp_y, track.wp_psi, track.wp_kappa, track.wp_vref, track.segment_lengths,
                        track.border, True)
    rng = np.random.default_rng(5)
    kmax = np.tan(0.66) / 0.12
    sm = 0.06 / np.sqrt(2)
    out = []
    for N in horizons:
        cfg = orc.mpc_cfg(N, [1.0, 0, 0], [0.5, 0], [1.0, 0, 0], [-np.inf] * 3, [np.inf] * 3, [0.0, -kmax], [1.0, kmax], 4.0,
                          0.12, sm)
        n, m, nnz = 5 * N + 3, 8 * N + 6, 16 * N + 6
        base = {k: [] for k in ("Pd", "q", "Ax", "l", "u")}
        while len(base["Pd"]) < 64:
            w = int(rng.integers(0, 200)); ey, eps_ = rng.uniform(-0.05, 0.05), rng.uniform(-0.1, 0.1)
            st, ub, lb, _ = orc.update_path_constraints(track.grid_obs, track.origin, track.res, pt, w + 1, N, 2 * sm, sm)
            if st:
                continue
            cc = np.zeros(2 * N)
            Pd, q, A, l, u = orc.mpc_assemble(pt, cfg, w, [ey, eps_, 0.0], cc, ub, lb)
            for k, v in zip(("Pd", "q", "Ax", "l", "u"), (Pd, q, A.data, l, u)):
                base[k].append(v)
        base = {k: np.array(v) for k, v in base.items()}
        Ap, Ai = fixed_pattern(N)
        for eps, prec in settings:
            # CPU: oracle port, all threads, 256-QP sample
            idx = np.arange(256) % 64
            t0 = time.perf_counter()
            xo, ito, sto = orc.batch_qp_solve(N, base["Pd"][idx], base["q"][idx], Ap, Ai, base["Ax"][idx], base["l"][idx],
                                              base["u"][idx], eps_abs=eps, eps_rel=eps)
            cpu_rate = 256 / (time.perf_counter() - t0)
            eng = mpc_b200.Engine(N=N, precision=prec, eps_abs=eps, eps_rel=eps)
            dbase = {k: torch.tensor(v, dtype=torch.float64, device=dev) for k, v in base.items()}
            for B in batches:
                bytes_in = B * (2 * n + nnz + 2 * m) * 8
                if bytes_in > 24e9 or (prec == 1 and B > 100000):
                    continue
                gi = torch.arange(B, device=dev) % 64
                args = [dbase[k][gi].contiguous() for k in ("Pd", "q", "Ax", "l", "u")]
                it = torch.zeros(B, dtype=torch.int32, device=dev)
                stt = torch.zeros(B, dtype=torch.int32, device=dev)
                for _ in range(2):
                    eng.solve_qp(*args, None, it, stt)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3 if B <= 100000 else 1
                e0.record()
                for _ in range(reps):
                    eng.solve_qp(*args, None, it, stt)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                same = bool((it[:64].cpu().numpy() == ito[:64]).all()) if B >= 64 else None
                row = dict(N=N, eps=eps, precision="f32" if prec == 0 else "f64", B=B, ms=ms, solves_per_s=B / ms * 1e3,
                           mean_iters=float(it.float().mean().item()), iters_equal_oracle=same,
                           solved_frac=float((stt == 1).float().mean().item()), cpu_port_solves_per_s=cpu_rate,
                           cpu_threads=orc.num_threads(), cpu_mean_iters=float(ito.mean()))
                out.append(row)
                if log:
                    print(row, file=log, flush=True)
                del args
            eng.close()
    return out


if __name__ == "__main__":
    import torch
    rows = sweep(torch.device("cuda:0"), log=sys.stdout)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(REPO, "gpurun_out", "qp_sweep.json"), "w"), indent=1)
