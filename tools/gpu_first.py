"""tools/gpu_first.py -- first on-GPU sanity run (development aid): QP-only parity against the golden
vectors / oracle and rough timing of the ADMM kernel.  Writes gpurun_out/gpu_first.json."""
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mpc_b200  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from scipy import sparse  # noqa: E402

out = {}
G = np.load(os.path.join(REPO, "tests", "golden", "teacher_forced.npz"))
N = 30
n, m, nnz = 5 * N + 3, 8 * N + 6, 16 * N + 6
Pd, q, Ax, l, u, xg = (G["qp_" + k] for k in ("Pd", "q", "Ax", "l", "u", "x"))
B0 = Pd.shape[0]
dev = torch.device("cuda:0")
tt = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


def run(eng, Pd, q, Ax, l, u):
    B = Pd.shape[0]
    x = torch.zeros((B, n), dtype=torch.float64, device=dev)
    it = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    eng.solve_qp(tt(Pd), tt(q), tt(Ax), tt(l), tt(u), x, it, st)
    torch.cuda.synchronize()
    return x.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()


for prec, refine, eps in ((1, 0, 1e-3), (1, 0, 1e-5), (0, 1, 1e-3), (0, 0, 1e-3), (0, 1, 1e-5)):
    eng = mpc_b200.Engine(precision=prec, eps_abs=eps, eps_rel=eps)
    x, it, st = run(eng, Pd, q, Ax, l, u)
    key = "prec%d_ref%d_eps%g" % (prec, refine, eps)
    out[key] = dict(iters=it.tolist(), status=st.tolist())
    eng.close()
    out[key]["x"] = x.tolist()

# oracle comparison through the batch API (pattern from the CSC walk)
sys.path.insert(0, os.path.join(REPO, "oracle"))
from make_golden import fixed_pattern  # noqa: E402
rows_, cols_ = fixed_pattern(N)
Ap = np.zeros(n + 1, np.int32)
for c_ in cols_:
    Ap[c_ + 1] += 1
Ap = np.cumsum(Ap).astype(np.int32)
Ai = rows_.astype(np.int32)
for key in list(out.keys()):
    prec, refine, eps = key.split("_")
    eps = float(eps[3:])
    xo, ito, sto = orc.batch_qp_solve(N, Pd, q, Ap, Ai, Ax, l, u, eps_abs=eps, eps_rel=eps)
    x = np.array(out[key].pop("x"))
    it, st = np.array(out[key]["iters"]), np.array(out[key]["status"])
    ok = sto > 0
    d = np.abs(x[ok] - xo[ok])
    res = dict(max_diff=float(d.max()), max_diff_states=float(d[:, :93].max()), max_diff_u0=float(d[:, 93:95].max()),
               same_iters=int((it == ito).sum()), same_status=int((st == sto).sum()), B=int(B0),
               mean_iters=float(it.mean()), oracle_mean_iters=float(ito.mean()),
               nan_match=bool(np.all(np.isnan(x[~ok])) if (~ok).any() else True))
    out[key] = res
    print(key, res, flush=True)

# timing: replicate the 48 QPs to a large batch
for prec, refine in ((0, 1), (0, 0), (1, 0)):
    eng = mpc_b200.Engine(precision=prec)
    B = 148 * 8 * 16 if prec == 0 else 148 * 8 * 4
    rep = (B + B0 - 1) // B0
    big = [np.tile(a, (rep, 1))[:B] for a in (Pd, q, Ax, l, u)]
    tb = [tt(a) for a in big]
    x = torch.zeros((B, n), dtype=torch.float64, device=dev)
    it = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    for _ in range(2):
        eng.solve_qp(*tb, x, it, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        eng.solve_qp(*tb, x, it, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    iters = float(it.sum().item())
    res = dict(B=B, ms=ms, solves_per_s=B / ms * 1e3, admm_iters_per_s=iters / ms * 1e3, mean_iters=iters / B)
    out["timing_prec%d_ref%d" % (prec, refine)] = res
    print("timing", prec, refine, res, flush=True)
    eng.close()

os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(REPO, "gpurun_out", "gpu_first.json"), "w"), indent=1)
