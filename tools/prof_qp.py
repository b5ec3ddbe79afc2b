"""tools/prof_qp.py [precision] [refine] [B] -- one short solve_qp run for ncu (development aid)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import mpc_b200
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
refine = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 8 * 2
G = np.load(os.path.join(REPO, "tests", "golden", "teacher_forced.npz"))
ok = G["status"] == 1
dev = torch.device("cuda:0")
arrs = [G["qp_" + k][ok] for k in ("Pd", "q", "Ax", "l", "u")]
rep = (B + arrs[0].shape[0] - 1) // arrs[0].shape[0]
tb = [torch.tensor(np.tile(a, (rep, 1))[:B], dtype=torch.float64, device=dev) for a in arrs]
eng = mpc_b200.Engine(precision=prec)
x = torch.zeros((B, 153), dtype=torch.float64, device=dev)
it = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(2):
    eng.solve_qp(*tb, x, it, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.solve_qp(*tb, x, it, st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("prec", prec, "refine", refine, "B", B, "ms %.3f" % ms, "iters/s %.3e" % (it.sum().item() / ms * 1e3), "mean iters", it.float().mean().item())
