"""tools/hbm_kernels.py -- the two HBM-bound kernels of the path at a batch that fills the machine: K4a localize_t2s
(get_current_waypoint + t2s) and K4b rollout (BicycleModel.drive), fp64 SoA, through the C ABI on torch tensors.
Algorithmic bytes per instance: rollout reads x, y, psi, s, e_y, e_psi, wp_id, v, delta, flags and writes
x, y, psi, s = 108 B; localise reads x, y, psi, s (+ flags) and writes wp_id, e_y, e_psi = 56 B (table gathers are
cache-resident).  Prints one JSON line per batch size with the achieved GB/s against MEASURED_PEAKS.json's HBM figure."""
import json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch, bench, mpc_b200
from mpc_b200 import _lib

T, grid = bench.load_track()
peak = 6549.1
try:
    peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda:0")
eng = mpc_b200.Engine(precision=0)
eng.set_path(_lib.path_table(T["wp_x"], T["wp_y"], T["wp_psi"], T["wp_kappa"], T["wp_vref"]), np.cumsum(T["segment_lengths"]), T["border"], True)
eng.set_base_grid(grid, T["origin"], float(T["resolution"]))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for B in (4096, 65536, 1 << 20, 1 << 22):
    st = torch.tensor(bench.scenario_states(T, min(B, 65536), 0, min(B, 65536)), device=dev)
    st = st.repeat(1, B // st.shape[1]).contiguous()
    wp = torch.zeros(B, dtype=torch.int32, device=dev)
    sp = torch.zeros((2, B), dtype=torch.float64, device=dev)
    fl = torch.zeros(B, dtype=torch.int32, device=dev)
    u = torch.full((B, 2), 0.3, dtype=torch.float64, device=dev)
    res = {}
    for name, fn, nbytes in (("localize_t2s", lambda: eng.localize_t2s(st, wp, sp, fl), 56), ("rollout", lambda: eng.rollout(st, sp, wp, u, fl), 108)):
        for _ in range(3):
            fn()
        ts = []
        for k in range(10):
            flush.fill_(float(k))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        res[name] = dict(ms=ms, gbs=nbytes * B / (ms * 1e-3) / 1e9, frac=nbytes * B / (ms * 1e-3) / 1e9 / peak, bytes_per_instance=nbytes)
    print(json.dumps(dict(B=B, hbm_peak_gbs=peak, **res)), flush=True)
eng.close()
