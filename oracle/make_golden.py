"""oracle/make_golden.py -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src, imported through oracle/ref_harness.py) in the build container.

Provenance classes recorded inside every file:
  "reference-code"        produced purely by the reference's numpy code (authoritative)
  "restated-dependency"   flowed through the oracle's restatement of skimage.draw.line_aa and/or
                          OSQP (authoritative only up to the correctness of that restatement)

Run:  python oracle/make_golden.py     (needs /root/reference; not runnable on the GPU box)
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import oracle as orc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.environ.get("MPC_GOLDEN_OUT") or os.path.join(REPO, "tests", "golden")   # MPC_GOLDEN_OUT: write elsewhere (live test)


def fixed_pattern(N):
    """(rows, cols) of the structural CSC pattern of the reference's A (MPC.py:128-135), zeros kept."""
    nx, nu = 3, 2
    neq = nx * (N + 1)
    rows, cols = [], []
    for col in range(neq):
        k, j = divmod(col, nx)
        rows.append(col); cols.append(col)
        if k < N:
            rr = {0: [0, 1, 2], 1: [0, 1], 2: [2]}[j]
            rows += [nx * (k + 1) + r for r in rr]; cols += [col] * len(rr)
        rows.append(neq + col); cols.append(col)
    for col in range(neq, neq + nu * N):
        k, j = divmod(col - neq, nu)
        rows.append(nx * (k + 1) + (2 if j == 0 else 1)); cols.append(col)
        rows.append(neq + col); cols.append(col)
    return np.array(rows), np.array(cols)


def path_arrays(rp):
    w = rp.waypoints
    return dict(
        wp_x=np.array([p.x for p in w]), wp_y=np.array([p.y for p in w]), wp_psi=np.array([p.psi for p in w]),
        wp_kappa=np.array([p.kappa for p in w]),
        wp_vref=np.array([np.nan if p.v_ref is None else p.v_ref for p in w]),
        wp_ub=np.array([p.ub for p in w]), wp_lb=np.array([p.lb for p in w]),
        border=np.array([[p.static_border_cells[0][0], p.static_border_cells[0][1],
                          p.static_border_cells[1][0], p.static_border_cells[1][1]] for p in w]),
        segment_lengths=np.array(rp.segment_lengths), length=np.float64(rp.length))


def qp_record(log_entry, N, rows, cols):
    A = log_entry["A"].toarray()
    chk = A.copy(); chk[rows, cols] = 0
    assert np.abs(chk).max() == 0, "reference A has entries outside the structural pattern"
    return dict(Pd=log_entry["P"].diagonal(), q=log_entry["q"], Ax=A[rows, cols], l=log_entry["l"],
                u=log_entry["u"], x=log_entry["x"], status=log_entry["status"], iters=log_entry["iter"])


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = rh.load()
    prov = ns.provenance
    print("provenance:", prov)
    orc.set_pow_mode(True)  # irrelevant for the reference's own code; documents the libm mode
    N = rh.SIM["N"]
    rows, cols = fixed_pattern(N)

    # ---------------- 1. track: map grid + path tables -------------------------------------
    mp0 = ns.Map(file_path=rh.SIM["map_file"], origin=rh.SIM["origin"], resolution=rh.SIM["resolution"])
    grid0 = mp0.data.copy()
    mp, rp, car, mpc = rh.build_sim(ns, use_obstacles=True)
    pa = path_arrays(rp)
    np.savez_compressed(
        os.path.join(OUT, "sim_track.npz"), grid_bits=np.packbits(grid0.astype(np.uint8), axis=1),
        grid_shape=np.array(grid0.shape), origin=np.array(rh.SIM["origin"], float),
        resolution=np.float64(rh.SIM["resolution"]), corner_x=np.array(rh.SIM["wp_x"]),
        corner_y=np.array(rh.SIM["wp_y"]), obstacles=np.array(rh.SIM["obstacles"]),
        grid_obstacles_bits=np.packbits(mp.data.astype(np.uint8), axis=1),
        provenance=np.array(["grid, waypoints x/y/psi/kappa, segment_lengths, obstacle raster: reference-code; "
                             "wp_ub/wp_lb/border (line_aa) and wp_vref (OSQP): restated-dependency; " + str(prov)]),
        **pa)
    print("sim_track.npz: n_wp", len(pa["wp_x"]), "length", float(pa["length"]))

    # ---------------- 2. C1: the reference's default closed loop ---------------------------
    ns.osqp.LOG_ENABLED = True
    ns.osqp.LOG.clear()
    rec = dict(state=[], control=[], wp_id=[], spatial=[], ub=[], lb=[], u=[], state_after=[], status=[], iters=[],
               infeas=[])
    qps = []
    k = 0
    while car.s < rp.length and k < 2000:
        rec["state"].append([car.temporal_state.x, car.temporal_state.y, car.temporal_state.psi, car.s])
        rec["control"].append(mpc.current_control.copy())
        rec["infeas"].append(mpc.infeasibility_counter)
        u = mpc.get_control()
        # ub/lb as the reference computed them inside _init_problem: recover from l,u of the QP
        L = ns.osqp.LOG[-1]
        neq = 3 * (N + 1)
        rec["lb"].append(L["l"][neq + 3::3][:N].copy())
        rec["ub"].append(L["u"][neq + 3::3][:N].copy())
        rec["wp_id"].append(car.wp_id)
        rec["spatial"].append(list(car.spatial_state[:]))
        car.drive(u)
        rec["u"].append(np.array(u, float))
        rec["state_after"].append([car.temporal_state.x, car.temporal_state.y, car.temporal_state.psi, car.s])
        rec["status"].append(L["status"]); rec["iters"].append(L["iter"])
        if k % 8 == 0 or L["status"] != 1:
            q = qp_record(L, N, rows, cols); q["step"] = k
            qps.append(q)
        k += 1
    out = {kk: np.array(v) for kk, v in rec.items()}
    for key in ("Pd", "q", "Ax", "l", "u", "x", "status", "iters", "step"):
        out["qp_" + key] = np.array([q[key] for q in qps])
    out["provenance"] = np.array(["closed loop of src/simulation.py:134-140 run by the reference's own code; "
                                  "ub/lb via restated line_aa, x/u/status/iters via restated OSQP at the "
                                  "reference's default settings (adaptive_rho_interval fixed to 25): "
                                  "restated-dependency; " + str(prov)])
    np.savez_compressed(os.path.join(OUT, "c1_lap.npz"), **out)
    print("c1_lap.npz: steps", k, "qps", len(qps), "statuses", sorted(set(rec["status"])))

    # ---------------- 3. teacher-forced single steps at random states (C2/C3-like) ----------
    rng = np.random.default_rng(20240517)
    ns.osqp.LOG.clear()
    tf = dict(state=[], control=[], wp_id=[], spatial=[], ub=[], lb=[], u=[], state_after=[], status=[], iters=[],
              control_after=[])
    tqps = []
    length_cum = np.cumsum(rp.segment_lengths)
    for i in range(48):
        w = int(rng.integers(0, rp.n_waypoints - 1))
        e_y, e_psi = rng.uniform(-0.05, 0.05), rng.uniform(-0.1, 0.1)
        wp = rp.waypoints[w]
        ts = car.s2t(wp, np.array([e_y, e_psi, 0.0]))
        car.temporal_state = ts
        car.s = float(length_cum[w]) + 1e-3 * rng.uniform(0, 1)
        cc = np.zeros(2 * N)
        if i % 3:
            cc[0::2] = rng.uniform(0.3, 1.0, N)
            cc[1::2] = rng.uniform(-0.4, 0.4, N)
        mpc.current_control = cc.copy()
        mpc.infeasibility_counter = 0
        tf["state"].append([ts.x, ts.y, ts.psi, car.s]); tf["control"].append(cc)
        u = mpc.get_control()
        L = ns.osqp.LOG[-1]
        neq = 3 * (N + 1)
        tf["lb"].append(L["l"][neq + 3::3][:N].copy()); tf["ub"].append(L["u"][neq + 3::3][:N].copy())
        tf["wp_id"].append(car.wp_id); tf["spatial"].append(list(car.spatial_state[:]))
        car.drive(u)
        tf["u"].append(np.array(u, float))
        tf["state_after"].append([car.temporal_state.x, car.temporal_state.y, car.temporal_state.psi, car.s])
        tf["status"].append(L["status"]); tf["iters"].append(L["iter"])
        tf["control_after"].append(np.array(mpc.current_control, float))
        q = qp_record(L, N, rows, cols); q["step"] = i
        tqps.append(q)
    out = {kk: np.array(v) for kk, v in tf.items()}
    for key in ("Pd", "q", "Ax", "l", "u", "x", "status", "iters", "step"):
        out["qp_" + key] = np.array([q[key] for q in tqps])
    out["provenance"] = np.array(["single get_control()+drive() calls of the reference from injected states "
                                  "(teacher forcing, SURVEY H1-ii), obstacles of simulation.py:40-48: "
                                  "restated-dependency; " + str(prov)])
    np.savez_compressed(os.path.join(OUT, "teacher_forced.npz"), **out)
    print("teacher_forced.npz:", len(tqps), "statuses", sorted(set(tf["status"])))

    # ---------------- 4. raycast with randomised obstacle sets ----------------------------
    ray = dict(obs=[], obs_off=[0], wp_id=[], ub=[], lb=[], cells_sm=[], grid_bits=[])
    sm = car.safety_margin
    for sc in range(12):
        mpx = ns.Map(file_path=rh.SIM["map_file"], origin=rh.SIM["origin"], resolution=rh.SIM["resolution"])
        rp.map = mpx  # static border cells were computed on the obstacle-free map (simulation.py order)
        K = int(rng.integers(4, 13))
        obs = []
        while len(obs) < K:
            w = int(rng.integers(0, rp.n_waypoints))
            wp = rp.waypoints[w]
            off = rng.uniform(-0.15, 0.15)
            cx, cy = wp.x - off * np.sin(wp.psi), wp.y + off * np.cos(wp.psi)
            r = rng.uniform(0.04, 0.08)
            obs.append((cx, cy, r))
        mpx.add_obstacles([ns.Obstacle(*o) for o in obs])
        ray["obs"] += obs; ray["obs_off"].append(len(ray["obs"]))
        ray["grid_bits"].append(np.packbits(mpx.data.astype(np.uint8), axis=1))
        for w in rng.integers(0, rp.n_waypoints, 6):
            try:
                ub, lb, cells = rp.update_path_constraints(int(w) + 1, N, 2 * sm, sm)
                cells = np.array([[c[0][0], c[0][1], c[1][0], c[1][1]] for c in cells])
                ok = 1
            except ValueError:  # max([]) at the first waypoint (rp.py:547)
                ub, lb, cells, ok = np.full(N, np.nan), np.full(N, np.nan), np.full((N, 4), np.nan), 0
            ray["wp_id"].append([sc, int(w), ok]); ray["ub"].append(ub); ray["lb"].append(lb)
            ray["cells_sm"].append(cells)
    rp.map = mp
    out = {kk: np.array(v) for kk, v in ray.items()}
    out["provenance"] = np.array(["ReferencePath.update_path_constraints (rp.py:522-648) run by the reference's own "
                                  "code on Map.add_obstacles rasters (reference-code) through restated line_aa: "
                                  "restated-dependency; widths use CPython/numpy `**2` = libm pow; " + str(prov)])
    np.savez_compressed(os.path.join(OUT, "raycast_random.npz"), **out)
    print("raycast_random.npz: cases", len(ray["wp_id"]), "no-segment cases", sum(1 for w in ray["wp_id"] if not w[2]))

    # ---------------- 5. line_aa known-answer vectors (restated-dependency) ---------------
    ends, offs, cells = [], [0], []
    for _ in range(200):
        r0, c0 = rng.integers(0, 500, 2)
        r1, c1 = r0 + rng.integers(-120, 121), c0 + rng.integers(-120, 121)
        rr, cc, _v = orc.line_aa(int(r0), int(c0), int(r1), int(c1))
        ends.append([r0, c0, r1, c1]); cells += list(zip(rr.tolist(), cc.tolist())); offs.append(len(cells))
    for e in ([5, 5, 5, 5], [0, 0, 10, 0], [0, 0, 0, 10], [10, 10, 0, 0], [3, 7, 4, 7], [0, 0, 7, 7], [0, 9, 9, 0]):
        rr, cc, _v = orc.line_aa(*e)
        ends.append(e); cells += list(zip(rr.tolist(), cc.tolist())); offs.append(len(cells))
    np.savez_compressed(os.path.join(OUT, "line_aa.npz"), ends=np.array(ends), offsets=np.array(offs),
                        cells=np.array(cells, dtype=np.int32),
                        provenance=np.array(["oracle restatement of skimage.draw.line_aa (parity unpinned); "
                                             "checked by invariants in tests/test_oracle_cpu.py"]))
    print("line_aa.npz:", len(ends), "lines")


if __name__ == "__main__":
    main()
