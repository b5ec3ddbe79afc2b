"""CPU tier: the oracle (oracle/*.c) against the golden vectors and solver-independent certificates."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as hst
from scipy import sparse

from conftest import fixed_pattern, load_golden, sim_cfg, ulps


# ---------------------------------------------------------------- line_aa --------------------
def test_line_aa_golden(orc):
    G = load_golden("line_aa.npz")
    for e, a, b in zip(G["ends"], G["offsets"][:-1], G["offsets"][1:]):
        rr, cc, _ = orc.line_aa(*[int(v) for v in e])
        assert np.array_equal(np.stack([rr, cc], 1), G["cells"][a:b])


@settings(max_examples=300, deadline=None)
@given(hst.integers(-300, 300), hst.integers(-300, 300), hst.integers(-300, 300), hst.integers(-300, 300))
def test_line_aa_invariants(r0, c0, r1, c1):
    """Properties any correct anti-aliased Bresenham line must have (the restatement cannot be checked
    against scikit-image here): starts at p0, ends at p1, main chain 8-connected, every cell within
    ~1.5 px of the ideal segment, reflection symmetry of the cell SET in r."""
    from oracle import oracle as orc
    rr, cc, _ = orc.line_aa(r0, c0, r1, c1)
    assert (rr[0], cc[0]) == (r0, c0)
    assert (rr[-1], cc[-1]) == (r1, c1)
    assert len(rr) <= 3 * (abs(r1 - r0) + abs(c1 - c0)) + 1
    d = np.array([r1 - r0, c1 - c0], float)
    L = np.hypot(*d)
    if L > 0:
        p = np.stack([rr - r0, cc - c0], 1).astype(float)
        t = np.clip(p @ d / (L * L), 0, 1)
        dist = np.hypot(*(p - t[:, None] * d).T)
        assert dist.max() <= 1.5 + 1e-9
    steps = np.abs(np.diff(rr)) + np.abs(np.diff(cc))
    assert steps.max(initial=0) <= 2 * 2  # consecutive emissions are neighbours or neighbours-of-neighbours
    rr2, cc2, _ = orc.line_aa(-r0, c0, -r1, c1)
    assert set(zip((-rr2).tolist(), cc2.tolist())) == set(zip(rr.tolist(), cc.tolist()))


def test_line_aa_degenerate(orc):
    rr, cc, _ = orc.line_aa(5, 5, 5, 5)
    assert (rr.tolist(), cc.tolist()) == ([5], [5])
    rr, cc, _ = orc.line_aa(0, 0, 0, 4)  # pure column walk: main cells plus the aa side cells
    assert (rr[0], cc[0], rr[-1], cc[-1]) == (0, 0, 0, 4)


# ---------------------------------------------------------------- map helpers -----------------
def test_w2m_fp64_pitfalls(orc):
    """0.07/0.005 = 14.000000000000002 -> ceil 15 ; (-0.8+1)/0.005 = 39.99.. -> 39"""
    import ctypes as C
    out = (C.c_long * 2)()
    orc.lib().orc_w2m(-1.0, -2.0, 0.005, -0.8, -1.5, out)
    assert out[0] == 39
    assert int(np.ceil(0.07 / 0.005)) == 15


def test_add_obstacles_matches_reference_raster(orc, track):
    g = track.grid.copy()
    for cx, cy, r in track.obstacles:
        assert orc.add_obstacle(g, track.origin, track.res, cx, cy, r) == 0
    assert np.array_equal(g, track.grid_obs)


# ---------------------------------------------------------------- static widths ---------------
def test_compute_width_matches_reference(orc, track, orc_path):
    orc.set_pow_mode(True)  # libm pow, as CPython evaluates `**2` in rp.py:282
    st, ub, lb, border = orc.compute_width(track.grid, track.origin, track.res, orc_path, 0.23)
    assert st == 0
    assert np.array_equal(border, track.border)
    assert np.array_equal(ub, track.wp_ub) and np.array_equal(lb, track.wp_lb)
    orc.set_pow_mode(False)  # IEEE x*x: cells identical, widths within 1 ulp
    st, ub2, lb2, border2 = orc.compute_width(track.grid, track.origin, track.res, orc_path, 0.23)
    assert np.array_equal(border2, track.border)
    assert ulps(ub2, track.wp_ub).max() <= 1 and ulps(lb2, track.wp_lb).max() <= 1


# ---------------------------------------------------------------- dynamic constraints ---------
def test_update_path_constraints_matches_reference(orc, track, orc_path):
    R = load_golden("raycast_random.npz")
    W = track.grid.shape[1]
    grids = np.unpackbits(R["grid_bits"], axis=2)[:, :, :W].astype(np.int8)
    sm = 0.06 / np.sqrt(2)
    worst = 0
    for (s, w, ok), ub_g, lb_g, cells_g in zip(R["wp_id"], R["ub"], R["lb"], R["cells_sm"]):
        orc.set_pow_mode(True)
        st, ub, lb, cells = orc.update_path_constraints(grids[s], track.origin, track.res, orc_path, int(w) + 1, 30,
                                                        2 * sm, sm)
        assert (st == 0) == bool(ok)
        if ok:
            assert np.array_equal(ub, ub_g) and np.array_equal(lb, lb_g)
            assert np.array_equal(cells, cells_g)
            orc.set_pow_mode(False)
            _, ub2, lb2, _ = orc.update_path_constraints(grids[s], track.origin, track.res, orc_path, int(w) + 1, 30,
                                                         2 * sm, sm)
            worst = max(worst, int(ulps(ub2, ub_g).max()), int(ulps(lb2, lb_g).max()))
    assert worst <= 1
    orc.set_pow_mode(False)


def test_sector_statistics(orc, track, orc_path):
    sm = 0.06 / np.sqrt(2)
    st, ub, lb, cells, n_cells, n_sec = orc.update_path_constraints(track.grid_obs, track.origin, track.res, orc_path, 1,
                                                                    30, 2 * sm, sm, want_stats=True)
    assert st == 0 and 30 * 40 < n_cells < 30 * 400 and 0 < n_sec <= 2 * track.grid.shape[0]


# ---------------------------------------------------------------- model -----------------------
def test_localize_and_drive_match_reference_lap(orc, track, orc_path):
    C1 = load_golden("c1_lap.npz")
    import ctypes as C
    L = orc.lib()
    for k in range(0, C1["state"].shape[0], 7):
        x, y, psi, s = C1["state"][k]
        w = L.orc_get_current_waypoint(orc_path.length_cum.ctypes.data_as(orc.c_double_p), track.n_wp, float(s))
        assert w == C1["wp_id"][k]
        sp = np.zeros(3)
        L.orc_t2s(x, y, psi, track.wp_x[w], track.wp_y[w], track.wp_psi[w], orc_path.cos_psi[w], orc_path.sin_psi[w],
                  sp.ctypes.data_as(orc.c_double_p))
        assert np.array_equal(sp, C1["spatial"][k])
        st4 = C1["state"][k].copy()
        L.orc_drive(st4.ctypes.data_as(orc.c_double_p), sp[0], sp[1], track.wp_kappa[w], C1["u"][k][0], C1["u"][k][1],
                    0.12, 0.05)
        # libm cos/sin/tan here vs numpy's in the reference: allow 2 ulp
        assert ulps(st4, C1["state_after"][k]).max() <= 2


# ---------------------------------------------------------------- QP assembly + OSQP ----------
def test_assembly_matches_reference_qp(orc, track, orc_path):
    TF = load_golden("teacher_forced.npz")
    cfg = sim_cfg(orc)
    Ap, Ai = fixed_pattern(30)
    for k in range(TF["state"].shape[0]):
        orc.set_pow_mode(False)  # IEEE x*x for kappa**2, v**2: within 1 ulp of the reference's libm pow
        A0 = orc.mpc_assemble(orc_path, cfg, int(TF["wp_id"][k]), TF["spatial"][k], TF["control"][k], TF["ub"][k],
                              TF["lb"][k])[2]
        assert ulps(A0.data, TF["qp_Ax"][k]).max() <= 2  # (1 ulp in the square) x (rounding of the product)
        orc.set_pow_mode(True)   # libm pow, exactly what CPython evaluates in sbm.py:405,410
        Pd, q, A, l, u = orc.mpc_assemble(orc_path, cfg, int(TF["wp_id"][k]), TF["spatial"][k], TF["control"][k],
                                          TF["ub"][k], TF["lb"][k])
        assert np.array_equal(A.indptr, Ap) and np.array_equal(A.indices, Ai)
        assert np.array_equal(Pd, TF["qp_Pd"][k])
        # tan() of libm vs numpy may differ in the last place inside the v_max bound
        assert np.allclose(A.data, TF["qp_Ax"][k], rtol=0, atol=0)
        assert np.array_equal(l, TF["qp_l"][k])
        assert ulps(u, TF["qp_u"][k]).max() <= 2
        assert np.array_equal(q, TF["qp_q"][k])
    orc.set_pow_mode(False)


@pytest.mark.parametrize("eps", [1e-3, 1e-5])
def test_osqp_restatement_kkt_certificate(orc, eps):
    """Whatever solver produced it, a returned point must satisfy the QP's optimality conditions to the
    requested tolerance (fp64, computed from P, q, A, l, u alone)."""
    TF = load_golden("teacher_forced.npz")
    Ap, Ai = fixed_pattern(30)
    n, m = 153, 246
    checked = 0
    for k in range(0, TF["state"].shape[0], 3):
        A = sparse.csc_matrix((TF["qp_Ax"][k], Ai, Ap), shape=(m, n))
        P = sparse.diags(TF["qp_Pd"][k])
        r = orc.osqp_solve(P, TF["qp_q"][k], A, TF["qp_l"][k], TF["qp_u"][k], perm=orc.stage_perm(30), eps_abs=eps,
                           eps_rel=eps, max_iter=20000)
        if r["status"] != 1:
            assert r["status"] == -3 and np.isnan(r["x"]).all()
            continue
        pri, dua, comp = orc.kkt_residuals(P, TF["qp_q"][k], A, TF["qp_l"][k], TF["qp_u"][k], r["x"], r["y"])
        scale = 1 + max(np.abs(A @ r["x"]).max(), np.abs(TF["qp_q"][k]).max(), np.abs(A.T @ r["y"]).max())
        assert pri <= eps * scale and dua <= eps * scale
        checked += 1
    assert checked >= 8


def test_osqp_restatement_reproduces_golden(orc):
    """The golden x / iteration counts came from the same restatement driven by the reference's Python:
    the standalone batch entry must reproduce them bit-for-bit (guards the fixture against drift)."""
    TF = load_golden("teacher_forced.npz")
    Ap, Ai = fixed_pattern(30)
    x, it, st = orc.batch_qp_solve(30, TF["qp_Pd"], TF["qp_q"], Ap, Ai, TF["qp_Ax"], TF["qp_l"], TF["qp_u"])
    assert np.array_equal(it, TF["qp_iters"]) and np.array_equal(st, TF["qp_status"])
    ok = st == 1
    assert np.array_equal(x[ok], TF["qp_x"][ok])
    assert np.isnan(x[~ok]).all()


def test_null_direction_is_the_only_one(orc):
    """rank [P; Aeq] = n - 1 with null vector (u_{N-1}.kappa, x_N.e_psi) = (1, ds)."""
    TF = load_golden("teacher_forced.npz")
    Ap, Ai = fixed_pattern(30)
    A = sparse.csc_matrix((TF["qp_Ax"][0], Ai, Ap), shape=(246, 153)).toarray()
    M = np.vstack([np.diag(TF["qp_Pd"][0]), A[:93]])
    sv = np.linalg.svd(M, compute_uv=False)
    assert (sv < 1e-10).sum() == 1


def test_h1_projection_helper_matches_the_svd_null_vector(orc):
    """tests/conftest.py::h1_null_direction (read off the constraint values) against the SVD null vector of [P; Aeq]."""
    from conftest import h1_null_direction
    TF = load_golden("teacher_forced.npz")
    Ap, Ai = fixed_pattern(30)
    d = h1_null_direction(30, TF["qp_Pd"][:6], TF["qp_Ax"][:6])
    for b in range(6):
        A = sparse.csc_matrix((TF["qp_Ax"][b], Ai, Ap), shape=(246, 153)).toarray()
        _, sv, vt = np.linalg.svd(np.vstack([np.diag(TF["qp_Pd"][b]), A[:93]]))
        assert sv[-1] < 1e-10 and abs(abs(vt[-1] @ d[b]) - 1.0) < 1e-12


def test_osqp_final_checks_after_max_iter(orc):
    """osqp.c after the main loop: a NORMAL termination check when the last iteration was not a check iteration, then the
    approximate one (statuses 2 / 3 / 4).  max_iter = 60 (60 % 25 != 0): QPs converging between iteration 50 and 60 are
    `solved` (1) at 60 iterations, the rest `solved inaccurate` (2); eps 1e-5 with max_iter 160 / 310 reaches the exact
    certificate at the final check (-3 at 160) and the inaccurate one (3), both without a solution."""
    TF, C1 = load_golden("teacher_forced.npz"), load_golden("c1_lap.npz")
    Pd, q, Ax, l, u = (np.concatenate([TF["qp_" + k], C1["qp_" + k]]) for k in ("Pd", "q", "Ax", "l", "u"))
    Ap, Ai = fixed_pattern(30)
    x, it, st = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u, max_iter=60)
    assert ((st == 1) & (it == 60)).sum() >= 3 and ((st == 2) & (it == 60)).sum() >= 30 and ((st == 1) & (it <= 50)).any()
    x, it, st = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u, max_iter=160, eps_abs=1e-5, eps_rel=1e-5)
    assert ((st == -3) & (it == 160)).any()
    x, it, st = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u, max_iter=310, eps_abs=1e-5, eps_rel=1e-5)
    assert ((st == 3) & (it == 310)).any() and np.isnan(x[st == 3]).all() and np.isfinite(x[st == -2]).all()


def test_full_step_matches_reference_lap(orc, track, orc_path):
    """orc_mpc_step (the C restatement of get_control + drive) against the reference's own closed loop."""
    C1 = load_golden("c1_lap.npz")
    world = orc.World(orc_path, sim_cfg(orc), track.grid.shape, track.origin, track.res, 0.05)
    orc.set_pow_mode(True)
    try:
        for k in list(range(0, 40)) + list(range(40, C1["state"].shape[0], 5)):
            r = world.step(track.grid_obs, C1["state"][k], C1["control"][k], int(C1["infeas"][k]))
            assert r["wp_id"] == C1["wp_id"][k]
            assert np.array_equal(r["ub"], C1["ub"][k]) and np.array_equal(r["lb"], C1["lb"][k])
            assert r["qp_status"] == C1["status"][k] and r["iters"] == C1["iters"][k]
            assert np.allclose(r["u"], C1["u"][k], rtol=0, atol=1e-9)
            assert np.allclose(r["state"], C1["state_after"][k], rtol=1e-12, atol=1e-12)
    finally:
        orc.set_pow_mode(False)


def test_osqp_restatement_against_the_real_solver_when_installed(orc):
    """Independent tier for the one third-party algorithm on the hot path: when a real `osqp` wheel is importable (it is not
    in the offline image, so this normally skips) the restatement must agree with it on the golden QPs -- same status, the
    primal solution within 1e-3 after the H1 projection at eps 1e-5 (the north-star's setting, where the solver's own
    adaptive-rho timing heuristic no longer matters), and iteration counts within one termination check at the defaults
    with adaptive_rho_interval pinned to 25 on both sides."""
    osqp = pytest.importorskip("osqp")
    from conftest import h1_split
    TF = load_golden("teacher_forced.npz")
    Ap, Ai = fixed_pattern(30)
    ks = [k for k in range(len(TF["qp_status"])) if TF["qp_status"][k] == 1][:12]
    for eps, tol_it in ((1e-3, 25), (1e-5, None)):
        xo, ito, sto = orc.batch_qp_solve(30, TF["qp_Pd"][ks], TF["qp_q"][ks], Ap, Ai, TF["qp_Ax"][ks], TF["qp_l"][ks],
                                          TF["qp_u"][ks], eps_abs=eps, eps_rel=eps)
        for j, k in enumerate(ks):
            P = sparse.diags(TF["qp_Pd"][k]).tocsc()
            A = sparse.csc_matrix((TF["qp_Ax"][k], Ai, Ap), shape=(246, 153))
            m = osqp.OSQP()
            m.setup(P=P, q=TF["qp_q"][k], A=A, l=TF["qp_l"][k], u=TF["qp_u"][k], verbose=False, eps_abs=eps, eps_rel=eps,
                    adaptive_rho_interval=25, polish=False)
            r = m.solve()
            assert r.info.status_val == sto[j]
            if tol_it is not None:
                assert abs(r.info.iter - ito[j]) <= tol_it
            rem, null = h1_split(30, TF["qp_Pd"][k][None], TF["qp_Ax"][k][None], (r.x - xo[j])[None])
            if eps < 1e-4:
                assert np.abs(rem).max() <= 1e-3
