"""CPU tier: the executable model of the ADMM kernel against the oracle, host-side logic of the
package, and the C-ABI export check (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import REPO, fixed_pattern, load_golden


def test_abi_exports_every_declared_symbol():
    """include/mpc_b200.h is the contract: every function it declares must be exported."""
    import mpc_b200
    hdr = open(os.path.join(REPO, "include", "mpc_b200.h")).read()
    declared = set(re.findall(r"\b(mpc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mpc_config", "mpc_engine"}
    lib = ctypes.CDLL(mpc_b200.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(mpc_b200.EXPORTS) <= declared
    assert lib.mpc_abi_version() == 1


def test_config_defaults_are_the_references():
    """mpc_config_default = src/simulation.py:100-111 + OSQP 0.6 defaults"""
    import mpc_b200
    c = mpc_b200.default_config()
    assert c.N == 30 and list(c.Q) == [1.0, 0.0, 0.0] and list(c.R) == [0.5, 0.0] and list(c.QN) == [1.0, 0.0, 0.0]
    assert c.umax[1] == np.tan(0.66) / 0.12 and c.umin[1] == -np.tan(0.66) / 0.12
    assert (c.rho, c.sigma, c.alpha, c.eps_abs, c.eps_rel) == (0.1, 1e-6, 1.6, 1e-3, 1e-3)
    assert (c.max_iter, c.scaling, c.check_termination) == (4000, 10, 25)
    assert np.isinf(c.xmin[0]) and np.isinf(c.xmax[2])


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    import mpc_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mpc_b200.MpcError):
        mpc_b200.Engine()
    lib = mpc_b200.load()
    h = ctypes.c_void_p()
    cfg = mpc_b200.default_config()
    assert lib.mpc_engine_create(ctypes.byref(cfg), ctypes.byref(h)) == -2  # MPC_E_CUDA
    assert b"no CPU fallback" in lib.mpc_last_error()


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: no product source may import, include or dlopen it."""
    pkg = os.path.join(REPO, "multi-purpose-mpc_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            for line in open(os.path.join(root, f)).read().splitlines():
                code = line.split("#")[0].split("//")[0]
                if re.search(r"\b(import|from|include|CDLL)\b", code):
                    assert "oracle" not in code and "liborc" not in code, (f, line)


def test_path_construction_matches_reference(track):
    """ReferencePath._construct_path / _compute_length (host numpy) are bit-identical to the reference's."""
    from mpc_b200.reference_path import ReferencePath
    rp = ReferencePath.__new__(ReferencePath)
    rp.eps, rp.resolution, rp.smoothing_distance, rp.circular = 1e-12, 0.05, 5, True
    wps = rp._construct_path(list(track.corner_x), list(track.corner_y))
    rp.waypoints = wps
    assert len(wps) == track.n_wp
    assert np.array_equal([w.x for w in wps], track.wp_x) and np.array_equal([w.y for w in wps], track.wp_y)
    assert np.array_equal([w.psi for w in wps], track.wp_psi)
    assert np.array_equal([w.kappa for w in wps], track.wp_kappa)
    length, seg = rp._compute_length()
    assert np.array_equal(seg, track.segment_lengths) and length == track.length


def test_map_add_obstacles_matches_reference(track):
    from mpc_b200.map import Map, Obstacle
    m = Map.from_grid(track.grid.copy(), list(track.origin), track.res)
    m.add_obstacles([Obstacle(*o) for o in track.obstacles])
    assert np.array_equal(m.data, track.grid_obs)
    assert m.w2m(-0.8, -1.5)[0] == 39  # fp64 pitfall
    assert m.m2w(39, 100) == ((39 + 0.5) * 0.005 + -1.0, (100 + 0.5) * 0.005 + -2.0)


def test_map_binarisation_and_hole_filling():
    from mpc_b200.map import Map
    raw = np.full((40, 40), 255, np.uint8)
    raw[10:12, 10:12] = 0        # 4-pixel hole: filled (area_threshold=5)
    raw[20:23, 20:23] = 0        # 9-pixel obstacle: kept
    raw[30, 30] = 99             # below threshold_occupied=100 -> occupied, single pixel -> filled
    m = Map(raw, origin=(0.0, 0.0), resolution=0.1)
    assert m.data.dtype == np.int8
    assert m.data[10:12, 10:12].all() and not m.data[20:23, 20:23].any() and m.data[30, 30] == 1


def test_host_line_aa_matches_oracle(orc):
    from mpc_b200.line_aa import line_aa_cells
    G = load_golden("line_aa.npz")
    for e, a, b in list(zip(G["ends"], G["offsets"][:-1], G["offsets"][1:]))[::5]:
        cells = np.array(list(line_aa_cells(*[int(v) for v in e])))
        assert np.array_equal(cells, G["cells"][a:b])


@pytest.mark.parametrize("dtype,eps,tol", [(np.float64, 1e-3, 1e-8), (np.float64, 1e-5, 1e-4), (np.float32, 1e-3, 1e-3)])  # fixture rows 7, 23 are infeasible QPs
def test_kernel_model_matches_oracle(orc, dtype, eps, tol):
    """tools/admm_pcr_model.py (stage layout + input elimination + PCR, the blueprint of
    csrc/admm.cuh) against oracle/osqp_oracle.c on QPs the reference assembled."""
    from tools import admm_pcr_model as M
    TF = load_golden("teacher_forced.npz")
    Ap, Ai = fixed_pattern(30)
    ks = [0, 7, 11, 23, 30]
    xo, ito, sto = orc.batch_qp_solve(30, TF["qp_Pd"][ks], TF["qp_q"][ks], Ap, Ai, TF["qp_Ax"][ks], TF["qp_l"][ks],
                                      TF["qp_u"][ks], eps_abs=eps, eps_rel=eps)
    for j, k in enumerate(ks):
        # fp64: the direct form; fp32: the increment form the CUDA kernel implements (no refinement needed)
        f = M.admm if dtype == np.float64 else M.admm_delta
        g = f(30, TF["qp_Pd"][k], TF["qp_q"][k], TF["qp_Ax"][k], TF["qp_l"][k], TF["qp_u"][k], dtype=dtype, eps_abs=eps,
              eps_rel=eps)
        assert g["status"] == sto[j] and g["iter"] == ito[j]
        if sto[j] == 1 and g["status"] == 1:
            assert np.abs(g["x"] - xo[j]).max() < tol


@pytest.mark.parametrize("dtype,eps,tol", [(np.float64, 1e-5, 1e-4), (np.float32, 1e-3, 1.5e-3)])
def test_paired_vform_model_matches_oracle(orc, dtype, eps, tol):
    """tools/admm_pcr_model.py::admm_vform with the paired factorisation (two stages per lane: in-lane cyclic reduction
    + PCR over the odd stages) -- the blueprint of csrc/admm_pair.cuh -- against oracle/osqp_oracle.c: same status and
    iteration count (incl. the infeasible rows 7 and 23), solution within the fp tolerance."""
    from tools import admm_pcr_model as M
    TF = load_golden("teacher_forced.npz")
    Ap, Ai = fixed_pattern(30)
    ks = [0, 7, 11, 23, 30]
    xo, ito, sto = orc.batch_qp_solve(30, TF["qp_Pd"][ks], TF["qp_q"][ks], Ap, Ai, TF["qp_Ax"][ks], TF["qp_l"][ks],
                                      TF["qp_u"][ks], eps_abs=eps, eps_rel=eps)
    for j, k in enumerate(ks):
        g = M.admm_vform(30, TF["qp_Pd"][k], TF["qp_q"][k], TF["qp_Ax"][k], TF["qp_l"][k], TF["qp_u"][k], dtype=dtype,
                         eps_abs=eps, eps_rel=eps, paired=True)
        if sto[j] in (1, -3):
            assert g["status"] == sto[j] and g["iter"] == ito[j], (k, sto[j], ito[j], g["status"], g["iter"])
        if sto[j] == 1 and g["status"] == 1:
            assert np.abs(g["x"] - xo[j]).max() < tol
