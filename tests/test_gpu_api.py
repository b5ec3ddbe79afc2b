"""GPU tier: the drop-in Python surface (Map, ReferencePath, BicycleModel, MPC) used the way
src/simulation.py uses the reference's, on the sim track rebuilt from the golden fixture."""
import numpy as np
import pytest
from scipy import sparse

from conftest import load_golden, ulps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim(track, tmp_path_factory):
    from PIL import Image
    import mpc_b200  # noqa: F401
    from mpc_b200.map import Map, Obstacle
    from mpc_b200.reference_path import ReferencePath
    from mpc_b200.spatial_bicycle_models import BicycleModel
    from mpc_b200.MPC import MPC
    # an image whose channel 0 binarises (>= 100) to the reference's sim_map grid
    img = np.repeat((track.grid.astype(np.uint8) * 255)[:, :, None], 3, axis=2)
    p = tmp_path_factory.mktemp("maps") / "sim_map.png"
    Image.fromarray(img).save(p)
    mp = Map(file_path=str(p), origin=[-1, -2], resolution=0.005)
    rp = ReferencePath(mp, list(track.corner_x), list(track.corner_y), 0.05, smoothing_distance=5, max_width=0.23,
                       circular=True)
    mp.add_obstacles([Obstacle(cx=o[0], cy=o[1], radius=o[2]) for o in track.obstacles])
    car = BicycleModel(length=0.12, width=0.06, reference_path=rp, Ts=0.05)
    N = 30
    Q, R, QN = sparse.diags([1.0, 0.0, 0.0]), sparse.diags([0.5, 0.0]), sparse.diags([1.0, 0.0, 0.0])
    ic = {'umin': np.array([0.0, -np.tan(0.66) / car.length]), 'umax': np.array([1.0, np.tan(0.66) / car.length])}
    sc = {'xmin': np.array([-np.inf] * 3), 'xmax': np.array([np.inf] * 3)}
    mpc = MPC(car, N, Q, R, QN, sc, ic, 4.0, precision=1)
    car.reference_path.compute_speed_profile({'a_min': -0.1, 'a_max': 0.5, 'v_min': 0.0, 'v_max': 1.0, 'ay_max': 4.0})
    return mp, rp, car, mpc


def test_setup_matches_reference(sim, track):
    mp, rp, car, mpc = sim
    assert np.array_equal(mp.data, track.grid_obs)
    assert rp.n_waypoints == 200 and rp.length == track.length
    assert np.array_equal([w.x for w in rp.waypoints], track.wp_x)
    assert np.array_equal([w.kappa for w in rp.waypoints], track.wp_kappa)
    border = np.array([[w.static_border_cells[0][0], w.static_border_cells[0][1], w.static_border_cells[1][0],
                        w.static_border_cells[1][1]] for w in rp.waypoints])
    assert np.array_equal(border, track.border)
    assert ulps([w.ub for w in rp.waypoints], track.wp_ub).max() <= 1
    assert np.abs(np.array([w.v_ref for w in rp.waypoints]) - track.wp_vref).max() <= 1e-9
    assert car.safety_margin == 0.06 / np.sqrt(2) and car.n_states == 3


def test_get_control_and_drive_follow_the_reference_lap(sim):
    """simulation.py:134-140 with the reference's class names; compared step by step with the reference run."""
    mp, rp, car, mpc = sim
    C1 = load_golden("c1_lap.npz")
    k = 0
    while car.s < rp.length and k < 60:
        u = mpc.get_control()
        car.drive(u)
        assert isinstance(u, np.ndarray) and u.shape == (2,)
        assert mpc.last_iters == C1["iters"][k]
        assert np.abs(u - C1["u"][k]).max() <= 1e-6
        got = np.array([car.temporal_state.x, car.temporal_state.y, car.temporal_state.psi, car.s])
        assert np.abs(got - C1["state_after"][k]).max() <= 1e-6
        assert car.wp_id == C1["wp_id"][k]
        k += 1
    assert mpc.current_control.shape == (60,) and len(mpc.current_prediction[0]) == 28


def test_update_path_constraints_api(sim, track):
    mp, rp, car, mpc = sim
    R = load_golden("c1_lap.npz")
    ub, lb, cells = rp.update_path_constraints(int(R["wp_id"][10]) + 1, 30, 2 * car.safety_margin, car.safety_margin)
    assert ulps(ub, R["ub"][10]).max() <= 1 and ulps(lb, R["lb"][10]).max() <= 1
    assert len(cells) == 30 and len(cells[0]) == 2


def test_batched_api_matches_single_car(sim, track):
    import torch
    from mpc_b200.spatial_bicycle_models import BatchedBicycleModel
    from mpc_b200.MPC import BatchedMPC
    mp, rp, car, mpc = sim
    B = 64
    cars = BatchedBicycleModel(rp, 0.12, 0.06, 0.05, B)
    ic = {'umin': np.array([0.0, -np.tan(0.66) / 0.12]), 'umax': np.array([1.0, np.tan(0.66) / 0.12])}
    sc = {'xmin': np.array([-np.inf] * 3), 'xmax': np.array([np.inf] * 3)}
    bm = BatchedMPC(cars, 30, sparse.diags([1.0, 0.0, 0.0]), sparse.diags([0.5, 0.0]), sparse.diags([1.0, 0.0, 0.0]), sc,
                    ic, 4.0, precision=1)
    C1 = load_golden("c1_lap.npz")
    for k in range(3):
        u = bm.get_control()
        cars.drive(u)
        assert tuple(u.shape) == (B, 2)
        assert torch.equal(u[0], u[B - 1])                                     # identical scenarios, identical answers
        assert np.abs(u[0].cpu().numpy() - C1["u"][k]).max() <= 1e-6
    # MPC.update_prediction, batched: same numbers as the single-car host method (MPC.py:224-248, sbm.py:155-181)
    xy = bm.update_prediction().cpu().numpy()
    assert xy.shape == (B, 28, 2)
    xs = bm.solution[0].cpu().numpy()
    wp0 = int(bm._b.wp_id[0].item())
    for n in range(2, 30):
        wp = rp.get_waypoint(wp0 + n)
        e_y = xs[3 * n]
        assert xy[0, n - 2, 0] == wp.x - e_y * np.sin(wp.psi) and xy[0, n - 2, 1] == wp.y + e_y * np.cos(wp.psi)
    assert np.array_equal(xy[0], xy[B - 1])
    stats = bm.run_closed_loop(10)
    assert stats["scenario_steps"] == 10 * B and stats["dead"] == 0
    assert np.abs(cars.s.cpu().numpy() - C1["state_after"][12][3]).max() <= 1e-6
    # a second call continues the fleet: scenarios that died or finished in an earlier call stay out of the loop (the
    # reference's exit(1) at MPC.py:218-220 is final) and the engine-owned buffers are not re-allocated
    bm.flags[3] = 2                                                            # MPC_ST_DEAD
    pose3 = cars.state[:, 3].clone() if hasattr(cars, "state") else None
    s_before = cars.s.cpu().numpy().copy()
    stats2 = bm.run_closed_loop(5)
    s_after = cars.s.cpu().numpy()
    assert s_after[3] == s_before[3] and (s_after[np.arange(B) != 3] > s_before[np.arange(B) != 3]).all()
    assert int(bm.flags[3].item()) & 2
    assert stats2["scenario_steps"] == 5 * (B - 1)
    assert np.abs(s_after[0] - C1["state_after"][17][3]) <= 1e-6
