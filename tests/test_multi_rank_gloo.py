"""CPU tier: the N > 1 host logic (sharding + the final statistics all-reduce) with world_size 2 on gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, REPO)
    import mpc_b200  # noqa: F401
    from mpc_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(B, rank, world)
    sc = D.make_scenarios(200, B, seed=2)
    local = dict(scenario_steps=float(hi - lo) * 3, qp_solves=float(hi - lo) * 3,
                 admm_iters=float(np.sum(sc["start_wp"][lo:hi])), qp_fallbacks=float(rank), dead=0.0, finished=0.0,
                 sum_abs_ey=float(np.abs(sc["e_y"][lo:hi]).sum()), max_abs_ey=float(np.abs(sc["e_y"][lo:hi]).max()))
    tot = D.allreduce_stats(local)
    t = D.max_over_ranks(1.0 + rank)
    q.put((rank, lo, hi, tot, t))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    from mpc_b200 import distributed as D
    for B in (1, 7, 4096, 65536, 262144 + 3):
        for G in (1, 2, 4, 8):
            r = [D.shard_range(B, g, G) for g in range(G)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = D.shard_sizes(B, G)
            assert max(sizes) - min(sizes) <= 1


def test_scenario_generator_is_rank_independent():
    from mpc_b200 import distributed as D
    a, b = D.make_scenarios(200, 64, seed=3), D.make_scenarios(200, 64, seed=3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    wx = np.linspace(0, 1, 200); wy = np.zeros(200); wpsi = np.zeros(200)
    o = D.make_scenarios(200, 16, seed=3, kind="obstacles", wp_xy_psi=(wx, wy, wpsi))
    K = np.diff(o["obs_off"])
    assert K.min() >= 4 and K.max() <= 12 and o["obs"].shape == (o["obs_off"][-1], 3)
    assert (o["obs"][:, 2] >= 0.04).all() and (o["obs"][:, 2] <= 0.08).all()


def test_world_size_2_allreduce_of_statistics():
    world, B = 2, 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from mpc_b200 import distributed as D
    sc = D.make_scenarios(200, B, seed=2)
    (r0, lo0, hi0, t0, m0), (r1, lo1, hi1, t1, m1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 50, 50, 101)
    for tot in (t0, t1):
        assert tot["scenario_steps"] == 3 * B and tot["qp_fallbacks"] == 1.0
        assert tot["admm_iters"] == float(sc["start_wp"].sum())
        assert np.isclose(tot["sum_abs_ey"], np.abs(sc["e_y"]).sum())
        assert tot["max_abs_ey"] == np.abs(sc["e_y"]).max()
    assert m0 == m1 == 2.0
