"""Multi-GPU plumbing: scenarios are independent, so each rank owns a contiguous shard and there is no
data-path collective.  The only exchange is one all-reduce of the small statistics
vector at the end of a run (SUM) plus one for the maxima (MAX) -- NCCL on GPUs, gloo in the CPU tests."""
import numpy as np

SUM_KEYS = ("scenario_steps", "qp_solves", "admm_iters", "qp_fallbacks", "dead", "finished", "sum_abs_ey")
MAX_KEYS = ("max_abs_ey",)


def shard_range(B, rank, world_size):
    """Contiguous batch range [lo, hi) of `rank`: [g*B/G, (g+1)*B/G) with integer arithmetic."""
    lo = (B * rank) // world_size
    hi = (B * (rank + 1)) // world_size
    return lo, hi


def shard_sizes(B, world_size):
    return [shard_range(B, r, world_size)[1] - shard_range(B, r, world_size)[0] for r in range(world_size)]


def allreduce_stats(stats, device=None):
    """All-reduce a statistics dict (as returned by Engine.run_closed_loop) over the default process
    group.  Works without torch.distributed being initialised (world size 1)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    s = torch.tensor([float(stats.get(k, 0.0)) for k in SUM_KEYS], dtype=torch.float64, device=dev)
    m = torch.tensor([float(stats.get(k, 0.0)) for k in MAX_KEYS], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    out = dict(stats)
    out.update({k: float(v) for k, v in zip(SUM_KEYS, s.tolist())})
    out.update({k: float(v) for k, v in zip(MAX_KEYS, m.tolist())})
    return out


def max_over_ranks(value, device=None):
    """max of a python float over ranks (timing: the slowest rank defines the step time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_scenarios(track_n_wp, B, seed, kind="tracking", wp_xy_psi=None, max_start_wp=None):
    """Synthetic scenario generator of SURVEY.md section 8d (same on every rank: generate all B, then slice).
    kind = "tracking": C2 -- start waypoint U{0..n_wp-1}, e_y ~ U(-0.05, 0.05), e_psi ~ U(-0.1, 0.1).
    kind = "obstacles": C3 -- additionally K ~ U{4..12} discs per scenario at random waypoints with lateral
    offset U(-0.15, 0.15) m and radius U(0.04, 0.08) m, rejected within 0.3 m of the start pose.
    Returns dict(start_wp, e_y, e_psi[, obs (n,3), obs_off (B+1,)])."""
    rng = np.random.default_rng(seed)
    # benchmarks keep the start waypoints away from the finish line so that no car ends its lap mid-run
    hi_wp = track_n_wp if max_start_wp is None else int(max_start_wp)
    out = dict(start_wp=rng.integers(0, hi_wp, B), e_y=rng.uniform(-0.05, 0.05, B),
               e_psi=rng.uniform(-0.1, 0.1, B))
    if kind == "obstacles":
        assert wp_xy_psi is not None
        wx, wy, wpsi = wp_xy_psi
        K = rng.integers(4, 13, B)
        obs, off = [], [0]
        for b in range(B):
            sx, sy = wx[out["start_wp"][b]], wy[out["start_wp"][b]]
            got = 0
            while got < K[b]:
                w = int(rng.integers(0, track_n_wp))
                o = rng.uniform(-0.15, 0.15)
                cx, cy = wx[w] - o * np.sin(wpsi[w]), wy[w] + o * np.cos(wpsi[w])
                r = rng.uniform(0.04, 0.08)
                if (cx - sx) ** 2 + (cy - sy) ** 2 < 0.3 ** 2:
                    continue
                obs.append((cx, cy, r))
                got += 1
            off.append(len(obs))
        out["obs"] = np.array(obs, dtype=np.float64)
        out["obs_off"] = np.array(off, dtype=np.int32)
    return out
