"""bench.py -- headline benchmark of the B200 batched MPC engine.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--precision 0|1]
  (N > 1: launched by torchrun, one rank per GPU; RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env)

Workload (BASELINE.json configs[1]): path tracking, 4096 independent cars per GPU on the sim track with
randomised start waypoints / offsets (seed 2), horizon N = 30, reference weights and OSQP defaults
(cold start, eps_abs = eps_rel = 1e-3).  A "step" is one closed-loop step of every car:
localise + t2s -> grid raycast -> LTV assembly + ADMM QP solve -> rollout (MPC.get_control + car.drive of
the reference, src/simulation.py:137-140), i.e. one QP solve per car.

Prints ONE JSON line (rank 0).  `value` = QP solves/s of the whole job with state resident in HBM;
`e2e` = the same through mpc_step_host (pinned host buffers, H2D + D2H inside the timed region);
`roofline` for the dominant kernel (K1+K2 assemble + ADMM), `cpu_baseline` = the CPU oracle port (the
reference's algorithm in C, all host cores) on a bounded sample of the same scenarios.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

N_HORIZON = 30
METRIC = "mpc_qp_solves_per_sec_N30"
UNIT = "QP solves/s"


def load_track():
    G = os.path.join(REPO, "tests", "golden")
    T = np.load(os.path.join(G, "sim_track.npz"))
    W = int(T["grid_shape"][1])
    grid = np.unpackbits(T["grid_bits"], axis=1)[:, :W].astype(np.int8)
    return T, grid


def scenario_states(T, B, lo, hi, seed=2, kind="tracking", return_obstacles=False):
    """C2 / C3 of SURVEY.md section 8d: all B scenarios are generated identically on every rank, then sliced."""
    from mpc_b200 import distributed as D
    # start waypoints U{0..119}: 80 waypoints (3.5 m = 70+ steps at 1 m/s) of headroom before s >= length would
    # end a car's lap (simulation.py:134), so every car is live in every timed step
    sc = D.make_scenarios(len(T["wp_x"]), B, seed=seed, max_start_wp=len(T["wp_x"]) - 80, kind=kind,
                          wp_xy_psi=(T["wp_x"], T["wp_y"], T["wp_psi"]))
    if return_obstacles:
        off = sc["obs_off"]
        return sc["obs"][off[lo]:off[hi]], (off[lo:hi + 1] - off[lo]).astype(np.int32)
    w = sc["start_wp"][lo:hi]
    e_y, e_psi = sc["e_y"][lo:hi], sc["e_psi"][lo:hi]
    lc = np.cumsum(T["segment_lengths"])
    x, y, psi = T["wp_x"][w], T["wp_y"][w], T["wp_psi"][w]
    return np.ascontiguousarray(np.stack([x - e_y * np.sin(psi), y + e_y * np.cos(psi), psi + e_psi, lc[w]]))


def flop_model(N, iters, n_factor=1):
    """F_iter = 340N + 264, F_factor = 400(N+1), F_check = 142N + 78 (every 25 iterations)."""
    f_iter, f_fac, f_chk = 340 * N + 264, 400 * (N + 1), 142 * N + 78
    return iters * f_iter + n_factor * f_fac + (iters // 25) * f_chk


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed region is tens of
    milliseconds, far shorter than one `nvidia-smi` process start, so the samples come from NVML in-process (pynvml,
    ~20 us per query, one every 2 ms); `nvidia-smi` is the fallback when pynvml cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nv = self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:  # NVML numbers the physical GPUs; CUDA's index may be remapped by CUDA_VISIBLE_DEVICES
                self.handle = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def sample(self):
        nv = self.nv
        if nv is None:
            return
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
            act = lambda bit: "Active" if (r & bit) else "Not Active"
            self.rows.append([sm, self.sm_max, pw, act(0x8), act(0x40), act(0x20), act(0x4)])
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            if self.nv is not None:
                self.sample()
                time.sleep(0.002)
                continue
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples: neither NVML nor nvidia-smi answered"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(str(r[col]).lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def cpu_oracle_world(T, grid):
    from oracle import oracle as orc
    pt = orc.PathTables(T["wp_x"], T["wp_y"], T["wp_psi"], T["wp_kappa"], T["wp_vref"], T["segment_lengths"],
                        T["border"], True)
    kmax = np.tan(0.66) / 0.12
    cfg = orc.mpc_cfg(N_HORIZON, [1.0, 0.0, 0.0], [0.5, 0.0], [1.0, 0.0, 0.0], [-np.inf] * 3, [np.inf] * 3,
                      [0.0, -kmax], [1.0, kmax], 4.0, 0.12, 0.06 / np.sqrt(2))
    return orc, orc.World(pt, cfg, grid.shape, T["origin"], float(T["resolution"]), 0.05)


def time_cpu_port(T, grid, states4xB, steps, threads=0):
    """The oracle port (reference algorithm, C, fp64, OpenMP over scenarios) on `states`; returns
    (solves/s, seconds, threads)."""
    orc, world = cpu_oracle_world(T, grid)
    B = states4xB.shape[1]
    st = np.ascontiguousarray(states4xB.T.copy())
    ctrl = np.zeros((B, 2 * N_HORIZON))
    infeas = np.zeros(B, np.int32)
    alive = np.ones(B, np.int32)
    nthr = threads or orc.num_threads()
    t0 = time.perf_counter()
    stats = world.batch_closed_loop(grid, st, ctrl, infeas, alive, steps, nthr)
    dt = time.perf_counter() - t0
    return stats[1] / dt, dt, nthr, stats


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path.  The reference is pure Python + OSQP + scikit-image, none of
    which exist on the GPU box, so this arm times the oracle port of the same algorithm (oracle/*.c: fp64,
    one scenario per OpenMP thread, all host cores) on a bounded sample of the same workload."""
    if rank != 0:
        return
    T, grid = load_track()
    B = args.batch
    sample = B  # every timed step = one closed-loop step of the whole per-GPU batch on the host cores
    states = scenario_states(T, B, 0, sample)
    # every host core this process may run on -- explicitly: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would
    # otherwise throttle this arm to a single thread when the driver launches it for N > 1
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count() or 1
    for _ in range(max(args.warmup, 0) and 1):
        time_cpu_port(T, grid, states[:, :64], 1, threads=ncores)
    per_step = []
    for _ in range(args.steps):
        v, dt, nthr, _ = time_cpu_port(T, grid, states, 1, threads=ncores)
        per_step.append((v, dt))
    value = float(np.mean([v for v, _ in per_step]))
    ms = float(np.mean([dt for _, dt in per_step]) * 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "path tracking, %d cars/GPU on sim_map, randomised start offsets (seed 2), N=30; "
                                   "reference arm steps a %d-car sample per step" % (B, sample),
                       "horizon": N_HORIZON, "batch_per_gpu": B, "eps_abs": 1e-3, "eps_rel": 1e-3, "cold_start": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "kind": "port",
                             "sample": "%d cars x 1 closed-loop step per timed step, OpenMP over cars" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = Python + OSQP + scikit-image, not installable offline; this arm is the C port of the "
                    "same algorithm (oracle/), which is FASTER than the reference's Python (no interpreter, no scipy "
                    "assembly at ~5 ms/step)"}
    emit_json(line)


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: whatever libraries print there meanwhile (NCCL's version banner, torchrun
    notes) is routed to stderr at the file-descriptor level until emit_json restores it."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="scenarios per GPU")
    ap.add_argument("--precision", type=int, default=0, help="0 = fp32 ADMM (production), 1 = fp64")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="cars in the CPU-baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="host time spent on the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="tracking", choices=["tracking", "obstacles"],
                    help="tracking = BASELINE configs[1] (headline); obstacles = configs[2] style: per-scenario random "
                         "obstacle sets, per-step raycast on per-scenario grids (use --batch 8192, seed 3)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import mpc_b200
    from mpc_b200 import _lib, distributed as D
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    T, grid = load_track()
    Bg = args.batch * world                      # weak scaling: fixed work per GPU
    lo, hi = D.shard_range(Bg, rank, world)
    B = hi - lo
    kind, seed = ("obstacles", 3) if args.workload == "obstacles" else ("tracking", 2)
    states = scenario_states(T, Bg, lo, hi, seed=seed, kind=kind)
    obstacles = scenario_states(T, Bg, lo, hi, seed=seed, kind=kind, return_obstacles=True) if kind == "obstacles" else None
    tab = _lib.path_table(T["wp_x"], T["wp_y"], T["wp_psi"], T["wp_kappa"], T["wp_vref"])
    lc = np.cumsum(T["segment_lengths"])

    def make_engine():
        e = mpc_b200.Engine(precision=args.precision)
        e.set_path(tab, lc, T["border"], True)
        e.set_base_grid(grid, T["origin"], float(T["resolution"]))
        if obstacles is not None:
            e.set_obstacles(obstacles[0], obstacles[1])
        e.scenarios_init(states)
        return e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    # ---------------- device-resident arm: K timed steps, CUDA events per step, L2 flushed in between ----
    eng = make_engine()
    for _ in range(args.warmup):
        eng.step()
    barrier()
    # Cars need 70+ steps from their start waypoint to the finish line (scenario_states); a run with more timed steps than
    # that would end with finished cars that the kernels skip.  Every RESTART timed steps the fleet is put back to its
    # state after the warm-up (poses, previous plans, infeasibility counters) -- between two timed steps, outside the
    # per-step event pairs.
    RESTART = 40
    snap = eng.scenarios_read()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = eng.launch_count()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        if k and k % RESTART == 0:
            eng.scenarios_set_state(snap["state"], snap["control"], snap["infeas"])
        flush.fill_(float(k))
        ev[k][0].record()
        eng.step()
        ev[k][1].record()
    while not ev[-1][1].query():  # the queue drains for tens of ms: sample the clocks under load from this thread too
        sampler.sample()
        time.sleep(0.001)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    launches = eng.launch_count() - l0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    if os.environ.get("MPC_BENCH_VERBOSE"):
        print("per-step ms:", ["%.3f" % v for v in ms_steps], file=sys.stderr)
    ms_total = D.max_over_ranks(float(np.sum(ms_steps)))
    out = eng.scenarios_read()
    iters_last = out["iters"].astype(np.int64)
    live = ((out["flags"] & (2 | 32)) == 0)
    n_live = int(live.sum())
    if n_live != B:
        print("WARNING: %d of %d cars finished / died during the timed region; value counts them as work" % (B - n_live, B),
              file=sys.stderr)
    value = Bg * args.steps / (ms_total * 1e-3)

    # ---------------- per-kernel durations (same workload, events around every kernel) -----------------
    eng.scenarios_set_state(snap["state"], snap["control"], snap["infeas"])
    eng.set_profiling(True)
    eng.run_closed_loop(min(args.steps, RESTART))
    prof, nl = eng.get_profile()
    kernel_ms = {k: v / max(n, 1) for (k, v), n in zip(prof.items(), nl)}
    eng.set_profiling(False)
    stats = eng.run_closed_loop(0)
    out2 = eng.scenarios_read()
    mean_iters = float(out2["iters"].mean())
    flops_per_launch = float(sum(flop_model(N_HORIZON, int(i)) for i in out2["iters"]))
    solve_ms = kernel_ms["assemble_solve"]
    achieved = flops_per_launch / (solve_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    # FP32 CUDA-core peak: MEASURED_PEAKS.json has no such figure, so the denominator is the FFMA issue rate measured on
    # this pool's B200 with tools/pipe_microbench.cu (profiles/measured_fp32_peak.json); nominal only if that is missing
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    fp32_src = "nominal 148 SM x 128 lanes x 2 x sm_max_mhz"
    try:
        mp_ = json.load(open(os.path.join(REPO, "profiles", "measured_fp32_peak.json")))
        fp32_peak = float(mp_["fp32_tflops"])
        fp32_src = ("measured: %.3f FFMA warp-instr/clk/SM x 32 x 2 x %d SMs x %.3f GHz (tools/pipe_microbench.cu, "
                    "profiles/measured_fp32_peak.json); nominal is %.1f" % (mp_["ffma_warp_instr_per_clk_per_sm"], mp_["sms"],
                                                                          mp_["clock_khz"] / 1e6, 148 * 128 * 2 * sm_max * 1e6 / 1e12))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    rollout_gbs = 100.0 * B / (kernel_ms["rollout"] * 1e-3) / 1e9 if kernel_ms["rollout"] > 0 else None
    # raycast: algorithmic bytes = 32 B x distinct 32-byte sectors of the bit grid holding a tested cell + 16N + 32
    # (SURVEY 8d); the oracle enumerates the tested cells, so it counts the sectors on a sample of scenarios
    ray_bytes = None
    if rank == 0:
        from oracle import oracle as orc
        pt = orc.PathTables(T["wp_x"], T["wp_y"], T["wp_psi"], T["wp_kappa"], T["wp_vref"], T["segment_lengths"],
                            T["border"], True)
        smg = 0.06 / np.sqrt(2)
        secs = []
        for b in range(0, B, max(B // 48, 1)):
            if obstacles is None:
                gb = grid
            else:
                gb = eng.get_grid(b)
            st_, _, _, _, ncell, nsec = orc.update_path_constraints(gb, T["origin"], float(T["resolution"]), pt,
                                                                    int(out2["wp_id"][b]) + 1, N_HORIZON, 2 * smg, smg,
                                                                    want_stats=True)
            if st_ == 0:
                secs.append(nsec)
        if secs:
            ray_bytes = 32.0 * float(np.mean(secs)) + 16 * N_HORIZON + 32
    ray_gbs = ray_bytes * B / (kernel_ms["raycast"] * 1e-3) / 1e9 if ray_bytes else None
    eng.close()

    # the same solve kernel with the machine evenly filled: 4096 cars are 2048 warps on 1184 resident warp slots
    # (1.73 waves, the second one 73 % full); 16 copies of the batch make the tail negligible.  Supplementary --
    # the headline roofline above stays on the BASELINE workload.
    saturated = None
    if world == 1 and obstacles is None and args.precision == 0:
        rep = 16
        e2 = mpc_b200.Engine(precision=args.precision)
        e2.set_path(tab, lc, T["border"], True)
        e2.set_base_grid(grid, T["origin"], float(T["resolution"]))
        e2.scenarios_init(np.ascontiguousarray(np.tile(states, (1, rep))))
        for _ in range(3):
            e2.step()
        e2.set_profiling(True)
        e2.run_closed_loop(max(args.steps // 2, 4))
        p2, n2 = e2.get_profile()
        e2.set_profiling(False)
        it2 = e2.scenarios_read()["iters"]
        ms2 = p2["assemble_solve"] / max(n2[2], 1)
        fl2 = float(sum(flop_model(N_HORIZON, int(i)) for i in it2))
        saturated = {"batch": int(B * rep), "assemble_solve_ms": ms2, "achieved": fl2 / (ms2 * 1e-3) / 1e12, "unit": "TFLOP/s",
                     "frac": fl2 / (ms2 * 1e-3) / 1e12 / fp32_peak, "qp_solves_per_sec_kernel_only": B * rep / (ms2 * 1e-3)}
        e2.close()

    # ---------------- end-to-end arm: host buffers, H2D + D2H inside the timed region -------------------
    # inputs (state[4][B]) and results (state, u[B][2], flags[B]) live in page-locked HOST memory; every step is
    # H2D + localise/raycast + assemble/solve/rollout + D2H, submitted as one graph launch by mpc_step_host
    def time_e2e(pinned):
        eng = make_engine()
        if pinned:
            hs, hu, hf = eng.host_io()
            hs[:] = states
        else:
            hs, hu, hf = states.copy(), np.zeros((B, 2)), np.zeros(B, np.int32)
        for _ in range(args.warmup):
            eng.step_host(hs, hu, hf)
        barrier()
        s0 = eng.scenarios_read() if args.steps > RESTART else None
        t_acc = 0.0
        for k in range(args.steps):
            if k and k % RESTART == 0:  # long runs only; not part of a step, so outside the clock
                hs[:] = s0["state"]
                eng.scenarios_set_state(None, s0["control"], s0["infeas"])
            t0 = time.perf_counter()
            eng.step_host(hs, hu, hf)  # synchronous: returns when the results are in the host buffers
            t_acc += time.perf_counter() - t0
        barrier()
        dt = D.max_over_ranks(t_acc)
        chk = np.array(hu, copy=True)  # the results are read on the host
        eng.close()
        return Bg * args.steps / dt, chk
    e2e_value, e2e_chk = time_e2e(True)
    e2e_pageable, e2e_chk2 = time_e2e(False)
    assert np.array_equal(e2e_chk, e2e_chk2, equal_nan=True) and np.nansum(np.abs(e2e_chk)) > 0, \
        "pinned and pageable step_host disagree"

    agg = D.allreduce_stats(stats)  # the one collective of the job: statistics only

    # ---------------- CPU baseline: the oracle port on a bounded sample, rank 0, N = 1 only --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and obstacles is None:
        # bounded sample: the whole batch, closed loop, for about args.cpu_seconds of host time (chunks of 8 steps)
        sample = min(args.cpu_sample, B)
        orc_, world_ = cpu_oracle_world(T, grid)
        st_ = np.ascontiguousarray(states[:, :sample].T.copy())
        ctrl_ = np.zeros((sample, 2 * N_HORIZON)); inf_ = np.zeros(sample, np.int32); alive_ = np.ones(sample, np.int32)
        nthr = orc_.num_threads()
        solves, nsteps, t0 = 0.0, 0, time.perf_counter()
        st_init = st_.copy()
        while time.perf_counter() - t0 < args.cpu_seconds and nsteps < 20000:
            stt = world_.batch_closed_loop(grid, st_, ctrl_, inf_, alive_, 8, nthr)
            solves += stt[1]; nsteps += 8
            if alive_.sum() < sample // 2:  # most cars have finished their lap: start them again
                st_[:] = st_init; ctrl_[:] = 0; inf_[:] = 0; alive_[:] = 1
        dt = time.perf_counter() - t0
        cpu = {"value": solves / dt, "unit": UNIT, "cores": nthr, "kind": "port",
               "sample": "%d of the %d cars x %d closed-loop steps (%.1f s of host time, %d QP solves), oracle/*.c fp64, "
                         "OpenMP over cars" % (sample, B, nsteps, dt, int(solves))}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == 0 else "f64", "data": "synthetic",
            "config": {"workload": ("path tracking, %d cars/GPU on sim_map, randomised start offsets (seed 2), N=30 "
                                    "(BASELINE configs[1]); one step = localise+raycast+QP solve+rollout for every car" % args.batch)
                       if obstacles is None else
                       ("obstacle avoidance, %d scenarios/GPU with randomised obstacle sets (seed 3), per-step raycast on "
                        "per-scenario grids, N=30 (BASELINE configs[2] style)" % args.batch),
                       "horizon": N_HORIZON, "batch_per_gpu": args.batch, "global_batch": Bg, "eps_abs": 1e-3,
                       "eps_rel": 1e-3, "cold_start": True, "l2": "flushed between timed steps (256 MB fill)",
                       "parallelism": "scenario shards, no data-path collective"},
            "closed_loop_steps_per_sec": value, "mean_admm_iters": mean_iters, "live_scenarios_rank0": n_live,
            "wall_ms_per_step_incl_flush": t_wall * 1e3 / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * B * 8),
                    "d2h_bytes_per_step": int(6 * B * 8 + 4 * B),
                    "path": "mpc_step_host on page-locked host buffers (H2D + 2 kernels + D2H = one graph launch), host "
                            "wall clock around each of the K synchronous calls, summed", "pageable_buffers_value": e2e_pageable},
            "gpu_launches": int(launches),
            "kernel_ms": kernel_ms,
            "roofline": {"kernel": ("assemble_solve_pair_kernel<16,loose> (K1+K2+K4b, paired-stage fp32)" if os.environ.get("MPC_ADMM_KERNEL", "p")[0] != "s"
                                    else "assemble_solve_kernel<float,5> (K1+K2)") if args.precision == 0 else "assemble_solve_kernel<double,5>",
                         "bound": "fp32_pipe" if args.precision == 0 else "fp64_pipe", "achieved": achieved,
                         "peak": fp32_peak if args.precision == 0 else fp32_peak / 2, "unit": "TFLOP/s",
                         "frac": achieved / (fp32_peak if args.precision == 0 else fp32_peak / 2),
                         "peak_source": fp32_src + " (MEASURED_PEAKS.json has no CUDA-core figure)",
                         "flops_per_launch": flops_per_launch,
                         "traffic": 4.02e6 * B / 4096 if args.precision == 0 else None,
                         "traffic_source": "dram__bytes_read+write of one launch at 4096 scenarios, ncu --set full "
                                           "(profiles/r1_assemble_solve_pair_fp32.txt), scaled by batch",
                         "model": "SURVEY 8d: iters*(340N+264) + 400(N+1) + (iters/25)*(142N+78) per instance, actual iteration counts",
                         "saturated": saturated},
            "roofline_hbm": [
                {"kernel": "rollout_kernel (K4)", "bound": "hbm", "achieved": rollout_gbs, "peak": hbm_peak, "unit": "GB/s",
                 "frac": (rollout_gbs / hbm_peak) if rollout_gbs else None, "bytes_per_instance": 100,
                 "note": "6-12 us launches: launch-latency bound at this batch size; 4.70 TB/s = 72 % of the HBM peak at 4 M "
                         "scenarios (tools/hbm_kernels.py, profiles/r1_hbm_kernels.jsonl)"},
                {"kernel": "raycast_kernel (K3)", "bound": "hbm", "achieved": ray_gbs, "peak": hbm_peak, "unit": "GB/s",
                 "frac": (ray_gbs / hbm_peak) if ray_gbs else None, "bytes_per_instance": ray_bytes,
                 "note": ("shared base grid: the 32 KB grid is L2-resident and staged once per CTA, the kernel is bound by "
                          "the per-cell walk, not by HBM") if obstacles is None else "per-scenario grids, row span staged per warp by TMA"}],
            "clocks": sampler.summary(),
            "stats": agg,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
