#!/usr/bin/env python
"""bench.py - env-steps/s of the fused step + window-observe hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c3|w5] [--impl reference]

One bench "step" = one rollout chunk: every environment of the job advances CHUNK env-steps through ONE
``ballenv_step_many`` call (one launch of the rollout kernel: a warp keeps its environments on chip for the whole
chunk), writing the whole chunk's observations / rewards / dones into a rollout buffer.
``value`` = env-steps of all ranks / max-over-ranks device time.  ``closed_loop`` repeats the measurement with one
launch per env-step (``ballenv_step``, what a policy-in-the-loop trainer calls); ``c5`` is that trainer (BASELINE.json
config 5); ``e2e`` goes through the host-buffer entry point of the C ABI; with N > 1 ``c4`` is BASELINE.json config 4.
See DESIGN.md (Measurement) for the byte model behind ``roofline``.

Workloads:
  c3  65 536 envs/GPU, WINDOW=10, 8 static + 24 moving obstacles ("dense moving")   <- headline
  w5  65 536 envs/GPU, WINDOW=5, the reference's default 13 static + 5 moving obstacles
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec"
UNIT = "env-steps/s"
CHUNK = 200          # env-steps per environment per bench step
ACTION_RING = 16     # pre-generated action chunks cycled through (> L2 in total)


# ----------------------------------------------------------------------------------------- workloads
def workload_spec(name):
    if name == "c3":
        goals = [(x, y) for y in (100, 200, 300, 400) for x in (50, 130, 210, 290, 370, 450)]
        return dict(name="c3", window=10, static_obstacles=8, dynamic_obstacles=24, speeds=[1] * 24, goals=goals,
                    change_step=50, rd_th_obs=60,
                    text="C3: 65536 envs/GPU, WINDOW=10, 8 static + 24 moving obstacles (dense), gym ruleset, "
                         "uniform 9-way actions, auto-reset + TimeLimit(1000), fp32 obs [N,104]")
    if name == "w5":
        goals = [(12, 122), (123, 93), (87, 150), (430, 440), (230, 11)]
        return dict(name="w5", window=5, static_obstacles=13, dynamic_obstacles=5, speeds=[1] * 5, goals=goals,
                    change_step=50, rd_th_obs=60,
                    text="65536 envs/GPU, WINDOW=5, 13 static + 5 moving obstacles (reference defaults), gym ruleset, "
                         "uniform 9-way actions, auto-reset + TimeLimit(1000), fp32 obs [N,29]")
    raise SystemExit("unknown workload %r" % name)


def alg_bytes_per_env_step(spec):
    """algorithmic bytes of one env-step (fp32 SoA state, int64 action, fp32 obs)."""
    ks, kd, w = spec["static_obstacles"], spec["dynamic_obstacles"], spec["window"]
    reads = 8 + 8 + 4 + 4 + 4 + 4 + 8 * ks + 8 * kd + 4 * kd + 8
    writes = 8 + 4 + 4 + 4 + 8 * kd + 4 * kd + 4 + 1 + 4 * (4 + w * w)
    return reads + writes


def moved_bytes_per_env_step(spec, chunk):
    """Bytes that must cross HBM per env-step in the rollout kernel: action in; observation, reward, done out;
    the state once per launch (read + write) amortised over the chunk."""
    w = spec["window"]
    io = 8 + 4 * (4 + w * w) + 4 + 1
    return io + (alg_bytes_per_env_step(spec) - io) / float(chunk)


def env_config(spec):
    from gym_ballenv_b200 import EnvConfig
    return EnvConfig(static_obstacles=spec["static_obstacles"], dynamic_obstacles=spec["dynamic_obstacles"],
                     obstacle_speed=list(spec["speeds"]), obs_goal_position=["%d,%d" % g for g in spec["goals"]],
                     time_step_for_change=spec["change_step"], rd_th_obs=spec["rd_th_obs"])


def config_dict(spec, n, world, chunk):
    """The same ``config`` object for both arms (the driver compares them)."""
    row = 4 + spec["window"] ** 2
    streamed = chunk * n * (8 + 4 * row + 4 + 1)
    return {"workload": spec["text"], "envs_per_gpu": n, "window": spec["window"],
            "env_steps_per_bench_step": chunk, "total_envs": world * n,
            "sharding": "envs partitioned by global id, %d per GPU, no data-path collective; one NCCL all-reduce of the "
                        "16-double episode statistics after the timed region, checked against the rank-local vectors" % n,
            "l2": "inputs larger than L2: every bench step streams %.2f GB of int64 actions + rollout "
                  "obs/reward/done per GPU through HBM (L2 is 126 MB); the env state itself is revisited every "
                  "env-step by design" % (streamed / 1e9)}


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while a timed region runs."""

    def __init__(self, uuid, index, period=0.001):
        super().__init__(daemon=True)
        self.period = period
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM)      # prime the (slow) first query
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _sample(self):
        nv = self._nv
        self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        for bit, name in ((nv.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                          (nv.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                          (nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                          (nv.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap"),
                          (nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown, "hw_power_brake")):
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if self._h is None:
            return
        while not self._stop_evt.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2.0)
        if self._h is not None and not self.samples:
            try:
                self._sample()
            except Exception:
                pass
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------- CPU legs
def _pool(procs):
    ctx = mp.get_context("spawn")
    return ctx.Pool(procs)


def cpu_baseline(spec, seconds=12.0, envs_per_worker=4):
    """Oracle port stepped in a multiprocessing loop on all host cores for a bounded time."""
    from oracle import cpu_rollout
    cores = os.cpu_count() or 1
    wl = {k: spec[k] for k in ("window", "static_obstacles", "dynamic_obstacles", "speeds", "goals",
                               "change_step", "rd_th_obs")}
    with _pool(cores) as pool:
        pool.map(cpu_rollout.run_steps, [(wl, i, envs_per_worker, 2) for i in range(cores)])   # import + reset
        t0 = time.perf_counter()
        res = pool.map(cpu_rollout.run_for, [(wl, i, envs_per_worker, seconds) for i in range(cores)])
        wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    return {"value": total / wall, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d worker processes x %d envs stepped+observed for %.0f s each (%d env-steps) of workload %s "
                      "through the Python oracle port of BallEnv.step + prep_state4" %
                      (cores, envs_per_worker, seconds, total, spec["name"])}


def run_reference(args, spec):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference cannot travel)
    on all host cores; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_rollout
    cores = os.cpu_count() or 1
    wl = {k: spec[k] for k in ("window", "static_obstacles", "dynamic_obstacles", "speeds", "goals",
                               "change_step", "rd_th_obs")}
    epw = 4
    # size a step to ~1 s per worker, and the whole run to <= ~3 min
    per_core = 600.0 if spec["window"] >= 10 else 3500.0
    budget = min(1.0, 170.0 / max(1, args.steps + args.warmup))
    n_steps = max(1, int(per_core * budget / epw))
    with _pool(cores) as pool:
        jobs = [(wl, i, epw, n_steps) for i in range(cores)]
        pool.map(cpu_rollout.run_steps, [(wl, i, epw, 1) for i in range(cores)])
        for _ in range(args.warmup):
            pool.map(cpu_rollout.run_steps, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(cpu_rollout.run_steps, jobs)
        wall = time.perf_counter() - t0
    total = cores * epw * n_steps * args.steps
    value = total / wall
    sample = ("each step = %d worker processes x %d envs x %d env-steps (step + prep_state4 W=%d) of workload %s "
              "through the Python oracle port (the reference is pure Python and is not on this box)" %
              (cores, epw, n_steps, spec["window"], spec["name"]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(spec, args.envs_per_gpu, max(1, args.gpus), args.chunk),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def kernel_facts(env, spec, n_steps):
    """Name of the kernel a step() / step_many() call launches and its ptxas figures (gym_ballenv_b200/build_info.json,
    written by build.py from `-Xptxas -v`)."""
    variant, lanes = env.kernel_variant(n_steps), env.kernel_lanes(n_steps)
    w, ks, kd = spec["window"], spec["static_obstacles"], spec["dynamic_obstacles"]
    roll = n_steps > 1
    if variant == "lean":
        name = "ballenv_lean_kernel<%d,%d,%d,lanes=%d,%s>" % (w, ks, kd, lanes, "rollout" if roll else "single-step")
        key = "_ZN7ballenv19ballenv_lean_kernelILi%dELi%dELi%dELi%dELb%dELb0EEEvNS_6ParamsE" % (w, ks, kd, lanes, 1 if roll else 0)
    else:
        name = "ballenv_kernel<float,%d,%s,%s>" % (w, "fast" if variant == "roles" else "generic",
                                                   "rollout" if roll else "single-step")
        key = None
    facts = {"kernel": name, "registers_per_thread": None, "spill_bytes": None}
    try:
        info = json.load(open(os.path.join(ROOT, "gym_ballenv_b200", "build_info.json")))
        if key in info:
            facts["registers_per_thread"] = info[key].get("registers")
            facts["spill_bytes"] = {"stores": info[key].get("spill_store_bytes"), "loads": info[key].get("spill_load_bytes")}
    except Exception:
        pass
    return facts


def measure(env, torch, spec, n, steps, warmup, chunk, dist, world, sampler=None, segments=5):
    """Device-timed rollout: EXACTLY `steps` bench steps between two events, barrier + synchronize on both sides, max
    over ranks.  The region is also cut into `segments` event-timed segments (best of 5) whose best rate
    is reported beside the whole-region value.  -> dict(ms, launches, clocks, best_launch_us)."""
    dev = env.device
    g = torch.Generator(device=dev).manual_seed(1 + env.global_env_offset)
    ring = min(ACTION_RING, steps + warmup)
    actions = torch.randint(0, 9, (ring, chunk, n), generator=g, device=dev, dtype=torch.int64)
    out = env.alloc_rollout(chunk, keep_all_obs=True)
    for k in range(warmup):
        env.step_many(actions[k % ring], keep_all_obs=True, out=out)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.start()
    l0 = env.launch_count
    segments = max(1, min(segments, steps))
    cuts = [steps * i // segments for i in range(segments + 1)]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(segments + 1)]
    evs[0].record()
    si = 1
    for k in range(steps):
        env.step_many(actions[(warmup + k) % ring], keep_all_obs=True, out=out)
        if k + 1 == cuts[si]:
            evs[si].record()
            si += 1
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler is not None else None
    ms = evs[0].elapsed_time(evs[-1])
    seg_us = [evs[i].elapsed_time(evs[i + 1]) * 1e3 / max(1, cuts[i + 1] - cuts[i]) for i in range(segments)]
    best = min(seg_us)
    launches = env.launch_count - l0
    if world > 1:
        t = torch.tensor([ms, best, float(launches)], device=dev, dtype=torch.float64)
        dist.all_reduce(t[:2], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[2:], op=dist.ReduceOp.SUM)
        ms, best, launches = float(t[0]), float(t[1]), int(t[2])
    return dict(ms=ms, launches=launches, clocks=clocks, best_launch_us=best)


def reduced_stats(env, torch, dist, world):
    """The job's only collective, checked: one all-reduce (NCCL over NVLink) of the 16-double statistics vector, and
    the same sum computed from an all-gather of the rank-local vectors.  -> (summed vector as dict, check dict)."""
    from gym_ballenv_b200 import STAT_NAMES, allreduce_stats
    local = env.stats_tensor.detach().clone()
    torch.cuda.synchronize(env.device)
    total, work = allreduce_stats(env.stats_tensor, async_op=True)
    if work is not None:
        work.wait()
    torch.cuda.synchronize(env.device)
    check = {"collective": "none (1 rank)" if world == 1 else "nccl all_reduce(SUM) of float64[16]", "ok": True}
    if world > 1:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        ref = torch.stack(parts).sum(0)
        # counters are integers held in doubles: exact whatever the order of the additions; the return sum is a sum of
        # fractions, and NCCL's reduction order is not the all-gather's: equal to rounding (seen differing in the last
        # bit on 4 ranks, not on 2 or 8)
        exact = torch.ones_like(ref, dtype=torch.bool)
        exact[STAT_NAMES.index("return_sum")] = False
        check["ok"] = bool(torch.equal(ref[exact], total[exact]) and
                           torch.allclose(ref[~exact], total[~exact], rtol=1e-12, atol=0.0))
        check["max_rel_diff"] = float(((ref - total).abs() / ref.abs().clamp_min(1e-300)).max())
        check["episodes_per_rank"] = [float(p_[0]) for p_ in parts]
    vec = total.cpu().tolist()
    return {name: vec[i] for i, name in enumerate(STAT_NAMES)}, check


def measure_e2e(env, torch, n, steps, chunk, dist, world, mode="many", t_call=25):
    """The same rollout through the host-buffer entry points of the C ABI: pinned host actions in, host obs / reward /
    done out, EVERY env-step.  mode "many": ballenv_step_many_host, t_call env-steps per call, copies and kernels of
    consecutive steps overlapped; mode "step": one synchronous ballenv_step_host per env-step."""
    from gym_ballenv_b200.hostmem import pinned_empty
    dev = env.device
    g = torch.Generator().manual_seed(7 + env.global_env_offset)
    obs_dtype = env._bufs[0]["obs"].dtype
    t_call = max(1, min(t_call, chunk))
    calls = max(1, chunk // t_call)
    act = pinned_empty((t_call, n), torch.int64)
    act.copy_(torch.randint(0, 9, (t_call, n), generator=g, dtype=torch.int64))
    if mode == "many":
        obs = pinned_empty((t_call, n, env.obs_row), obs_dtype)
        rew = pinned_empty((t_call, n), torch.float32)
        done = pinned_empty((t_call, n), torch.uint8)
        run = lambda: env.step_many_host(act, obs, rew, done)
    else:
        obs = pinned_empty((n, env.obs_row), obs_dtype)
        rew = pinned_empty((n,), torch.float32)
        done = pinned_empty((n,), torch.uint8)

        def run():
            for t in range(t_call):
                env.step_host(act[t], obs, rew, done)
    run()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = env.launch_count
    ev0.record()
    for k in range(steps * calls):
        run()
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    env_steps = steps * calls * t_call
    row_b = env.obs_row * (4 if obs_dtype in (torch.float32, torch.int32) else 1)
    h2d = calls * t_call * n * 8
    d2h = calls * t_call * n * (row_b + 4 + 1)
    return dict(ms=ms, env_steps_per_env=env_steps, h2d=h2d, d2h=d2h, launches=env.launch_count - l0)


def measure_c5(torch, dev, iterations=12, n_steps=32, n=16384):
    """BASELINE.json config 5: the actor-critic loop of examples/ball_cnn_ac3.py:528-646 driving 16 K GPU environments
    end to end.  An iteration = one n_steps-step policy-in-the-loop rollout of all environments - ONE launch
    (ballenv_rollout_policy: the environments' own lanes evaluate Policy(5) and draw the action between two steps) - and
    the batched finish_episode update: discounted returns (ballenv_discounted_returns), normalisation, loss and all
    gradients from hand-written kernels (ballenv_a2c_grads: two launches, checked against autograd), torch's Adam; the
    whole iteration replayed as one CUDA graph (a2c.GraphedTrainer).  Beside it: the rollout alone, the same iteration
    with the autograd update, and with the per-step torch policy (GraphedRollout: Policy forward + multinomial +
    ballenv_step per env-step, replayed as a graph; what this leg measured before the fused launches existed)."""
    from gym_ballenv_b200 import BallVecEnv
    from gym_ballenv_b200.a2c import FusedRollout, GraphedRollout, GraphedTrainer, Policy, a2c_loss

    def timed(fn, reps):
        torch.cuda.synchronize(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            out = fn()
        ev1.record()
        torch.cuda.synchronize(dev)
        return ev0.elapsed_time(ev1) / reps, out

    torch.manual_seed(0)
    env = BallVecEnv(n, window=5, seed=0, device=dev)
    policy = Policy(5).to(dev)
    env.reset()
    trainer = GraphedTrainer(env, policy, n_steps)
    trainer.step()      # three eager iterations, then the capture
    trainer.step()
    l0 = env.launch_count
    ms_it, loss = timed(trainer.step, iterations)
    launches = env.launch_count - l0      # (replays do not pass through the library: counted below)
    roll = FusedRollout(env, policy, n_steps)
    roll.run()
    ms_roll, _ = timed(roll.run, iterations)
    errs = env.error_flags()
    env.close()
    torch.manual_seed(0)
    env = BallVecEnv(n, window=5, seed=0, device=dev)
    policy_ag = Policy(5).to(dev)
    env.reset()
    trainer_ag = GraphedTrainer(env, policy_ag, n_steps, fused_update=False)
    trainer_ag.step()
    trainer_ag.step()
    ms_ag, _ = timed(trainer_ag.step, iterations)
    errs |= env.error_flags()
    env.close()

    # the per-step torch policy, for comparison
    torch.manual_seed(0)
    env = BallVecEnv(n, window=5, seed=0, device=dev)
    policy2 = Policy(5).to(dev)
    opt = torch.optim.Adam(policy2.parameters(), lr=1e-3)
    env.reset()
    groll = GraphedRollout(env, policy2, n_steps)

    def torch_iteration():
        raw = groll.run()
        batch = groll.evaluate(raw)
        with torch.no_grad():
            _, v_last = policy2(raw["obs"][n_steps])
        l = a2c_loss(batch, 0.99, bootstrap=v_last.squeeze(-1))
        opt.zero_grad(set_to_none=True)
        l.backward()
        opt.step()
        return l

    for _ in range(3):
        torch_iteration()
    ms_torch, _ = timed(torch_iteration, max(4, iterations // 2))
    env.close()

    # the same loop at the headline size (65 536 environments), and the pixel policies' observation (rgb patches)
    torch.manual_seed(0)
    n_big = 65536
    env = BallVecEnv(n_big, window=5, seed=0, device=dev)
    policy_big = Policy(5).to(dev)
    env.reset()
    trainer_big = GraphedTrainer(env, policy_big, n_steps)
    trainer_big.step()
    trainer_big.step()
    ms_big, _ = timed(trainer_big.step, iterations)
    roll_big = FusedRollout(env, policy_big, n_steps)
    roll_big.run()
    ms_roll_big, _ = timed(roll_big.run, iterations)
    errs |= env.error_flags()
    env.close()
    env = BallVecEnv(n, window=5, seed=0, device=dev)
    env.reset()
    env.step_many(torch.randint(0, 9, (100, n), device=dev))
    patches = env.rgb_patches()
    ms_patch, _ = timed(lambda: env.rgb_patches(out=patches), 10)
    errs |= env.error_flags()
    env.close()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak = 6650.0
    return {"workload": "config 5: 16384 envs, WINDOW=5, reference defaults (13 + 5 obstacles), Policy(5) MLP 29-128-{9,1} in "
                        "the loop (Categorical by inverse CDF of the env's Philox stream), %d-step rollouts in ONE launch "
                        "(ballenv_rollout_policy), batched finish_episode (loss + gradients by ballenv_a2c_grads) + Adam per "
                        "rollout, the iteration replayed as one CUDA graph" % n_steps,
            "value": n * n_steps / (ms_it * 1e-3), "unit": UNIT, "ms_per_iteration": ms_it,
            "rollout_only": {"value": n * n_steps / (ms_roll * 1e-3), "unit": UNIT, "ms_per_rollout": ms_roll,
                             "us_per_env_step_of_all_envs": ms_roll * 1e3 / n_steps},
            "autograd_update": {"value": n * n_steps / (ms_ag * 1e-3), "unit": UNIT, "ms_per_iteration": ms_ag,
                                "what": "same rollout launch, update by torch autograd over the stored pairs"},
            "torch_policy_per_step": {"value": n * n_steps / (ms_torch * 1e-3), "unit": UNIT, "ms_per_iteration": ms_torch,
                                      "what": "GraphedRollout: torch Policy forward + multinomial + ballenv_step per env-step "
                                              "(one CUDA graph per rollout), same update issued eagerly"},
            "at_65536_envs": {"value": n_big * n_steps / (ms_big * 1e-3), "unit": UNIT, "ms_per_iteration": ms_big,
                              "rollout_only": n_big * n_steps / (ms_roll_big * 1e-3), "ms_per_rollout": ms_roll_big},
            "rgb_patches": {"value": n / (ms_patch * 1e-3), "unit": "patches/s", "us_per_launch": ms_patch * 1e3,
                            "what": "ballenv_observe_patches: float32 [16384, 3, 40, 40] (extract_patch of the pixel policies)",
                            "hbm_frac_on_written_bytes": patches.numel() * 4 / (ms_patch * 1e-3) / 1e9 / peak},
            "env_kernel": "ballenv_lean_kernel<5,13,5,lanes=2,rollout,policy>",
            "gpu_launches_per_iteration": 4, "iterations": iterations,
            "loss_finite": bool(torch.isfinite(loss).item()), "device_error_flags": errs}


def run_gpu(args, spec):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    # pinned host buffers of the end-to-end legs: this rank's CPUs and pages on the GPU's NUMA node (torchrun does not bind)
    from gym_ballenv_b200.hostmem import bind_to_gpu_numa, copy_ceiling_gbs
    all_cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from gym_ballenv_b200 import make_sharded_env
    n = args.envs_per_gpu
    chunk = args.chunk

    def make_env(sp, n_=None, **kw):   # weak scaling: n envs per GPU, global ids [rank * n, (rank + 1) * n)
        n_ = n if n_ is None else n_
        e = make_sharded_env(world * n_, rank=rank, world=world, device=dev, window=sp["window"],
                             config=env_config(sp), seed=0, **kw)
        assert e.num_envs == n_ and e.global_env_offset == rank * n_
        e.reset()
        return e

    env = make_env(spec)
    assert env.launch_count == 1
    props = torch.cuda.get_device_properties(dev)
    uuid = "GPU-%s" % props.uuid if hasattr(props, "uuid") else ""
    sampler = ClockSampler(uuid, local)
    m = measure(env, torch, spec, n, args.steps, args.warmup, chunk, dist, world, sampler)
    ms, launches, clocks = m["ms"], m["launches"], m["clocks"]
    env_steps = float(world) * n * chunk * args.steps
    value = env_steps / (ms * 1e-3)
    # the job's only collective, checked: summed statistics == sum of the rank-local vectors, and every step counted
    stats, stats_check = reduced_stats(env, torch, dist, world)
    stats_check["steps_expected"] = float(world) * n * chunk * (args.steps + args.warmup)
    stats_check["ok"] = bool(stats_check["ok"] and stats["steps"] == stats_check["steps_expected"])
    balg = alg_bytes_per_env_step(spec)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_us = ms * 1e3 / args.steps                      # average duration of one rollout launch (CHUNK env-steps)
    moved = moved_bytes_per_env_step(spec, chunk)
    achieved = moved * n * chunk / (kernel_us * 1e-6) / 1e9            # bytes that must cross HBM, per second
    achieved_alg = balg * n * chunk / (kernel_us * 1e-6) / 1e9         # the contract's algorithmic bytes, per second
    traffic, traffic_source = None, None
    try:   # per-launch DRAM bytes of the step kernel from the committed ncu capture (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(spec["name"])
        if tr:
            traffic = tr["dram_bytes_per_env_step"] * n * chunk
            traffic_source = ("not measured in this run: %s (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set "
                              "full` launch, per env-step) x env-steps of one launch" % tr.get("source", "profiles/traffic.json"))
    except Exception:
        pass
    facts = kernel_facts(env, spec, chunk)

    launches_per_step = launches / float(world * args.steps)
    # closed loop: the same rollout with one launch per env-step (ballenv_step), i.e. what a trainer with the policy in
    # the loop pays; bounded to a few chunks
    closed = None
    if args.closed_loop_steps > 0:
        os.environ["BALLENV_NO_ROLLOUT"] = "1"
        env_c = make_env(spec)
        os.environ["BALLENV_NO_ROLLOUT"] = "0"
        cs = args.closed_loop_steps
        mc = measure(env_c, torch, spec, n, cs, 3, chunk, dist, world)
        us = mc["ms"] * 1e3 / (cs * chunk)
        closed = {"value": float(world) * n * chunk * cs / (mc["ms"] * 1e-3), "unit": UNIT, "avg_launch_us": us,
                  "best_of_5_launch_us": mc["best_launch_us"] / chunk, "gpu_launches": mc["launches"],
                  "roofline_frac_alg": balg * n / (us * 1e-6) / 1e9 / peak, "alg_bytes_per_env_step": balg,
                  "note": "one launch per env-step, state through L2 / HBM every step (programmatic dependent launch: the "
                          "next launch's blocks start while this one drains)"}
        closed.update(kernel_facts(env_c, spec, 1))
        env_c.close()

    # ---- end to end through the C ABI with host buffers
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    row_bytes = env.obs_row * 4
    ceiling = copy_ceiling_gbs(dev, n * row_bytes, iters=10)          # all ranks at once: the box's D2H ceiling per GPU
    if world > 1:
        tt = torch.tensor([ceiling], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        ceiling = float(tt[0])
    em = measure_e2e(env, torch, n, e2e_steps, chunk, dist, world, "many")
    e2e_value = float(world) * n * em["env_steps_per_env"] / (em["ms"] * 1e-3)
    e2e_gbs = em["d2h"] * e2e_steps / (em["ms"] * 1e-3) / 1e9
    es = measure_e2e(env, torch, n, 1, chunk, dist, world, "step")
    e2e_serial = {"value": float(world) * n * es["env_steps_per_env"] / (es["ms"] * 1e-3), "unit": UNIT,
                  "api": "ballenv_step_host, one synchronous call per env-step (round 1's e2e)"}
    errs = env.error_flags()
    e2e_alt = {}
    if args.e2e_u8:   # the same host round trip with narrower observation rows (identical 0 / 1 values)
        for key, dt in (("e2e_uint8_obs", torch.uint8), ("e2e_bits_obs", "bits")):
            env8 = make_env(spec, obs_dtype=dt)
            m8 = measure_e2e(env8, torch, n, e2e_steps, chunk, dist, world, "many")
            e2e_alt[key] = {"value": float(world) * n * m8["env_steps_per_env"] / (m8["ms"] * 1e-3), "unit": UNIT,
                            "h2d_bytes_per_step": m8["h2d"], "d2h_bytes_per_step": m8["d2h"],
                            "obs": "uint8 rows" if key == "e2e_uint8_obs" else "bit-packed rows (uint32 words)",
                            "kernel": env8.kernel_variant(1)}
            env8.close()

    secondary = None
    if args.secondary and spec["name"] == "c3":
        sp2 = workload_spec("w5")
        env.close()
        env2 = make_env(sp2)
        st2 = max(5, args.steps // 4)
        m2 = measure(env2, torch, sp2, n, st2, 3, chunk, dist, world)
        v2 = float(world) * n * chunk * st2 / (m2["ms"] * 1e-3)
        b2 = alg_bytes_per_env_step(sp2)
        secondary = {"workload": sp2["text"], "value": v2, "unit": UNIT,
                     "frac": (v2 / world) * moved_bytes_per_env_step(sp2, chunk) / 1e9 / peak,
                     "frac_alg": (v2 / world) * b2 / 1e9 / peak, "alg_bytes_per_env_step": b2,
                     "hbm_bytes_per_env_step_moved": moved_bytes_per_env_step(sp2, chunk)}
        secondary.update(kernel_facts(env2, sp2, chunk))
        env2.close()
        if args.closed_loop_steps > 0:      # the same configuration with one launch per env-step
            os.environ["BALLENV_NO_ROLLOUT"] = "1"
            env2c = make_env(sp2)
            os.environ["BALLENV_NO_ROLLOUT"] = "0"
            cs2 = args.closed_loop_steps
            mc2 = measure(env2c, torch, sp2, n, cs2, 3, chunk, dist, world)
            secondary["closed_loop"] = {"value": float(world) * n * chunk * cs2 / (mc2["ms"] * 1e-3), "unit": UNIT,
                                        "avg_launch_us": mc2["ms"] * 1e3 / (cs2 * chunk),
                                        "best_of_5_launch_us": mc2["best_launch_us"] / chunk,
                                        "kernel": kernel_facts(env2c, sp2, 1)["kernel"]}
            env2c.close()

    # BASELINE.json config 4: 2^20 environments over the job's GPUs (W=5, reference defaults), stats all-reduced
    c4 = None
    if world > 1 and args.c4:
        sp4 = workload_spec("w5")
        n4 = (1 << 20) // world
        env4 = make_env(sp4, n_=n4)
        st4 = max(5, args.steps // 10)
        m4 = measure(env4, torch, sp4, n4, st4, 3, chunk, dist, world)
        stats4, check4 = reduced_stats(env4, torch, dist, world)
        check4["steps_expected"] = float(world) * n4 * chunk * (st4 + 3)
        check4["ok"] = bool(check4["ok"] and stats4["steps"] == check4["steps_expected"])
        c4 = {"workload": "config 4: %d envs total = %d per GPU, WINDOW=5, 13 + 5 obstacles" % (world * n4, n4),
              "value": float(world) * n4 * chunk * st4 / (m4["ms"] * 1e-3), "unit": UNIT, "steps": st4,
              "episode_stats": {k: stats4[k] for k in ("episodes", "goals", "hits_static", "hits_dynamic", "timeouts", "steps")},
              "stats_check": check4}
        env4.close()

    c5 = measure_c5(torch, dev) if (rank == 0 and args.c5) else None
    if world > 1:
        dist.barrier()

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if all_cpus is not None:
            os.sched_setaffinity(0, all_cpus)   # the CPU baseline runs on all host cores, not on the GPU's node only
        base = cpu_baseline(spec, seconds=args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(spec, n, world, chunk),
            "roofline": dict({"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                              "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                              "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                              "hbm_bytes_per_env_step_moved": moved, "env_steps_per_launch": n * chunk,
                              "avg_launch_us": kernel_us, "best_of_5_launch_us": m["best_launch_us"],
                              "launches_per_bench_step": launches_per_step,
                              "achieved_alg": achieved_alg, "frac_alg": achieved_alg / peak, "alg_bytes_per_env_step": balg,
                              "note": "achieved / frac = the bytes that must cross HBM (action in; observation, reward, "
                                      "done out; the state once per launch) per second, against the measured copy peak. "
                                      "achieved_alg / frac_alg use SURVEY 8(d)'s algorithmic bytes, which count the state "
                                      "as read and written every env-step although the rollout kernel keeps it on chip: "
                                      "that figure can exceed 1 and is not an HBM fraction."}, **facts),
            "closed_loop": closed,
            "c5": c5,
            "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": em["h2d"], "d2h_bytes_per_step": em["d2h"],
                    "steps": e2e_steps,
                    "api": "ballenv_step_many_host (C ABI), 25 env-steps per call: pinned host actions in and fp32 obs / "
                           "reward / done out EVERY env-step, copies of consecutive steps overlapped with the kernels",
                    "pcie": {"d2h_ceiling_gbs_per_gpu": ceiling, "achieved_d2h_gbs_per_gpu": e2e_gbs,
                             "pcie_frac": e2e_gbs / ceiling if ceiling else None,
                             "how": "ceiling = bare cudaMemcpyAsync of one step's rows (%d MB) device -> pinned host, all "
                                    "ranks at once, min over ranks" % (n * row_bytes // (1 << 20))},
                    "numa": numa},
            "e2e_step_host": e2e_serial,
            "gpu_launches": launches,
            "clocks": clocks,
            "episode_stats": {k: stats[k] for k in ("episodes", "goals", "hits_static", "hits_dynamic", "timeouts", "steps")},
            "stats_check": stats_check,
            "device_error_flags": errs,
        }
        line.update(e2e_alt)
        if secondary is not None:
            line["secondary"] = secondary
        if c4 is not None:
            line["c4"] = c4
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "w5"])
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--chunk", type=int, default=CHUNK)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--closed-loop-steps", type=int, default=5)
    ap.add_argument("--e2e-u8", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", type=int, default=1)
    ap.add_argument("--c4", type=int, default=1, help="N > 1: also run config 4 (2^20 envs over the job's GPUs)")
    ap.add_argument("--c5", type=int, default=1, help="also run config 5 (A2C loop, 16 K envs) on rank 0")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    spec = workload_spec(args.workload)
    if args.impl == "reference":
        run_reference(args, spec)
    else:
        run_gpu(args, spec)


if __name__ == "__main__":
    main()
