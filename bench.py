#!/usr/bin/env python
"""bench.py - env-steps/s of the fused step + window-observe hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c3|w5] [--impl reference]

One bench "step" = one rollout chunk: every environment of the job advances CHUNK env-steps through ONE
``ballenv_step_many`` call (one launch of the rollout kernel: the block keeps its 32 environments on chip for
the whole chunk), writing the whole chunk's observations / rewards / dones into a rollout buffer.
``value`` = env-steps of all ranks / max-over-ranks device time.  ``closed_loop`` repeats the measurement with one
launch per env-step (``ballenv_step``, what a policy-in-the-loop trainer calls).  See DESIGN.md (Measurement) for
the byte model behind ``roofline``.

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
            "sharding": "envs partitioned by global id, %d per GPU, no data-path collective; one async NCCL "
                        "all-reduce of the 16-double episode statistics per bench step" % n,
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
def measure(env, torch, spec, n, steps, warmup, chunk, dist, world, sampler=None):
    """Device-timed rollout: returns (elapsed_ms max over ranks, launches in the timed region)."""
    from gym_ballenv_b200 import allreduce_stats
    dev = env.device
    g = torch.Generator(device=dev).manual_seed(1 + env.global_env_offset)
    ring = min(ACTION_RING, steps + warmup)
    actions = torch.randint(0, 9, (ring, chunk, n), generator=g, device=dev, dtype=torch.int64)
    out = env.alloc_rollout(chunk, keep_all_obs=True)
    stats_work = None
    for k in range(warmup):
        env.step_many(actions[k % ring], keep_all_obs=True, out=out)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.start()
    l0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(steps):
        env.step_many(actions[(warmup + k) % ring], keep_all_obs=True, out=out)
        if world > 1:   # the job's only collective: episode statistics, off the critical path
            _, stats_work = allreduce_stats(env.stats_tensor, async_op=True)
    ev1.record()
    torch.cuda.synchronize(dev)
    if stats_work is not None:
        stats_work.wait()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler is not None else None
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_count - l0
    if world > 1:
        t = torch.tensor([ms, float(launches)], device=dev, dtype=torch.float64)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        ms, launches = float(t[0]), int(t[1])
    bytes_streamed = actions[0].numel() * 8 + sum(o.numel() * o.element_size() for o in out)
    return ms, launches, clocks, bytes_streamed


def measure_e2e(env, torch, n, steps, chunk, dist, world):
    """Same rollout through ballenv_step_host: pinned host actions in, host obs/reward/done out, every env-step."""
    dev = env.device
    g = torch.Generator().manual_seed(7 + env.global_env_offset)
    act = torch.randint(0, 9, (chunk, n), generator=g, dtype=torch.int64).pin_memory()
    obs = torch.empty((n, env.obs_row), dtype=env._bufs[0]["obs"].dtype).pin_memory()
    rew = torch.empty(n, dtype=torch.float32).pin_memory()
    done = torch.empty(n, dtype=torch.uint8).pin_memory()
    for t in range(3):
        env.step_host(act[t], obs, rew, done)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = env.launch_count
    ev0.record()
    for k in range(steps):
        for t in range(chunk):
            env.step_host(act[t], obs, rew, done)
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    h2d = chunk * act[0].numel() * 8
    d2h = chunk * (obs.numel() * obs.element_size() + rew.numel() * 4 + done.numel())
    return ms, h2d, d2h, env.launch_count - l0


def run_gpu(args, spec):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
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

    def make_env(sp):   # weak scaling: n envs per GPU, global ids [rank * n, (rank + 1) * n)
        e = make_sharded_env(world * n, rank=rank, world=world, device=dev, window=sp["window"],
                             config=env_config(sp), seed=0)
        assert e.num_envs == n and e.global_env_offset == rank * n
        e.reset()
        return e

    env = make_env(spec)
    assert env.launch_count == 1
    props = torch.cuda.get_device_properties(dev)
    uuid = "GPU-%s" % props.uuid if hasattr(props, "uuid") else ""
    sampler = ClockSampler(uuid, local)
    ms, launches, clocks, streamed = measure(env, torch, spec, n, args.steps, args.warmup, chunk, dist, world, sampler)
    env_steps = float(world) * n * chunk * args.steps
    value = env_steps / (ms * 1e-3)
    balg = alg_bytes_per_env_step(spec)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_us = ms * 1e3 / args.steps                      # average duration of one rollout launch (CHUNK env-steps)
    achieved = balg * n * chunk / (kernel_us * 1e-6) / 1e9
    moved = moved_bytes_per_env_step(spec, chunk)
    traffic = None
    try:   # per-launch DRAM bytes of the step kernel from the committed ncu capture (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(spec["name"])
        if tr:   # measured DRAM bytes per env-step (ncu --set full, see profiles/) x env-steps of one launch
            traffic = tr["dram_bytes_per_env_step"] * n * chunk
    except Exception:
        pass

    launches_per_step = launches / float(world * args.steps)
    # closed loop: the same rollout with one launch per env-step (ballenv_step), i.e. what a trainer with the policy
    # in the loop pays; bounded to a few chunks
    closed = None
    if args.closed_loop_steps > 0:
        os.environ["BALLENV_NO_ROLLOUT"] = "1"
        env_c = make_env(spec)
        os.environ["BALLENV_NO_ROLLOUT"] = "0"
        cs = args.closed_loop_steps
        msc, lc, _, _ = measure(env_c, torch, spec, n, cs, 3, chunk, dist, world)
        us = msc * 1e3 / (cs * chunk)
        closed = {"value": float(world) * n * chunk * cs / (msc * 1e-3), "unit": UNIT, "avg_launch_us": us,
                  "gpu_launches": lc, "kernel": "ballenv_kernel<float,%d,fast,single-step>" % spec["window"],
                  "roofline_frac": balg * n / (us * 1e-6) / 1e9 / peak, "alg_bytes_per_env_step": balg}
        env_c.close()

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ems, h2d, d2h, e2e_launches = measure_e2e(env, torch, n, e2e_steps, chunk, dist, world)
    e2e_value = float(world) * n * chunk * e2e_steps / (ems * 1e-3)
    errs = env.error_flags()
    stats = env.stats()
    # the same host round trip with the observation rows as uint8 (identical 0/1 values, a quarter of the PCIe bytes;
    # the reference casts its observation with .float() anyway, examples/ball_cnn_ac3.py:211) - reported beside e2e
    e2e_u8 = None
    if args.e2e_u8:
        from gym_ballenv_b200 import make_sharded_env as _mk
        env8 = _mk(world * n, rank=rank, world=world, device=dev, window=spec["window"], config=env_config(spec),
                   seed=0, obs_dtype=torch.uint8)
        env8.reset()
        ms8, h8, d8, _ = measure_e2e(env8, torch, n, e2e_steps, chunk, dist, world)
        e2e_u8 = {"value": float(world) * n * chunk * e2e_steps / (ms8 * 1e-3), "unit": UNIT,
                  "h2d_bytes_per_step": h8, "d2h_bytes_per_step": d8, "obs": "uint8"}
        env8.close()

    secondary = None
    if args.secondary and spec["name"] == "c3":
        sp2 = workload_spec("w5")
        env.close()
        env2 = make_env(sp2)
        ms2, _, _, _ = measure(env2, torch, sp2, n, max(3, args.steps // 4), 3, chunk, dist, world)
        v2 = float(world) * n * chunk * max(3, args.steps // 4) / (ms2 * 1e-3)
        b2 = alg_bytes_per_env_step(sp2)
        secondary = {"workload": sp2["text"], "value": v2, "unit": UNIT,
                     "roofline_frac": (v2 / world) * b2 / 1e9 / peak, "alg_bytes_per_env_step": b2,
                     "frac_moved": (v2 / world) * moved_bytes_per_env_step(sp2, chunk) / 1e9 / peak}
        env2.close()

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(spec, seconds=args.cpu_seconds)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(spec, n, world, chunk),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                         "kernel": "ballenv_kernel<float,%d,fast,rollout>" % spec["window"],
                         "alg_bytes_per_env_step": balg, "env_steps_per_launch": n * chunk,
                         "avg_launch_us": kernel_us, "launches_per_bench_step": launches_per_step,
                         "note": "achieved = SURVEY 8(d) algorithmic bytes x env-steps per launch / launch time; the "
                                 "rollout kernel keeps the state on chip, so only hbm_bytes_per_env_step_moved must "
                                 "cross HBM per env-step (frac_moved is that stream against the same peak)",
                         "hbm_bytes_per_env_step_moved": moved,
                         "frac_moved": moved * n * chunk / (kernel_us * 1e-6) / 1e9 / peak},
            "closed_loop": closed,
            "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "ballenv_step_host (C ABI), pinned host buffers, fp32 obs to host "
                                               "every env-step"},
            "e2e_uint8_obs": e2e_u8,
            "gpu_launches": launches,
            "clocks": clocks,
            "episode_stats": {k: stats[k] for k in ("episodes", "goals", "hits_static", "hits_dynamic", "timeouts")},
            "device_error_flags": errs,
        }
        if secondary is not None:
            line["secondary"] = secondary
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
