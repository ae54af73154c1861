"""Run-time-count lean kernels against the fixed instances and the block-of-roles kernel (rollout and per-step rates)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_ballenv_b200 import BallVecEnv
T, n = 200, 65536
for wl in ("c3", "w5"):
    spec = bench.workload_spec(wl)
    for name, envs in [("fixed", {}), ("run-time", {"BALLENV_LEAN_RT": "1"}), ("run-time G=1", {"BALLENV_LEAN_RT": "1", "BALLENV_LEAN_G": "1"}),
                       ("run-time G=2", {"BALLENV_LEAN_RT": "1", "BALLENV_LEAN_G": "2"}), ("roles", {"BALLENV_NO_LEAN": "1"})]:
        for mode in ("rollout", "per-step"):
            os.environ["BALLENV_NO_ROLLOUT"] = "1" if mode == "per-step" else "0"
            for k, v in envs.items():
                os.environ[k] = v
            env = BallVecEnv(n, window=spec["window"], config=bench.env_config(spec), seed=0, device="cuda:0")
            for k in envs:
                del os.environ[k]
            env.reset()
            a = torch.randint(0, 9, (T, n), device="cuda:0")
            out = env.alloc_rollout(T, keep_all_obs=True)
            for _ in range(2):
                env.step_many(a, keep_all_obs=True, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                env.step_many(a, keep_all_obs=True, out=out)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (4 * T)
            print("%s %-12s %-8s %.2f us/env-step-of-all-envs  %.2f G env-steps/s  errs=%d" %
                  (wl, name, mode, us, n / us / 1e3, env.error_flags()), flush=True)
            env.close()
