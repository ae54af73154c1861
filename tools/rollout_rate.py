"""Per-env-step device time of ballenv_step_many: rollout kernel (one launch for T steps) vs one launch per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_ballenv_b200 import BallVecEnv
T = int(os.environ.get("T", "200"))
NS = [int(x) for x in os.environ.get("NS", "65536,262144").split(",")]
for wl in ("c3", "w5"):
    spec = bench.workload_spec(wl)
    for n in NS:
        for mode in ("rollout", "per-step"):
            os.environ["BALLENV_NO_ROLLOUT"] = "1" if mode == "per-step" else "0"
            env = BallVecEnv(n, window=spec["window"], config=bench.env_config(spec), seed=0, device="cuda:0")
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
            print("%s n=%7d %-8s %.2f us/env-step-of-all-envs  %.2f G env-steps/s  errs=%d" %
                  (wl, n, mode, us, n / us / 1e3, env.error_flags()), flush=True)
            env.close()
