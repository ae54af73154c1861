"""Profiling experiment: launch-to-launch time of the step kernel with stages disabled (BALLENV_DEBUG_SKIP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_ballenv_b200 import BallVecEnv
T = 200
for wl in ("c3",):
    spec = bench.workload_spec(wl)
    for n in [int(x) for x in os.environ.get("NS", "65536").split(",")]:
        for skip in (0, 2, 4, 8, 16, 2 + 4, 4 + 8, 2 + 4 + 8, 2 + 4 + 8 + 16):
            os.environ["BALLENV_DEBUG_SKIP"] = str(skip)
            env = BallVecEnv(n, window=spec["window"], config=bench.env_config(spec), seed=0, device="cuda:0",
                             max_episode_steps=0)
            os.environ["BALLENV_DEBUG_SKIP"] = "0"
            env.reset()
            a = torch.randint(0, 9, (T, n), device="cuda:0")
            out = env.alloc_rollout(T, keep_all_obs=True)
            for _ in range(2):
                env.step_many(a, keep_all_obs=True, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                env.step_many(a, keep_all_obs=True, out=out)
            e1.record()
            torch.cuda.synchronize()
            print("%s n=%7d skip=%2d  %.2f us/launch" % (wl, n, skip, e0.elapsed_time(e1) * 1e3 / (3 * T)), flush=True)
            env.close()
