"""Per-env-step time of the rollout kernel as a function of how often environments reset (time limit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_ballenv_b200 import BallVecEnv
T, n = 200, 65536
spec = bench.workload_spec("c3")
for limit in (0, 1000, 200, 50, 10):
    os.environ["BALLENV_DEBUG_SKIP"] = os.environ.get("SKIP", "0")
    env = BallVecEnv(n, window=10, config=bench.env_config(spec), seed=0, device="cuda:0", max_episode_steps=limit)
    env.reset()
    # desynchronise the time limits: random episode lengths so far
    if limit:
        env.set_state(ep_len=torch.randint(0, limit, (n,), device="cuda:0", dtype=torch.int32))
    a = torch.randint(0, 9, (T, n), device="cuda:0")
    out = env.alloc_rollout(T, keep_all_obs=True)
    for _ in range(2):
        env.step_many(a, keep_all_obs=True, out=out)
    env.reset_stats()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        env.step_many(a, keep_all_obs=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (4 * T)
    st = env.stats()
    print("limit %5d: %.2f us per env-step of all envs, %.1f resets per step (%.2f%% of blocks if spread)" %
          (limit, us, st["episodes"] / (4 * T), 100 * min(1.0, st["episodes"] / (4 * T) / 2048)), flush=True)
    env.close()
