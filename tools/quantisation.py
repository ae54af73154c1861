"""Per-env-step time of the rollout kernel vs number of environments around the resident-block boundaries
(592 resident blocks x 32 envs = 18 944 envs per round)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_ballenv_b200 import BallVecEnv
T = 200
spec = bench.workload_spec("c3")
NS = [int(x) for x in os.environ.get("NS", "18944,37888,56832,65536,75776,131072").split(",")]
for n in NS:
    env = BallVecEnv(n, window=10, config=bench.env_config(spec), seed=0, device="cuda:0")
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
    print("n=%7d (%.2f rounds): %.2f us per step, %.3f ns per env-step, %.2f G env-steps/s" %
          (n, n / 18944.0, us, us * 1e3 / n, n / us / 1e3), flush=True)
    env.close()
