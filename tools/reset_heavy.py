"""A reset-heavy rollout (time limit 10 steps, desynchronised): every block resets some environment every step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_ballenv_b200 import BallVecEnv
T, n = 20, 65536
spec = bench.workload_spec("c3")
env = BallVecEnv(n, window=10, config=bench.env_config(spec), seed=0, device="cuda:0", max_episode_steps=10)
env.reset()
env.set_state(ep_len=torch.randint(0, 10, (n,), device="cuda:0", dtype=torch.int32))
a = torch.randint(0, 9, (T, n), device="cuda:0")
out = env.alloc_rollout(T, keep_all_obs=True)
for _ in range(3):
    env.step_many(a, keep_all_obs=True, out=out)
torch.cuda.synchronize()
print("ok", env.error_flags())
