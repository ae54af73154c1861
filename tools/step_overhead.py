"""Host-side cost of BallVecEnv.step(): wall clock per call against the device time of the launch it enqueues."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_ballenv_b200 import BallVecEnv
for n, w in ((4096, 5), (65536, 5), (65536, 10)):
    env = BallVecEnv(n, window=w, seed=0, device="cuda:0")
    env.reset()
    a = torch.randint(0, 9, (n,), device="cuda:0")
    for _ in range(200):
        env.step(a)
    torch.cuda.synchronize()
    K = 3000
    t0 = time.perf_counter()
    for _ in range(K):
        env.step(a)
    t1 = time.perf_counter()          # enqueue rate (the queue may run ahead of the device)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    obs = torch.empty((n, env.obs_row), device="cuda:0"); rew = torch.empty(n, device="cuda:0"); done = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    t3 = time.perf_counter()
    for _ in range(K):
        env.step_into(a, obs, rew, done)
    t4 = time.perf_counter()
    torch.cuda.synchronize()
    print("n=%6d W=%2d: step() %.1f us per call on the host, %.1f us per call end to end; step_into() %.1f us on the host" %
          (n, w, (t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6, (t4 - t3) / K * 1e6))
    env.close()
