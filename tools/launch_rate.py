"""Launch-to-launch time of ballenv_step_many vs number of environments (is the loop launch-bound?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gym_ballenv_b200 import BallVecEnv

T = 200
for wl in ("c3", "w5"):
    spec = bench.workload_spec(wl)
    for n in (32, 2048, 16384, 65536, 262144):
        env = BallVecEnv(n, window=spec["window"], config=bench.env_config(spec), seed=0, device="cuda:0")
        env.reset()
        a = torch.randint(0, 9, (T, n), device="cuda:0")
        out = env.alloc_rollout(T, keep_all_obs=True)
        for _ in range(3):
            env.step_many(a, keep_all_obs=True, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(5):
            env.step_many(a, keep_all_obs=True, out=out)
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print("%s n=%7d  device %.2f us/launch   host-issue %.2f us/launch" %
              (wl, n, e0.elapsed_time(e1) * 1e3 / (5 * T), t_host * 1e6 / (5 * T)), flush=True)
        env.close()
