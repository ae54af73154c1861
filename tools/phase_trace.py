"""Phase timeline of one block of the rollout kernel (library built with -DBALLENV_TRACE): clock64 stamps that
block 0 leaves in the reward rows of environments 32.., averaged over the steps of one launch.
  stamps: scalar 0 loop top, 1 before arrive(agent), 2 before sync(near), 3 after sync(near), 4 before arrive(done)
          static thread 8.. / first dynamic thread 16..: +0 loop top, +1 before sync(agent), +2 after, +3 before
          sync(near), +4 after, +5 before sync(done), +6 after"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gym_ballenv_b200 import BallVecEnv
T = 60
spec = bench.workload_spec(os.environ.get("WL", "c3"))
for n in [int(x) for x in os.environ.get("NS", "4736,65536").split(",")]:
    os.environ["BALLENV_DEBUG_SKIP"] = os.environ.get("SKIP", "0")   # read when the handle is created
    env = BallVecEnv(n, window=spec["window"], config=bench.env_config(spec), seed=0, device="cuda:0", max_episode_steps=0)
    os.environ["BALLENV_DEBUG_SKIP"] = "0"
    env.reset()
    a = torch.randint(0, 9, (T, n), device="cuda:0")
    out = env.alloc_rollout(T, keep_all_obs=True)
    for _ in range(2):
        res = env.step_many(a, keep_all_obs=True, out=out)
    torch.cuda.synchronize()
    rew = res[1].cpu().numpy()[:, 32:64].astype(np.int64)   # [T, 32] stamps (24 bits of clock64)
    def d(a_, b_, dt=0):   # mean over steps of stamp b (of step t + dt) - stamp a (of step t), modulo 2^24
        x = (rew[dt:, b_] - rew[:T - dt, a_]) & 0xffffff
        x = x[5:-2]
        return float(np.median(x))
    print("n=%d  step period (scalar loop top to loop top): %.0f cycles" % (n, d(0, 0, 1)))
    print("  scalar : top->arriveA %.0f | ->syncN(before) %.0f | wait N %.0f | ->arriveD %.0f | ->next top %.0f" %
          (d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(4, 0, 1)))
    for name, b in (("static ", 8), ("dynamic", 16)):
        print("  %s: top->A %.0f | wait A %.0f | A->N(before) %.0f | wait N %.0f | N->D(before) %.0f | wait D %.0f | ->next top %.0f" %
              (name, d(b, b + 1), d(b + 1, b + 2), d(b + 2, b + 3), d(b + 3, b + 4), d(b + 4, b + 5), d(b + 5, b + 6), d(b + 6, b, 1)))
    env.close()
