"""Device time of ballenv_observe_patches (40 x 40 rgb patches) for the reference's default obstacle set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_ballenv_b200 import BallVecEnv
for n in (16384, 65536):
    env = BallVecEnv(n, window=5, device="cuda:0")
    env.reset()
    a = torch.randint(0, 9, (100, n), device="cuda:0")
    env.step_many(a)      # a state off the reset distribution
    for dt in (torch.float32, torch.uint8):
        out = env.rgb_patches(dtype=dt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            env.rgb_patches(dtype=dt, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print("n=%d %s: %.1f us per launch, %.1f M patches/s, %.0f GB/s written" %
              (n, dt, us, n / us, out.numel() * out.element_size() / us / 1e3))
    env.close()
