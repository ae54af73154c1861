"""One graphed A2C iteration (config 5) between cudaProfilerStart / Stop: `ncu --profile-from-start off --metrics gpu__time_duration.sum` lists its kernels."""
import os, sys, torch
sys.path.insert(0, ".")
from gym_ballenv_b200 import BallVecEnv
from gym_ballenv_b200.a2c import GraphedTrainer, Policy
env = BallVecEnv(16384, window=5, seed=0, device="cuda:0"); env.reset()
torch.manual_seed(0)
tr = GraphedTrainer(env, Policy(5).to("cuda:0"), 32)
tr.step(); tr.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", env.error_flags())
