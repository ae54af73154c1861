"""Config 5 (16 384 envs, W = 5, reference defaults, Policy(5) in the loop): the graphed torch loop against the fused
launch (FusedRollout), rollout only and whole iterations (rollout + batched update)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_ballenv_b200 import BallVecEnv
from gym_ballenv_b200.a2c import FusedRollout, GraphedRollout, Policy, a2c_loss
N, T = int(os.environ.get("N", "16384")), 32
for lanes in ("1", "2"):
    os.environ["BALLENV_LEAN_G"] = lanes
    for cls in (FusedRollout,) if lanes == "1" else (FusedRollout, GraphedRollout):
        env = BallVecEnv(N, window=5, seed=0, device="cuda:0")
        env.reset()
        torch.manual_seed(0)
        pol = Policy(5).to("cuda:0")
        opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
        roll = cls(env, pol, T)
        for _ in range(3):
            roll.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            roll.run()
        e1.record()
        torch.cuda.synchronize()
        ms_roll = e0.elapsed_time(e1) / 10
        def iteration():
            raw = roll.run()
            batch = roll.evaluate(raw)
            with torch.no_grad():
                _, v_last = pol(raw["obs"][T])
            loss = a2c_loss(batch, 0.99, bootstrap=v_last.squeeze(-1))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        for _ in range(2):
            iteration()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            iteration()
        e1.record()
        torch.cuda.synchronize()
        ms_it = e0.elapsed_time(e1) / 10
        print("%-15s lanes=%s n=%d: rollout of %d steps %.3f ms (%.1f us per step, %.1f M env-steps/s); iteration %.3f ms (%.1f M env-steps/s) errs=%d" %
              (cls.__name__, lanes, N, T, ms_roll, ms_roll * 1e3 / T, N * T / ms_roll / 1e3, ms_it, N * T / ms_it / 1e3, env.error_flags()), flush=True)
        env.close()

from gym_ballenv_b200.a2c import GraphedTrainer
os.environ["BALLENV_LEAN_G"] = "2"
env = BallVecEnv(N, window=5, seed=0, device="cuda:0")
env.reset()
torch.manual_seed(0)
pol = Policy(5).to("cuda:0")
FUSED = os.environ.get("FUSED_UPDATE", "1") == "1"
tr = GraphedTrainer(env, pol, T, fused_update=FUSED)
tr.step()
tr.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    loss = tr.step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(("GraphedTrainer fused_update=%s " % FUSED) + "lanes=2 n=%d: iteration %.3f ms (%.1f M env-steps/s) loss %.4f errs=%d" % (N, ms, N * T / ms / 1e3, float(loss), env.error_flags()))
env.close()
