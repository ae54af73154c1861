#!/bin/bash
# usage: tools/ncu_extra.sh <tag>  - `ncu --set full` captures of the policy-in-the-loop rollout kernel (config 5: 16 384
# environments, 32 steps per launch) and of the rgb patch kernel (16 384 environments); each program runs plain first.
tag=$1
cat > gpurun_out/_pol.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from gym_ballenv_b200 import BallVecEnv
from gym_ballenv_b200.a2c import FusedRollout, Policy
env = BallVecEnv(16384, window=5, seed=0, device="cuda:0"); env.reset()
torch.manual_seed(0)
roll = FusedRollout(env, Policy(5).to("cuda:0"), 32)
for _ in range(6): roll.run()
torch.cuda.synchronize(); print("ok", env.error_flags())
PY
cat > gpurun_out/_patch.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from gym_ballenv_b200 import BallVecEnv
env = BallVecEnv(16384, window=5, seed=0, device="cuda:0"); env.reset()
env.step_many(torch.randint(0, 9, (100, 16384), device="cuda:0"))
out = env.rgb_patches()
for _ in range(5): env.rgb_patches(out=out)
torch.cuda.synchronize(); print("ok", env.error_flags())
PY
python gpurun_out/_pol.py > gpurun_out/plain_pol.log 2>&1 || { tail -5 gpurun_out/plain_pol.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:ballenv_lean_kernel -s 4 -c 1 -f -o gpurun_out/prof_${tag}_policy python gpurun_out/_pol.py > gpurun_out/ncu_pol.log 2>&1; tail -1 gpurun_out/ncu_pol.log
python gpurun_out/_patch.py > gpurun_out/plain_patch.log 2>&1 || { tail -5 gpurun_out/plain_patch.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:ballenv_patch_kernel -s 3 -c 1 -f -o gpurun_out/prof_${tag}_patch python gpurun_out/_patch.py > gpurun_out/ncu_patch.log 2>&1; tail -1 gpurun_out/ncu_patch.log
