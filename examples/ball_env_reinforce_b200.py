"""The reference's older scripts on the GPU environment:

    python examples/ball_env_reinforce_b200.py --envs 4096 --iterations 20
        REINFORCE on the 29-float block-count observation (examples/ball_env_reinforce.py), N environments at once
    python examples/ball_env_reinforce_b200.py --states State_info_trail_no2 --actions Trial_no_2 --epochs 50
        first fit the policy to a demonstration log (examples/train_supervise.py), then evaluate it on the GPU env
    python examples/ball_env_reinforce_b200.py --resume stored_models/supervised/episode_9999.pth --iterations 0
        evaluate a checkpoint of the reference
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gym_ballenv_b200 as gb  # noqa: E402
from gym_ballenv_b200.legacy import BlockPolicy, reinforce_loss, rollout_blocks, train_supervised  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--envs', type=int, default=4096)
    p.add_argument('--steps', type=int, default=32, help='steps per update')
    p.add_argument('--iterations', type=int, default=20)
    p.add_argument('--gamma', type=float, default=0.99)
    p.add_argument('--seed', type=int, default=543)
    p.add_argument('--states', type=str, default=None, help='demonstration log: pickled 29-float vectors')
    p.add_argument('--actions', type=str, default=None, help='demonstration log: pickled [+-10, +-10] actions')
    p.add_argument('--epochs', type=int, default=50)
    p.add_argument('--resume', type=str, default=None, help='state_dict with affine1/2/3 (reference or this script)')
    a = p.parse_args()
    torch.manual_seed(a.seed)
    dev = torch.device('cuda:0')
    supervised = a.states is not None
    policy = BlockPolicy(logits=supervised).to(dev)
    if a.resume:
        sd = torch.load(a.resume, map_location=dev)
        policy.logits = 'affine3.weight' in sd and supervised or policy.logits
        policy.load_state_dict(sd)
    if supervised:
        x, y = gb.pathlogs.load_path_log(a.states, a.actions, device=dev)
        loss = train_supervised(policy, x, y, a.epochs)
        acc = (policy(x).argmax(-1) == y).float().mean().item()
        print('supervised: %d samples, %d epochs, loss %.4f, accuracy %.3f' % (len(x), a.epochs, loss, acc))
    env = gb.BallVecEnv(a.envs, window=5, seed=a.seed)
    env.reset()
    opt = torch.optim.Adam(policy.parameters(), lr=1e-2)       # ball_env_reinforce.py:197
    for it in range(a.iterations):
        t0 = time.time()
        batch = rollout_blocks(env, policy, a.steps)
        loss = reinforce_loss(batch, a.gamma)
        opt.zero_grad()
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
        if it % 5 == 0:
            print('iter %4d  loss %12.4f  mean reward %.5f  (%.0f k env-steps/s)' %
                  (it, loss.item(), batch['reward'].mean().item(), a.envs * a.steps / (time.time() - t0) / 1e3))
    env.stats_reset() if hasattr(env, 'stats_reset') else None
    with torch.no_grad():
        rollout_blocks(env, policy, 200, greedy=True)
    st = env.stats()
    print('greedy evaluation, 200 steps x %d envs: %s' % (a.envs, {k: st[k] for k in ('episodes', 'goals', 'hits_static', 'hits_dynamic')}))


if __name__ == '__main__':
    main()
