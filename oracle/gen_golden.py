"""Generate the golden fixtures under tests/golden/ by running the reference's
own, unedited code through oracle/ref_shim.py.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.gen_golden            # writes tests/golden/*.npz

The fixtures are what travels to the GPU box (the reference tree does not).
Each .npz carries a JSON ``meta`` string with the config, the draw mode and the
reference commit-independent description of how it was made.

Fixture families
  rollout_philox_*.npz  N envs x T steps with the vector wrapper's auto-reset /
                        time-limit semantics wrapped around the reference env;
                        every draw answered from Philox4x32-10 at its address
                        (oracle/draws.py) - the production draw stream.
  rollout_mt_*.npz      same, but the reference draws from np.random (MT19937,
                        np.random.seed(s)) and the recorded draws are stored as
                        slot-addressed tapes (parity / tape mode).
  edge_gym.npz          hand-built single-step cases on injected states.
  window_kat.npz        RNG-free window-observation known answers (W=5,10,21).
  blocks_kat.npz        legacy 29-float block-count observation of the REINFORCE scripts, 200 states.
  patches_kat.npz       40 x 40 rgb patches of the pixel policies: the reference's patch path cannot run (pyglet /
                        OpenGL viewer, Python-2 slicing), so the frame is the restatement of oracle/patches.py and
                        the pad / crop / resize is done here with the installed Pillow itself (BICUBIC and BILINEAR).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from . import draws as D
from . import ref_shim as R
from .ballenv_oracle import AGENT_MOVES

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CFG_DEFAULT = dict(static_obstacles=13, dynamic_obstacles=5, obstacle_speed=[1, 1, 1, 1, 1],
                   obs_goal_position=['12,122', '123,93', '87,150', '430,440', '230,11'],
                   time_step_for_change=50, rd_th_obs=60, static_penalty=[1, 1], dynamic_penalty=[4000, 8000])

# config 3 of BASELINE.json ("dense moving obstacles"), SURVEY.md section 8d
_LATTICE = ['%d,%d' % (x, y) for y in (100, 200, 300, 400) for x in (50, 130, 210, 290, 370, 450)]
CFG_DENSE = dict(static_obstacles=8, dynamic_obstacles=24, obstacle_speed=[1] * 24,
                 obs_goal_position=_LATTICE, time_step_for_change=50, rd_th_obs=60,
                 static_penalty=[1, 1], dynamic_penalty=[4000, 8000])

# small, fast-changing config: exercises goal changes, axis moves and truncation often
CFG_BUSY = dict(static_obstacles=3, dynamic_obstacles=4, obstacle_speed=[1, 2, 1, 3],
                obs_goal_position=['40,30', '250,5', '250,60', '460,20'],
                time_step_for_change=7, rd_th_obs=45, static_penalty=[1, 1], dynamic_penalty=[4000, 8000])


# repeated obstacle goals (indices 0 and 2 are the same point): the goal-change step draws among the goals whose
# POSITION differs from the current one (ballenv_env.py:351-352), duplicates included as separate list entries
CFG_DUPGOALS = dict(static_obstacles=2, dynamic_obstacles=3, obstacle_speed=[1, 2, 1],
                    obs_goal_position=['40,30', '250,5', '40,30', '460,20'],
                    time_step_for_change=5, rd_th_obs=50, static_penalty=[1, 1], dynamic_penalty=[4000, 8000])


def _pack_rows(obs, window):
    """float obs [4 + W*W] -> (quadrant index, W row bitmasks)."""
    q = int(np.argmax(obs[:4]))
    assert obs[:4].sum() == 1.0
    grid = obs[4:].reshape(window, window)
    rows = [int(sum(int(grid[r, c]) << c for c in range(window))) for r in range(window)]
    return q, rows


def _action_for(policy, src, g, t, ref):
    if callable(policy):      # the reference's own network in the loop (greedy), fed by its own prep_state4
        return policy(ref)
    w = src.action_word(g, t)
    if policy == "random":
        return D.mulhi(w, 9)
    # goal seeking with 25 % random moves
    if D.mulhi(w, 4) == 0:
        return D.mulhi(D.mullo(w, 4), 9)
    (ax, ay), (gx, gy) = ref.env.state[0], ref.env.state[1]
    sx = (gx > ax) - (gx < ax)
    sy = (gy > ay) - (gy < ay)
    return AGENT_MOVES.index((sx, sy))


def rollout(cfg, n_envs, T, seed, mode, max_steps, windows=(5, 10), g0=0, policies=("random", "seek")):
    args = R.default_args(**cfg)
    ks, kd = cfg["static_obstacles"], cfg["dynamic_obstacles"]
    K = ks + kd
    act_src = D.PhiloxDraws(seed ^ 0x5EED)
    if mode == "philox":
        src = D.PhiloxDraws(seed)
        refs = [R.ReferenceEnv(args, R.AddressedRouter(src, g0 + i)) for i in range(n_envs)]
    else:
        np.random.seed(seed)
        refs = [R.ReferenceEnv(args, R.MTRouter()) for i in range(n_envs)]

    rec = dict(
        actions=np.zeros((T, n_envs), np.uint8),
        agent=np.zeros((T, n_envs, 2), np.float64), goal=np.zeros((T, n_envs, 2), np.float64),
        dist=np.zeros((T, n_envs), np.float64), total_distance=np.zeros((T, n_envs), np.float64),
        acc=np.zeros((T, n_envs), np.float64),
        reward=np.zeros((T, n_envs), np.float64), done=np.zeros((T, n_envs), np.uint8),
        flags=np.zeros((T, n_envs), np.uint8), ep_len=np.zeros((T, n_envs), np.int32),
        obst=np.zeros((T, n_envs, K, 2), np.float64),
        dyn_goal=np.zeros((T, n_envs, kd), np.uint8), dyn_counter=np.zeros((T, n_envs, kd), np.int32),
        quadrant=np.zeros((T, n_envs), np.uint8),
    )
    for w in windows:
        rec["rows%d" % w] = np.zeros((T, n_envs, w), np.uint32)
    init = dict(agent=np.zeros((n_envs, 2)), goal=np.zeros((n_envs, 2)), dist=np.zeros(n_envs),
                total_distance=np.zeros(n_envs), obst=np.zeros((n_envs, K, 2)))
    for w in windows:
        init["rows%d" % w] = np.zeros((n_envs, w), np.uint32)
    init["quadrant"] = np.zeros(n_envs, np.uint8)

    # tape recording (mt mode): per env list of per-episode reset draws and per-step draws
    A = 8
    step_val = np.zeros((T, n_envs, kd, 2), np.uint8)
    step_n = np.zeros((T, n_envs, kd, 2), np.uint8)
    reset_logs = [[] for _ in range(n_envs)]   # env -> list of episodes -> list of (lo, hi, v)

    def take_reset_log(i):
        if mode != "mt":
            return
        log = refs[i].router.log
        reset_logs[i].append([(lo, hi, v) for (c, lo, hi, v) in log if c == "reset"])
        del log[:]

    def take_step_log(i, t):
        if mode != "mt":
            return
        log = refs[i].router.log
        j = -1
        second = False
        for (c, lo, hi, v) in log:
            assert c == "step" and lo == 0
            if second:
                step_val[t, i, j, 1], step_n[t, i, j, 1] = v, hi
                second = False
                continue
            j += 1
            step_val[t, i, j, 0], step_n[t, i, j, 0] = v, hi
            if hi == 100 and v >= cfg["rd_th_obs"]:
                second = True
        assert j == kd - 1 and not second, (j, kd)
        del log[:]

    ep_len = [0] * n_envs
    stats = dict(episodes=0, return_sum=0.0, length_sum=0, goals=0, hits_static=0, hits_dynamic=0,
                 timeouts=0, steps=0)
    for i, ref in enumerate(refs):
        if mode == "mt":
            del ref.router.log[:]
        s = ref.reset()
        take_reset_log(i)
        init["agent"][i], init["goal"][i], init["dist"][i] = s[0], s[1], s[2]
        init["total_distance"][i] = ref.env.total_distance
        init["obst"][i] = np.array([tuple(p) for p in s[3:]], dtype=np.float64).reshape(K, 2)
        for w in windows:
            q, rows = _pack_rows(ref.observe(s, w), w)
            init["rows%d" % w][i] = rows
            init["quadrant"][i] = q

    for t in range(T):
        for i, ref in enumerate(refs):
            a = _action_for(policies[i % len(policies)], act_src, g0 + i, t, ref)
            rec["actions"][t, i] = a
            s, r, d, _ = ref.step(AGENT_MOVES[a])
            take_step_log(i, t)
            ep_len[i] += 1
            env = ref.env
            goal_flag = s[2] < env.threshold_goal
            hit_idx = next((k for k, p in enumerate(s[3:]) if env.check_overlap(s[0], p)), -1)
            hit = hit_idx >= 0
            assert bool(d) == (goal_flag or hit)
            trunc = max_steps > 0 and ep_len[i] >= max_steps
            done = bool(d) or trunc
            f = (1 if goal_flag else 0) | (2 if hit else 0) | (4 if trunc else 0) | (8 if hit_idx >= ks else 0)
            rec["reward"][t, i], rec["done"][t, i], rec["flags"][t, i] = r, done, f
            stats["steps"] += 1
            if done:
                stats["episodes"] += 1
                stats["return_sum"] += env.total_reward_accumulated
                stats["length_sum"] += ep_len[i]
                stats["goals"] += int(goal_flag)
                stats["hits_static"] += int(hit and hit_idx < ks)
                stats["hits_dynamic"] += int(hit_idx >= ks)
                stats["timeouts"] += int(trunc and not d)
                s = ref.reset()
                take_reset_log(i)
                ep_len[i] = 0
            rec["agent"][t, i], rec["goal"][t, i], rec["dist"][t, i] = s[0], s[1], s[2]
            rec["total_distance"][t, i] = env.total_distance
            rec["acc"][t, i] = env.total_reward_accumulated
            rec["ep_len"][t, i] = ep_len[i]
            rec["obst"][t, i] = np.array([tuple(p) for p in s[3:]], dtype=np.float64).reshape(K, 2)
            for j, o in enumerate(env.obstacle_list[ks:]):
                rec["dyn_goal"][t, i, j] = env.obstacle_goal_list.index(o.curr_goal)
                rec["dyn_counter"][t, i, j] = o.curr_counter
            for w in windows:
                q, rows = _pack_rows(ref.observe(s, w), w)
                rec["rows%d" % w][t, i] = rows
                rec["quadrant"][t, i] = q

    out = {"rec_" + k: v for k, v in rec.items()}
    out.update({"init_" + k: v for k, v in init.items()})
    if mode == "mt":
        E = max(len(x) for x in reset_logs)
        reset_val = np.zeros((E, n_envs, D.reset_tape_width(ks, kd, A)), np.uint16)
        for i, eps in enumerate(reset_logs):
            for e, log in enumerate(eps):
                vals = [v - lo for (lo, hi, v) in log]
                reset_val[e, i, 0:4] = vals[0:4]
                pairs = [(vals[k], vals[k + 1]) for k in range(4, len(vals), 2)]
                # replay acceptance to slot the pairs: static i attempts, then dynamic
                gx, gy = vals[0], 480 + vals[1]
                ax, ay = vals[2], vals[3]
                p = 0
                for si in range(ks):
                    a = 0
                    while True:
                        x, y = pairs[p][0], 20 + pairs[p][1]
                        assert a < A, "raise A"
                        reset_val[e, i, 4 + (si * A + a) * 2: 4 + (si * A + a) * 2 + 2] = pairs[p]
                        p += 1
                        a += 1
                        rej = (abs(x - ax) < 25 and abs(y - ay) < 15) or (abs(x - gx) < 25 and abs(y - gy) < 15)
                        if not rej:
                            break
                for dj in range(kd):
                    reset_val[e, i, 4 + 2 * A * ks + 2 * dj: 4 + 2 * A * ks + 2 * dj + 2] = pairs[p]
                    p += 1
                assert p == len(pairs)
        out.update(tape_step_val=step_val, tape_step_n=step_n, tape_reset_val=reset_val)
    meta = dict(kind="rollout", ruleset="gym", mode=mode, seed=seed, n_envs=n_envs, T=T, g0=g0,
                max_episode_steps=max_steps, auto_reset=True, windows=list(windows), cfg=cfg,
                tape_attempts=A, stats=stats, action_seed=seed ^ 0x5EED,
                policies=[p if isinstance(p, str) else "checkpoint" for p in policies],
                made_by="oracle/gen_golden.py running /root/reference unedited via oracle/ref_shim.py")
    out["meta"] = np.array(json.dumps(meta))
    return out


def pygame_action(src, g, t):
    """Raw (dx, dy) of the pygame-ruleset fixtures: multiples of 0.75 in [-3, 3] (non-integral on purpose)."""
    w = src.action_word(g, t)
    return (D.mulhi(w, 9) - 4) * 0.75, (D.mulhi(D.mullo(w, 9), 9) - 4) * 0.75


def rollout_pygame(n_static, n_envs, T, seed, g0=0):
    """createBoard (ballenv_pygame.py:314-706) stepped with raw float actions; a finished episode is followed by
    reset(), like the vector wrapper's auto-reset.  Recorded state is the one after that reset."""
    src = D.PhiloxDraws(seed)
    act_src = D.PhiloxDraws(seed ^ 0x5EED)
    refs = [R.ReferencePygameEnv(n_static, R.PygameRouter(src, g0 + i)) for i in range(n_envs)]
    K = n_static
    rec = dict(actions=np.zeros((T, n_envs, 2)), agent=np.zeros((T, n_envs, 2)), goal=np.zeros((T, n_envs, 2)),
               dist=np.zeros((T, n_envs)), total_distance=np.zeros((T, n_envs)), acc=np.zeros((T, n_envs)),
               reward=np.zeros((T, n_envs)), done=np.zeros((T, n_envs), np.uint8),
               obst=np.zeros((T, n_envs, K, 2)))
    init = dict(agent=np.zeros((n_envs, 2)), goal=np.zeros((n_envs, 2)), dist=np.zeros(n_envs),
                total_distance=np.zeros(n_envs), obst=np.zeros((n_envs, K, 2)))

    def snap(dst, idx, ref):
        st = ref.env.state
        dst["agent"][idx] = st[0]
        dst["goal"][idx] = st[1]
        dst["dist"][idx] = st[2]
        dst["total_distance"][idx] = ref.env.total_distance
        for k in range(K):
            dst["obst"][idx + (k,)] = st[3 + k][:2]

    episodes = 0
    for i, ref in enumerate(refs):
        ref.reset()
        snap(init, (i,), ref)
    for t in range(T):
        for i, ref in enumerate(refs):
            a = pygame_action(act_src, g0 + i, t)
            rec["actions"][t, i] = a
            _, reward, done, _ = ref.step(a)
            rec["reward"][t, i], rec["done"][t, i] = reward, done
            rec["acc"][t, i] = ref.env.total_reward_accumulated
            if done:
                episodes += 1
                ref.reset()
            snap(rec, (t, i), ref)
    out = {"rec_" + k: v for k, v in rec.items()}
    out.update({"init_" + k: v for k, v in init.items()})
    out["meta"] = json.dumps(dict(ruleset="pygame", n_static=n_static, n_envs=n_envs, T=T, seed=seed, g0=g0,
                                  episodes=episodes, agent_radius=10, static_obstacle_radius=10))
    return out


def reset_fixed_kat(n_static=9, n_envs=48, seed=33, g0=9, rounds=3, steps=6):
    """createBoard.resetFixedstate (ballenv_pygame.py:589-624) on the running reference: reset(), a few raw-action steps,
    then per round resetFixedstate() and a few more steps.  Recorded after every resetFixedstate: agent, goal, state[2],
    total_distance, accumulated reward, the (kept) obstacles, the number of agent draws it took; and the rewards / dones
    of the steps that follow (they depend on old_dist and total_distance)."""
    src = D.PhiloxDraws(seed)
    act_src = D.PhiloxDraws(seed ^ 0x5EED)
    refs = [R.ReferencePygameEnv(n_static, R.PygameRouter(src, g0 + i)) for i in range(n_envs)]
    rec = dict(agent=np.zeros((rounds, n_envs, 2)), goal=np.zeros((rounds, n_envs, 2)), dist=np.zeros((rounds, n_envs)),
               total_distance=np.zeros((rounds, n_envs)), acc=np.zeros((rounds, n_envs)),
               obst=np.zeros((rounds, n_envs, n_static, 2)), draws=np.zeros((rounds, n_envs), np.int32),
               actions=np.zeros((rounds, steps, n_envs, 2)), reward=np.zeros((rounds, steps, n_envs)),
               done=np.zeros((rounds, steps, n_envs), np.uint8))
    pre_actions = np.zeros((steps, n_envs, 2))
    t = 0
    for i, ref in enumerate(refs):
        ref.reset()
    for k in range(steps):                       # episodes under way, accumulated reward non-zero, before the first call
        for i, ref in enumerate(refs):
            a = pygame_action(act_src, g0 + i, t)
            pre_actions[k, i] = a
            _, _, done, _ = ref.step(a)
            assert not done or True
        t += 1
    for r in range(rounds):
        for i, ref in enumerate(refs):
            n0 = len(ref.router.trace)
            ref.reset_fixed()
            st = ref.env.state
            rec["agent"][r, i], rec["goal"][r, i], rec["dist"][r, i] = st[0], st[1], st[2]
            rec["total_distance"][r, i] = ref.env.total_distance
            rec["acc"][r, i] = ref.env.total_reward_accumulated
            rec["draws"][r, i] = (len(ref.router.trace) - n0) // 2
            for j in range(n_static):
                rec["obst"][r, i, j] = st[3 + j][:2]
        for k in range(steps):
            for i, ref in enumerate(refs):
                a = pygame_action(act_src, g0 + i, t)
                rec["actions"][r, k, i] = a
                _, reward, done, _ = ref.step(a)
                rec["reward"][r, k, i], rec["done"][r, k, i] = reward, done
            t += 1
    out = {"rec_" + k: v for k, v in rec.items()}
    out["pre_actions"] = pre_actions
    out["meta"] = np.array(json.dumps(dict(kind="reset_fixed_kat", ruleset="pygame", n_static=n_static, n_envs=n_envs,
                                           seed=seed, g0=g0, rounds=rounds, steps=steps, goal=[145, 120],
                                           made_by="oracle/gen_golden.py running createBoard.resetFixedstate of "
                                                   "/root/reference/ballenv_pygame.py unedited via oracle/ref_shim.py")))
    return out


def features_kats():
    """States -> the 20 floats of the reference's featureExtractor helpers (numpy part, featureExtractor.py:36-257)."""
    rng = np.random.RandomState(12)
    agents, goals, obsts, feats, rads = [], [], [], [], []
    K = 6
    for i in range(160):
        world = 100.0 if i % 2 == 0 else 500.0
        agent = tuple(rng.uniform(0, world, 2)) if i % 4 else tuple(rng.randint(0, int(world), 2).astype(float))
        goal = tuple(rng.uniform(0, world, 2))
        if i == 7:
            goal = agent                       # zero goal vector: arccos(0) -> the "left" sector
        if i == 9:
            goal = (agent[0], agent[1] + 30.0)  # straight up
        if i == 11:
            goal = (agent[0] - 10.0, agent[1] - 40.0)
        ob = [(float(x), float(y)) for x, y in rng.randint(0, int(world), (K, 2))]
        if i % 5 == 0:
            ob[0] = (agent[0] + 31.0, agent[1])        # surface distance ~1: social force near its threshold
            ob[1] = (agent[0], agent[1] - 45.0)
        rad = 10 if i % 3 else 5
        agents.append(agent); goals.append(goal); obsts.append(ob); rads.append(rad)
        feats.append(R.reference_features(agent, goal, ob, agent_rad=rad))
    return dict(agent=np.array(agents), goal=np.array(goals), obst=np.array(obsts), agent_rad=np.array(rads, np.float64),
                features=np.array(feats), meta=json.dumps(dict(K=K, n=len(agents), obstacle_rad=20)))


def blocks_kats():
    """States -> the legacy 29-float block-count observation (examples/ball_env_reinforce.py:130-172, lifted)."""
    prep = R.load_legacy_block_state()
    rng = np.random.RandomState(29)
    K = 8
    agents, goals, obsts, outs = [], [], [], []
    for i in range(200):
        if i % 5 == 4:     # non-integral coordinates (the pygame ruleset moves by floats)
            agent = tuple(float(v) for v in rng.uniform(20, 480, 2).round(2))
            goal = tuple(float(v) for v in rng.uniform(0, 500, 2).round(2))
            ob = [tuple(float(v) for v in (np.array(agent) + rng.uniform(-75, 75, 2)).round(2)) for _ in range(K)]
        else:
            agent = tuple(int(v) for v in rng.randint(20, 480, 2))
            goal = tuple(int(v) for v in rng.randint(0, 500, 2))
            ob = [tuple(int(v) for v in (np.array(agent) + rng.randint(-75, 76, 2))) for _ in range(K)]
        if i % 4 == 0:
            ob[0] = (agent[0], agent[1] + 33)            # dx == 0: both blocks 0 whatever dy is
            ob[1] = (agent[0] - 47, agent[1])            # dy == 0
            ob[2] = (agent[0] - 10, agent[1] - 10)       # dx - 10 == 0 on the positive side
            ob[3] = (agent[0] + 10, agent[1] + 50)       # |dx| + 10 == 20: block 1; |dy| + 10 == 60: block 3, outside
            ob[4] = (agent[0] - 69, agent[1] - 70)       # (69 - 10) // 20 == 2 inside, (70 - 10) // 20 == 3 outside
        if i == 3:
            goal = agent
        agents.append(agent); goals.append(goal); obsts.append(ob)
        outs.append(np.asarray(prep([agent, goal, 0.0] + ob), np.float64))
    return dict(agent=np.array(agents, np.float64), goal=np.array(goals, np.float64), obst=np.array(obsts, np.float64),
                blocks=np.array(outs), meta=json.dumps(dict(K=K, n=len(agents),
                made_by="oracle/gen_golden.py: prep_state2 of examples/ball_env_reinforce.py:130-172, AST-lifted")))


def patches_kats():
    """States -> uint8 [3, 40, 40] patches: restated frame (oracle/patches.py), then the steps of extract_patch
    (examples/ball_cnn_reinforce.py:130-147) with numpy padding and Pillow's own resize."""
    from PIL import Image
    from . import patches as P
    rng = np.random.RandomState(40)
    KS, KD = 4, 4
    agents, goals, stat, dyn, out_c, out_l = [], [], [], [], [], []
    for i in range(48):
        agent = tuple(int(v) for v in rng.randint(0, 501, 2))
        goal = tuple(int(v) for v in (np.array(agent) + rng.randint(-60, 61, 2))) if i % 3 == 0 else tuple(int(v) for v in rng.randint(0, 500, 2))
        st = [tuple(int(v) for v in (np.array(agent) + rng.randint(-75, 76, 2))) for _ in range(KS)]
        dy = [tuple(int(v) for v in (np.array(agent) + rng.randint(-75, 76, 2))) for _ in range(KD)]
        if i == 0:
            agent = (0, 0)                      # the padding fills three quarters of the patch
        if i == 1:
            agent = (500, 500)
        if i == 2:
            st[0] = (agent[0] + 10, agent[1]); dy[0] = (agent[0] + 20, agent[1] + 5)   # overlapping discs: draw order
            goal = (agent[0] - 12, agent[1] + 9)
        if i == 3:
            agent = (250, 3); st[1] = (260, -15); dy[1] = (230, -30)                   # obstacles outside the world
        agents.append(agent); goals.append(goal); stat.append(st); dyn.append(dy)
        span = 50
        frame = P.codes_to_rgb(P.render_frame(agent, goal, st, dy))
        padded = np.pad(frame, ((span, span), (span, span), (0, 0)), mode="constant", constant_values=255)   # :137
        ax, ay = agent[0] + span, frame.shape[0] - agent[1] + span                                           # :133-135
        patch = padded[max(ay - span, 0):ay + span, ax - span:ax + span, :]                                  # :144
        full = np.full((100, 100, 3), 255, np.uint8)       # (agent_y = 500: the slice starts above the padded frame)
        full[100 - patch.shape[0]:, :patch.shape[1]] = patch
        img = Image.fromarray(full, "RGB")
        out_c.append(np.asarray(img.resize((40, 40), Image.BICUBIC)).transpose(2, 0, 1))
        out_l.append(np.asarray(img.resize((40, 40), Image.BILINEAR)).transpose(2, 0, 1))
    import PIL
    return dict(agent=np.array(agents, np.float64), goal=np.array(goals, np.float64), stat=np.array(stat, np.float64),
                dyn=np.array(dyn, np.float64), bicubic=np.array(out_c, np.uint8), bilinear=np.array(out_l, np.uint8),
                meta=json.dumps(dict(KS=KS, KD=KD, n=len(agents), pillow=PIL.__version__,
                made_by="oracle/gen_golden.py: oracle/patches.py frame + Pillow resize (see the module header)")))


def pathlog_kat():
    """The head of the shipped demonstration log (examples/State_info_trail_no2 + Trial_no_2, Python 2 pickles) and
    the labels examples/train_supervise.py:43-59 derives from its actions (restated here: that script is Python 2
    syntax and cannot be lifted): action / 10 as a tuple -> its index in move_list, 8 if it is not there."""
    import pickle
    ex = os.path.join(R.REFERENCE_ROOT, "examples")
    with open(os.path.join(ex, "State_info_trail_no2"), "rb") as f:
        X = pickle.load(f, encoding="latin1")
    with open(os.path.join(ex, "Trial_no_2"), "rb") as f:
        Y = pickle.load(f, encoding="latin1")
    move_list = [(1, 1), (1, -1), (1, 0), (0, 1), (0, -1), (0, 0), (-1, 1), (-1, 0), (-1, -1)]
    labels = []
    for a in Y:
        tup = (int(a[0] // 10), int(a[1] // 10))
        j = 8
        for k in range(9):
            if move_list[k] == tup:
                j = k
                break
        labels.append(j)
    n = 96
    return dict(states=np.asarray(X[:n], np.float64), actions=np.asarray(Y[:n], np.int64), labels=np.asarray(labels[:n], np.int64),
                meta=json.dumps(dict(n=n, total=len(X), label_histogram=np.bincount(labels, minlength=9).tolist())))


def a2c_kats():
    """Known answers of the reference's actor-critic update and of its shipped WINDOW = 5 checkpoint.

    * finish_episode (examples/ball_cnn_ac3.py:222-246), AST-lifted and run unedited on one recorded episode: the
      observations, greedy actions and rewards that went in, the loss it printed and the gradients it left on every
      parameter of ``Policy(5)`` (its optimizer stepping with lr = 0 so that they survive).
    * the state_dict the reference ships (stored_models/ball_state3/2layer+dropout+randpos/episode_2500.pth, the
      README's run) - 6 tensors, 21 KB: a data file of the reference, kept as a fixture so that the GPU box (which has
      no reference tree) can load the same network.
    """
    import torch
    ns = R.load_a2c(5, gamma=0.99)
    policy = ns["policy"]
    sd = torch.load(R.REFERENCE_CHECKPOINT_W5, map_location="cpu")
    policy.load_state_dict(sd)
    rng = np.random.RandomState(5)
    T = 24
    states = (rng.rand(T, 29) < 0.15).astype(np.float32)
    states[:, :4] = 0
    states[np.arange(T), rng.randint(0, 4, T)] = 1
    rewards = (rng.randn(T) * 0.01).astype(np.float64)
    rewards[-1] -= 1.0                                   # the episode ends on a static obstacle
    actions = np.zeros(T, np.int64)
    probs_rec = np.zeros((T, 9), np.float32)
    values_rec = np.zeros(T, np.float32)
    for t in range(T):
        probs, value = policy(torch.from_numpy(states[t:t + 1]))          # select_action (:210-220), greedy
        a = int(probs.argmax(-1))
        actions[t] = a
        probs_rec[t], values_rec[t] = probs.detach().numpy()[0], float(value)
        policy.saved_actions.append(ns["SavedAction"](torch.log(probs[0, a]), value))
        policy.rewards.append(float(rewards[t]))
    ns["finish_episode"]()
    loss = float(ns["printed"][-1][1])
    out = {"ckpt_" + k.replace(".", "_"): v.numpy() for k, v in sd.items()}
    out.update({"grad_" + k.replace(".", "_"): p.grad.numpy().copy() for k, p in policy.named_parameters()})
    out.update(states=states, actions=actions, rewards=rewards, probs=probs_rec, values=values_rec,
               loss=np.array(loss, np.float64))
    out["meta"] = np.array(json.dumps(dict(kind="a2c_kat", gamma=0.99, T=T, window=5,
                                           checkpoint="examples/stored_models/ball_state3/2layer+dropout+randpos/episode_2500.pth",
                                           made_by="oracle/gen_golden.py: Policy and finish_episode AST-lifted from "
                                                   "examples/ball_cnn_ac3.py, run unedited")))
    return out


def rollout_checkpoint():
    """The shipped WINDOW = 5 checkpoint driving the reference environment in closed loop: the reference's own
    ``Policy`` (AST-lifted) fed by its own ``prep_state4``, greedy actions, default obstacles, Philox-addressed draws
    (so the CUDA path can replay the episodes), TimeLimit 60, auto-reset - examples/ball_cnn_ac3.py:553-613 without
    the sampling noise."""
    import torch
    ns = R.load_a2c(5)
    policy = ns["policy"]
    policy.load_state_dict(torch.load(R.REFERENCE_CHECKPOINT_W5, map_location="cpu"))

    def greedy(ref):
        with torch.no_grad():
            probs, _ = policy(ref.prep_state4(ref.env.state, 5).float())
        return int(probs.argmax(-1))

    return compress_rollout(rollout(CFG_DEFAULT, 12, 180, 29, "philox", 60, windows=(5,), g0=77, policies=(greedy,)))


def compress_rollout(out):
    """Shrink dtypes where the values are integral (checked)."""
    for k in list(out):
        v = out[k]
        if isinstance(v, np.ndarray) and v.dtype == np.float64 and k.split("_", 1)[1] in ("agent", "goal", "obst"):
            assert np.all(v == np.round(v)) and np.abs(v).max() < 32000
            out[k] = v.astype(np.int16)
    return out


# --------------------------------------------------------------------------- edge cases
def edge_cases():
    """Single-step transitions on injected states (gym ruleset, 2 static + 3
    dynamic obstacles).  Every case: state, action, tape words -> next state,
    reward, done, flags, W=5 / W=10 observation."""
    cfg = dict(static_obstacles=2, dynamic_obstacles=3, obstacle_speed=[1, 2, 1],
               obs_goal_position=['100,100', '300,300', '50,400'], time_step_for_change=50, rd_th_obs=60,
               static_penalty=[1, 1], dynamic_penalty=[4000, 8000])
    args = R.default_args(**cfg)
    far = [(400, 100), (420, 300)]                 # statics far from everything
    dfar = [(300, 50), (310, 200), (320, 350)]      # dynamics far from everything
    W100 = lambda v: D.word_for(v, 100)
    W9 = lambda v: D.word_for(v, 9)
    W2 = lambda v: D.word_for(v, 2)
    cases = []

    def case(name, agent, goal, obst, action, words, dist=None, goal_idx=(0, 1, 2), counter=(0, 0, 0)):
        d = float(np.hypot(goal[0] - agent[0], goal[1] - agent[1])) if dist is None else dist
        cases.append(dict(name=name, agent=agent, goal=goal, obst=list(obst), action=action, words=words,
                          dist=d, goal_idx=list(goal_idx), counter=list(counter)))

    follow = [(W100(0), 0)] * 3      # every dynamic obstacle follows its goal direction
    # walls: agent at 0 / 500 pushed outwards
    case("wall_low", (0, 0), (250, 490), far + dfar, (-1, -1), follow)
    case("wall_high", (500, 500), (250, 490), far + dfar, (1, 1), follow)
    case("wall_mixed", (500, 0), (250, 490), far + dfar, (1, -1), follow)
    case("big_action_clamped", (495, 3), (250, 490), far + dfar, (10, -10), follow)
    # overlap boundary: distance exactly 25 after the move (15,20) and just outside (26)
    case("hit_exact_25_static", (99, 100), (250, 490), [(115, 120), far[1]] + dfar, (1, 0), follow)
    case("miss_26_static", (99, 100), (250, 490), [(126, 100), far[1]] + dfar, (0, 0), follow)
    case("hit_25_axis", (99, 100), (250, 490), [(125, 100), far[1]] + dfar, (1, 0), follow)
    # dynamic obstacle moves INTO the agent this step (post-move positions are tested, :262-278)
    case("dynamic_moves_into_agent", (200, 200), (250, 490), far + [(216, 221), dfar[1], dfar[2]], (0, 0), follow)
    # first hit in list order: static before dynamic -> static penalty
    case("static_and_dynamic_hit", (200, 200), (250, 490), [(210, 210), far[1]] + [(190, 190), dfar[1], dfar[2]],
         (0, 0), follow)
    # goal: distance exactly 10 is NOT goal (<), 9.x is
    case("goal_exact_10_not", (239, 490), (250, 490), far + dfar, (1, 0), follow)
    case("goal_inside", (240, 490), (250, 490), far + dfar, (1, 0), follow)
    case("goal_and_hit_same_step", (240, 490), (250, 490), [(260, 480), far[1]] + dfar, (1, 0), follow)
    # obstacle on an axis of its goal: single randint(9) draw, all nine table entries
    for k in range(9):
        case("axis_move_%d" % k, (10, 5), (250, 490), far + [(100, 50), dfar[1], dfar[2]], (0, 0),
             [(W9(k), 0), (W100(0), 0), (W100(0), 0)])
    # off-axis, u >= threshold: second draw picks the table entry (speed 2 obstacle)
    for k in range(9):
        case("random_move_%d" % k, (10, 5), (250, 490), far + dfar, (0, 0),
             [(W100(0), 0), (W100(60 + k), W9(k)), (W100(59), 0)])
    # goal-change step: counter reached -> no move, new goal among the others
    case("goal_change_pick0", (10, 5), (250, 490), far + dfar, (0, 0),
         [(W2(0), 0), (W2(1), 0), (W2(1), 0)], counter=(50, 50, 50))
    case("goal_change_mixed", (10, 5), (250, 490), far + dfar, (0, 0),
         [(W2(1), 0), (W100(10), 0), (W2(0), 0)], counter=(50, 49, 51), goal_idx=(2, 0, 1))
    # obstacle outside the world keeps moving (no clamp, :334-347)
    case("obstacle_outside_world", (10, 5), (250, 490), far + [(-30, 520), (510, -12), dfar[2]], (0, 0),
         [(W100(99), W9(7)), (W100(99), W9(1)), (W100(0), 0)])
    # obstacle straddling the window edge, agent near a wall (window pokes outside the world)
    case("window_outside_world", (1, 1), (250, 490), [(20, 20), far[1]] + dfar, (-1, -1), follow)
    case("window_partial", (100, 100), (250, 490), [(127, 100), (100, 128)] + dfar, (0, 0), follow)
    # stored dist that disagrees with the positions (old_dist is the stored value, :236)
    case("stale_dist", (100, 100), (250, 490), far + dfar, (1, 1), follow, dist=123.5)

    n = len(cases)
    ks, kd, K = 2, 3, 5
    out = dict(
        in_agent=np.zeros((n, 2)), in_goal=np.zeros((n, 2)), in_dist=np.zeros(n), in_total=np.zeros(n),
        in_obst=np.zeros((n, K, 2)), in_goal_idx=np.zeros((n, kd), np.uint8), in_counter=np.zeros((n, kd), np.int32),
        action=np.zeros((n, 2)), words=np.zeros((n, kd, 2), np.uint32),
        out_agent=np.zeros((n, 2)), out_dist=np.zeros(n), out_obst=np.zeros((n, K, 2)),
        out_goal_idx=np.zeros((n, kd), np.uint8), out_counter=np.zeros((n, kd), np.int32),
        out_reward=np.zeros(n), out_done=np.zeros(n, np.uint8), out_flags=np.zeros(n, np.uint8),
        out_acc=np.zeros(n), out_quadrant=np.zeros(n, np.uint8),
        out_rows5=np.zeros((n, 5), np.uint32), out_rows10=np.zeros((n, 10), np.uint32),
    )
    names = []
    for i, c in enumerate(cases):
        tape = np.array(c["words"], dtype=np.uint64).astype(np.uint32).reshape(1, 1, kd, 2)
        src = D.TapeDraws(step_tape=tape, n_static=ks, n_dynamic=kd)
        ref = R.ReferenceEnv(args, R.AddressedRouter(D.PhiloxDraws(99), 0))
        ref.reset()                                 # creates the obstacle objects
        ref.router.src = src
        ref.router.tick = 0
        total = 400.0
        ref.inject(c["agent"], c["goal"], c["dist"], c["obst"], c["goal_idx"], c["counter"],
                   total_distance=total, acc=0.25)
        s, r, d, _ = ref.step(c["action"])
        env = ref.env
        goal_flag = s[2] < env.threshold_goal
        hit_idx = next((k for k, p in enumerate(s[3:]) if env.check_overlap(s[0], p)), -1)
        assert bool(d) == (goal_flag or hit_idx >= 0), c["name"]
        f = (1 if goal_flag else 0) | (2 if hit_idx >= 0 else 0) | (8 if hit_idx >= ks else 0)
        names.append(c["name"])
        out["in_agent"][i], out["in_goal"][i], out["in_dist"][i], out["in_total"][i] = c["agent"], c["goal"], c["dist"], total
        out["in_obst"][i] = c["obst"]
        out["in_goal_idx"][i], out["in_counter"][i] = c["goal_idx"], c["counter"]
        out["action"][i], out["words"][i] = c["action"], tape[0, 0]
        out["out_agent"][i], out["out_dist"][i] = s[0], s[2]
        out["out_obst"][i] = np.array([tuple(p) for p in s[3:]], dtype=np.float64)
        for j, o in enumerate(env.obstacle_list[ks:]):
            out["out_goal_idx"][i, j] = env.obstacle_goal_list.index(o.curr_goal)
            out["out_counter"][i, j] = o.curr_counter
        out["out_reward"][i], out["out_done"][i], out["out_flags"][i] = r, d, f
        out["out_acc"][i] = env.total_reward_accumulated
        for w in (5, 10):
            q, rows = _pack_rows(ref.observe(s, w), w)
            out["out_rows%d" % w][i] = rows
            out["out_quadrant"][i] = q
    meta = dict(kind="edge", ruleset="gym", cfg=cfg, names=names, in_acc=0.25,
                made_by="oracle/gen_golden.py running /root/reference unedited via oracle/ref_shim.py")
    out["meta"] = np.array(json.dumps(meta))
    return out


def window_kats():
    """RNG-free window observations straight from the reference's prep_state4."""
    args = R.default_args(static_obstacles=0, dynamic_obstacles=0, obstacle_speed=[], obs_goal_position=[])
    ref = R.ReferenceEnv(args, R.MTRouter())
    rng = np.random.RandomState(12345)
    states = [[(82, 82), (400, 490), 0.0, (100, 100)]]          # SURVEY.md section 8c KAT
    for n in range(120):
        ax, ay = int(rng.randint(0, 501)), int(rng.randint(0, 501))
        k = int(rng.randint(0, 7))
        obst = []
        for _ in range(k):
            # cluster obstacles around the agent so windows are not all empty
            obst.append((ax + int(rng.randint(-45, 46)), ay + int(rng.randint(-45, 46))))
        states.append([(ax, ay), (int(rng.randint(0, 500)), int(rng.randint(0, 500))), 1.0] + obst)
    # goal exactly on an axis of the agent (quadrant tie-breaks, >= vs <)
    states.append([(50, 50), (50, 50), 0.0])
    states.append([(50, 50), (49, 50), 0.0])
    states.append([(50, 50), (50, 49), 0.0])
    states.append([(50, 50), (49, 49), 0.0])
    kmax = max(len(s) - 3 for s in states)
    n = len(states)
    out = dict(agent=np.zeros((n, 2), np.int32), goal=np.zeros((n, 2), np.int32),
               n_obst=np.zeros(n, np.int32), obst=np.full((n, kmax, 2), 10000, np.int32),
               quadrant=np.zeros(n, np.uint8))
    for w in (5, 10, 21):
        out["rows%d" % w] = np.zeros((n, w), np.uint32)
    for i, s in enumerate(states):
        out["agent"][i], out["goal"][i], out["n_obst"][i] = s[0], s[1], len(s) - 3
        for k, p in enumerate(s[3:]):
            out["obst"][i, k] = p
        for w in (5, 10, 21):
            q, rows = _pack_rows(ref.observe(s, w), w)
            out["rows%d" % w][i] = rows
            out["quadrant"][i] = q
    out["meta"] = np.array(json.dumps(dict(
        kind="window_kat", made_by="oracle/gen_golden.py: prep_state4 of examples/ball_cnn_ac3.py:384-412, AST-lifted")))
    return out


def main(argv):
    if not R.reference_available():
        print("reference tree missing; cannot generate", file=sys.stderr)
        return 1
    os.makedirs(OUT_DIR, exist_ok=True)
    jobs = {
        "window_kat": lambda: window_kats(),
        "edge_gym": lambda: edge_cases(),
        "rollout_philox_default": lambda: compress_rollout(rollout(CFG_DEFAULT, 24, 260, 7, "philox", 120, g0=1000)),
        "rollout_philox_busy": lambda: compress_rollout(rollout(CFG_BUSY, 16, 200, 11, "philox", 40, g0=5)),
        "rollout_philox_dupgoals": lambda: compress_rollout(rollout(CFG_DUPGOALS, 16, 160, 13, "philox", 40, g0=3)),
        "rollout_philox_dense": lambda: compress_rollout(rollout(CFG_DENSE, 8, 120, 3, "philox", 1000, windows=(10,))),
        "rollout_mt_default": lambda: compress_rollout(rollout(CFG_DEFAULT, 8, 150, 0, "mt", 60)),
        "a2c_kat": lambda: a2c_kats(),
        "rollout_checkpoint": lambda: rollout_checkpoint(),
        "features_kat": lambda: features_kats(),
        "blocks_kat": lambda: blocks_kats(),
        "pathlog_kat": lambda: pathlog_kat(),
        "rollout_pygame": lambda: rollout_pygame(9, 12, 200, 21, g0=40),
        "reset_fixed_kat": lambda: reset_fixed_kat(),
        "patches_kat": lambda: patches_kats(),
    }
    only = set(argv[1:])
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        out = fn()
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-28s %8.1f KiB" % (name, os.path.getsize(path) / 1024.0))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
