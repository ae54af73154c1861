"""40 x 40 rgb patches of the pixel policies (examples/ball_cnn_reinforce.py:120-163).

The rendering half restates the gym viewer (parity unpinned: pyglet / OpenGL cannot run here - oracle/patches.py says
what is restated); the resize half is pinned to Pillow bit for bit.  CPU tests: the oracle against Pillow and against
the committed fixture.  GPU tests: ballenv_observe_patches against the oracle and the fixture, bit-exact on the uint8
values; the float32 form is value / 255 exactly (ToTensor)."""
import numpy as np
import pytest

from helpers import load_golden
from oracle import patches as P


def _states(rng, n, ks, kd):
    agent = rng.integers(0, 501, (n, 2))
    goal = np.where(rng.random((n, 1)) < 0.4, agent + rng.integers(-60, 61, (n, 2)), rng.integers(0, 500, (n, 2)))
    stat = agent[:, None, :] + rng.integers(-80, 81, (n, ks, 2))
    dyn = agent[:, None, :] + rng.integers(-80, 81, (n, kd, 2))
    return agent, goal, stat, dyn


# ---------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("interp", ["bicubic", "bilinear"])
def test_resize_restatement_is_pillow(interp):
    from PIL import Image
    pil = {"bicubic": Image.BICUBIC, "bilinear": Image.BILINEAR}[interp]
    rng = np.random.default_rng(1)
    for size_in, size_out in ((100, 40), (64, 40), (128, 32), (100, 7), (20, 40)):
        for kind in range(3):
            if kind == 0:
                img = rng.integers(0, 256, (size_in, size_in, 3), dtype=np.uint8)
            elif kind == 1:
                img = np.where(rng.random((size_in, size_in, 3)) < 0.5, 255, 0).astype(np.uint8)   # hard edges: the ringing clips
            else:
                a, g, s, d = _states(rng, 1, 4, 4)
                img = P.codes_to_rgb(P.crop(P.render_frame(a[0], g[0], s[0], d[0]), a[0], size_in if size_in % 2 == 0 else 100))
            ref = np.asarray(Image.fromarray(img, "RGB").resize((size_out, size_out), pil))
            assert np.array_equal(P.resize_u8(img, size_out, interp), ref), (size_in, size_out, kind)


def test_sprites_are_the_viewer_polygons():
    sp = P.sprites()
    ob = sp["obstacle"]
    assert len(ob) == 40 and all(m < (1 << 40) for m in ob)
    # the 30-gon is symmetric about both axes (vertices at multiples of 12 degrees) and a little smaller than the disc
    assert ob == ob[::-1]
    assert all(m == int(format(m, "040b")[::-1], 2) for m in ob)
    area = sum(bin(m).count("1") for m in ob)
    assert 0.985 * np.pi * 400 < area < np.pi * 400
    # the self-intersecting goal quad as a triangle fan: everything but the bottom wedge
    goal = [format(m, "010b")[::-1] for m in sp["goal"]]
    assert goal[9] == "1111111111" and goal[5] == "1111111111"
    assert goal[0] == "1000000001" and goal[3] == "1111001111"


def test_oracle_matches_the_fixture():
    z, meta = load_golden("patches_kat")
    for i in range(meta["n"]):
        for interp in ("bicubic", "bilinear"):
            got = P.extract_patch_u8(z["agent"][i], z["goal"][i], z["stat"][i], z["dyn"][i], interp=interp)
            assert np.array_equal(got, z[interp][i]), (i, interp)


def test_draw_order_and_padding():
    # a static obstacle drawn over the agent, a moving one over the static one; the padding outside the world stays white
    codes = P.crop(P.render_frame((3, 3), (400, 400), [(10, 3)], [(40, 8)]), (3, 3), 100)
    # patch row i <-> frame row y = ay + 49 - i, patch column j <-> frame column c = ax - 50 + j
    assert codes[49, 50] == P.RED                # the agent's own pixel (3, 3): under the static disc
    assert codes[49, 50 + 24] == P.GREEN         # (27, 3): inside both discs, the moving one is drawn later
    assert codes[44, 50 + 50 - 1] == P.GREEN     # (52, 8): the moving disc alone
    assert (codes[:, :47] == P.WHITE).all() and (codes[53:, :] == P.WHITE).all()   # x < 0 and y < 0: padding
    speed0 = P.render_frame((3, 3), (400, 400), [], [(100, 100)], dynamic_speeds=[0])
    assert speed0[499 - 100, 100] == P.RED       # obstacle.speed == 0 -> red (ballenv_env.py:298-301)


# ---------------------------------------------------------------------------------------------------- GPU
def _env(n, ks, kd, parity=False, speeds=None):
    import torch
    from gym_ballenv_b200 import BallVecEnv, EnvConfig
    goals = ["%d,%d" % (40 + 37 * i, 30 + 11 * i) for i in range(max(kd, 2))]
    cfg = EnvConfig(static_obstacles=ks, dynamic_obstacles=kd, obstacle_speed=speeds or [1.0] * kd, obs_goal_position=goals)
    env = BallVecEnv(n, window=5, config=cfg, seed=3, device="cuda:0", parity=parity)
    env.reset()
    return env, torch


def _inject(env, torch, agent, goal, stat, dyn):
    f = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
    env.set_state(agent_x=f(agent[:, 0]), agent_y=f(agent[:, 1]), goal_x=f(goal[:, 0]), goal_y=f(goal[:, 1]),
                  static_x=f(stat[:, :, 0].T), static_y=f(stat[:, :, 1].T), dynamic_x=f(dyn[:, :, 0].T), dynamic_y=f(dyn[:, :, 1].T))


@pytest.mark.gpu
@pytest.mark.parametrize("parity", [False, True])
def test_gpu_patches_match_the_fixture(parity):
    z, meta = load_golden("patches_kat")
    env, torch = _env(meta["n"], meta["KS"], meta["KD"], parity)
    _inject(env, torch, z["agent"], z["goal"], z["stat"], z["dyn"])
    for interp in ("bicubic", "bilinear"):
        u8 = env.rgb_patches(interp=interp, dtype=torch.uint8).cpu().numpy()
        assert np.array_equal(u8, z[interp]), interp
        f32 = env.rgb_patches(interp=interp).cpu().numpy()
        assert f32.dtype == np.float32 and np.array_equal(f32, z[interp].astype(np.float32) / np.float32(255))
    assert env.error_flags() == 0
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ks,kd,width,size", [(13, 5, 100, 40), (8, 24, 100, 40), (3, 2, 64, 32), (1, 1, 128, 64), (40, 3, 100, 40)])
def test_gpu_patches_match_the_oracle(ks, kd, width, size):
    rng = np.random.default_rng(ks * 100 + kd)
    n = 40
    speeds = [float(j % 3) for j in range(kd)]     # speed 0: drawn red
    env, torch = _env(n, ks, kd, speeds=speeds)
    agent, goal, stat, dyn = _states(rng, n, ks, kd)
    agent[0] = (0, 0); agent[1] = (500, 500); agent[2] = (0, 500); agent[3] = (499, 1)
    stat[4, 0] = agent[4] + (3, -2); dyn[4, 0] = agent[4] + (-4, 6)       # everything on top of the agent
    dyn[5, :] = (-30, 520)                                                 # far outside the world
    _inject(env, torch, agent, goal, stat, dyn)
    for interp in ("bicubic", "bilinear"):
        got = env.rgb_patches(width=width, size=size, interp=interp, dtype=torch.uint8).cpu().numpy()
        for i in range(n):
            want = P.extract_patch_u8(agent[i], goal[i], stat[i], dyn[i], width, size, interp, dynamic_speeds=speeds)
            assert np.array_equal(got[i], want), (i, interp)
    env.close()


@pytest.mark.gpu
def test_gpu_patches_follow_the_rollout():
    """Patches of the live state after resets and steps (not injected): the state the kernels themselves produced."""
    env, torch = _env(256, 13, 5)
    g = torch.Generator(device="cuda:0").manual_seed(5)
    for _ in range(30):
        env.step(torch.randint(0, 9, (256,), device="cuda:0", generator=g))
    st = env.get_state()
    got = env.rgb_patches(dtype=torch.uint8).cpu().numpy()
    ax, ay = st["agent_x"].cpu().numpy(), st["agent_y"].cpu().numpy()
    gx, gy = st["goal_x"].cpu().numpy(), st["goal_y"].cpu().numpy()
    sx, sy = st["static_x"].cpu().numpy(), st["static_y"].cpu().numpy()
    dx, dy = st["dynamic_x"].cpu().numpy(), st["dynamic_y"].cpu().numpy()
    for i in range(0, 256, 9):
        want = P.extract_patch_u8((ax[i], ay[i]), (gx[i], gy[i]), list(zip(sx[:, i], sy[:, i])), list(zip(dx[:, i], dy[:, i])))
        assert np.array_equal(got[i], want), i
    env.close()


@pytest.mark.gpu
def test_gpu_patches_reject_bad_arguments():
    from gym_ballenv_b200._lib import BallenvError
    env, torch = _env(4, 2, 2)
    for kw in (dict(width=101), dict(width=130), dict(size=65), dict(width=128, size=1)):
        with pytest.raises(BallenvError):
            env.rgb_patches(**kw)
    with pytest.raises(ValueError):
        env.rgb_patches(interp="nearest")
    env.close()


@pytest.mark.gpu
def test_facade_extract_patch_and_the_cnn_policy():
    """extract_patch(state) of the single-environment facade (current state and an arbitrary state list) and one forward
    pass of PolicyCNN over rgb_patches + the goal-quadrant floats, as examples/ball_cnn_reinforce.py:360 does."""
    import torch
    from gym_ballenv_b200 import BallEnv
    from gym_ballenv_b200.a2c import PolicyCNN
    env = BallEnv(device="cuda:0", seed=1, window=5)
    state = env.reset()
    for _ in range(3):
        state, _, _, _ = env.step((1, 1))
    ks, kd = env._cfg.static_obstacles, env._cfg.dynamic_obstacles
    want = P.extract_patch(state[0], state[1], list(state[3:3 + ks]), list(state[3 + ks:]))
    got = env.extract_patch()
    assert tuple(got.shape) == (1, 3, 40, 40) and np.array_equal(got[0].cpu().numpy(), want)
    assert np.array_equal(env.extract_patch(list(state))[0].cpu().numpy(), want)
    other = [(100, 100), (300, 300), 0.0, (110, 95), (90, 130)]
    got2 = env.extract_patch(other, interp="bilinear")[0].cpu().numpy()
    assert np.array_equal(got2, P.extract_patch(other[0], other[1], other[3:], [], interp="bilinear"))
    policy = PolicyCNN().to("cuda:0").eval()
    with torch.no_grad():
        probs, value = policy(got, env.window_observation()[:, :4])
    assert tuple(probs.shape) == (1, 9) and tuple(value.shape) == (1, 1) and abs(float(probs.sum()) - 1.0) < 1e-5
    env.close()
