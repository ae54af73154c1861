"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/ballenv.h declares,
the ctypes mirrors have the C layout, and the no-GPU paths fail loudly (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ballenv.h")


@pytest.fixture(scope="module")
def L():
    from gym_ballenv_b200 import _lib
    return _lib


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ballenv_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(L):
    names = _declared_functions()
    assert len(names) >= 17
    lib = C.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libballenv_b200.so does not export %s" % n
    assert sorted(L.EXPORTS) == names, "the ctypes binding and include/ballenv.h disagree on the entry points"


def test_abi_version_and_constants(L):
    src = open(HEADER).read()
    assert L.LIB.ballenv_abi_version() == L.ABI_VERSION == int(re.search(r"#define BALLENV_ABI_VERSION (\d+)", src).group(1))
    for cname, val in (("BALLENV_MAX_DYNAMIC", L.MAX_DYNAMIC), ("BALLENV_MAX_GOALS", L.MAX_GOALS),
                       ("BALLENV_MAX_STATIC", L.MAX_STATIC), ("BALLENV_MAX_WINDOW", L.MAX_WINDOW),
                       ("BALLENV_NUM_STATS", L.NUM_STATS), ("BALLENV_OBS_BITS", L.OBS_BITS),
                       ("BALLENV_ACT_XY_F64", L.ACT_XY_F64), ("BALLENV_FLAG_HIT_DYNAMIC", L.FLAG_HIT_DYNAMIC)):
        assert int(re.search(r"#define %s (\d+)" % cname, src).group(1)) == val, cname


def test_struct_layouts_match_the_header(L, tmp_path):
    """Compile a tiny C program against the header (plain gcc, no CUDA) and compare sizeof / offsetof."""
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "ballenv.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(BallenvConfig), offsetof(BallenvConfig, static_penalty),
         offsetof(BallenvConfig, obstacle_speed), offsetof(BallenvConfig, obs_goal_y), sizeof(BallenvStatePtrs));
  printf("%zu %zu %zu\n", offsetof(BallenvStatePtrs, static_stride), offsetof(BallenvStatePtrs, agent_x),
         offsetof(BallenvStatePtrs, error_flags));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(BallenvPolicyMLP), offsetof(BallenvPolicyMLP, fc1_weight),
         offsetof(BallenvPolicyMLP, value_bias), sizeof(BallenvA2CUpdate), offsetof(BallenvA2CUpdate, fc1_weight_grad),
         offsetof(BallenvA2CUpdate, policy_out));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = [int(v) for v in out]
    cfg, ptr = L.BallenvConfig, L.BallenvStatePtrs
    pol, upd = L.BallenvPolicyMLP, L.BallenvA2CUpdate
    want = [C.sizeof(cfg), cfg.static_penalty.offset, cfg.obstacle_speed.offset, cfg.obs_goal_y.offset, C.sizeof(ptr),
            ptr.static_stride.offset, ptr.agent_x.offset, ptr.error_flags.offset,
            C.sizeof(pol), pol.fc1_weight.offset, pol.value_bias.offset, C.sizeof(upd), upd.fc1_weight_grad.offset,
            upd.policy_out.offset]
    assert got == want


def test_config_default_and_validation(L):
    cfg = L.BallenvConfig()
    assert L.LIB.ballenv_config_default(C.byref(cfg), L.RULESET_GYM) == 0
    assert (cfg.static_obstacles, cfg.dynamic_obstacles, cfg.window, cfg.max_episode_steps) == (13, 5, 5, 1000)
    assert [cfg.obs_goal_x[i] for i in range(5)] == [12, 123, 87, 430, 230]      # examples/ball_cnn_ac3.py:45
    assert L.LIB.ballenv_state_bytes(C.byref(cfg), 4096) > 4096 * 4 * (4 + 2 * 16 + 3 * 8)
    # a single distinct goal would make np.random.randint(0) raise in the reference (ballenv_env.py:351-352)
    bad = L.BallenvConfig()
    L.LIB.ballenv_config_default(C.byref(bad), L.RULESET_GYM)
    for i in range(5):
        bad.obs_goal_x[i], bad.obs_goal_y[i] = 7.0, 7.0
    assert L.LIB.ballenv_state_bytes(C.byref(bad), 16) == -1
    assert b"distinct" in L.LIB.ballenv_last_error()
    bad2 = L.BallenvConfig()
    L.LIB.ballenv_config_default(C.byref(bad2), L.RULESET_GYM)
    bad2.window = 33
    assert L.LIB.ballenv_state_bytes(C.byref(bad2), 16) == -1
    bad3 = L.BallenvConfig()
    L.LIB.ballenv_config_default(C.byref(bad3), L.RULESET_PYGAME)
    bad3.dynamic_obstacles = 1          # createBoard's dynamic branch is broken in the reference (ballenv_pygame.py:502-506)
    assert L.LIB.ballenv_state_bytes(C.byref(bad3), 16) == -1
    with pytest.raises(L.BallenvError):
        L.check(L.LIB.ballenv_state_bytes(C.byref(bad3), 16))


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of falling back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gym_ballenv_b200 import BallVecEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BallVecEnv(8)
    from gym_ballenv_b200 import _lib as L
    cfg = L.BallenvConfig()
    L.LIB.ballenv_config_default(C.byref(cfg), L.RULESET_GYM)
    h = C.c_void_p()
    rc = L.LIB.ballenv_create(C.byref(cfg), 8, 0, 0, C.c_uint64(0), None, C.byref(h))
    assert rc < 0 and not h.value


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under gym_ballenv_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "gym_ballenv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle." not in text.replace("oracle/", "") or f.endswith((".cuh", ".cu")), f
