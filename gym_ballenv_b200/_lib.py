"""ctypes binding of the C ABI declared in include/ballenv.h (libballenv_b200.so).

There is no CPU fallback: if the CUDA library is missing this module raises at import.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BALLENV_LIB_PATH") or os.path.join(HERE, "libballenv_b200.so")   # override: A/B experiments

ABI_VERSION = 5
MAX_DYNAMIC = 64
MAX_GOALS = 64
MAX_STATIC = 1024
MAX_WINDOW = 32
NUM_STATS = 16

RULESET_GYM, RULESET_PYGAME = 0, 1
F32, F64 = 0, 1
OBS_F32, OBS_U8, OBS_BITS = 0, 1, 2
ACT_INDEX_I64, ACT_INDEX_I32, ACT_INDEX_U8, ACT_XY_F32, ACT_XY_F64 = 0, 1, 2, 3, 4
FLAG_GOAL, FLAG_HIT, FLAG_TRUNCATED, FLAG_HIT_DYNAMIC = 1, 2, 4, 8
DEVERR_BAD_ACTION, DEVERR_TAPE_EXHAUSTED, DEVERR_RESET_STUCK = 1, 2, 4
KERNEL_GENERIC, KERNEL_ROLES, KERNEL_LEAN, KERNEL_LEAN2 = 0, 1, 2, 3
STAT_NAMES = ("episodes", "return_sum", "length_sum", "goals", "hits_static", "hits_dynamic", "timeouts", "steps")


class BallenvConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("ruleset", C.c_int32), ("window", C.c_int32),
        ("static_obstacles", C.c_int32), ("dynamic_obstacles", C.c_int32), ("n_goals", C.c_int32),
        ("time_step_for_change", C.c_int32), ("rd_th_obs", C.c_int32), ("max_episode_steps", C.c_int32),
        ("auto_reset", C.c_int32), ("precision", C.c_int32), ("obs_format", C.c_int32),
        ("static_penalty", C.c_double), ("dynamic_penalty", C.c_double),
        ("agent_radius", C.c_double), ("static_obstacle_radius", C.c_double),
        ("obstacle_speed", C.c_double * MAX_DYNAMIC),
        ("obs_goal_x", C.c_double * MAX_GOALS), ("obs_goal_y", C.c_double * MAX_GOALS),
    ]


class BallenvStatePtrs(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int64), ("n_stride", C.c_int64), ("static_stride", C.c_int64),
        ("dynamic_stride", C.c_int64), ("real_bytes", C.c_int32), ("obs_row_elems", C.c_int32),
        ("agent_x", C.c_void_p), ("agent_y", C.c_void_p), ("goal_x", C.c_void_p), ("goal_y", C.c_void_p),
        ("dist", C.c_void_p), ("total_distance", C.c_void_p), ("acc_reward", C.c_void_p),
        ("ep_len", C.c_void_p), ("episode", C.c_void_p), ("tick", C.c_void_p),
        ("static_x", C.c_void_p), ("static_y", C.c_void_p), ("dynamic_x", C.c_void_p), ("dynamic_y", C.c_void_p),
        ("dynamic_meta", C.c_void_p), ("flags", C.c_void_p), ("stats", C.c_void_p), ("error_flags", C.c_void_p),
    ]


class BallenvPolicyMLP(C.Structure):
    """include/ballenv.h: BallenvPolicyMLP (ballenv_rollout_policy)."""
    _fields_ = [("n_inputs", C.c_int32), ("hidden", C.c_int32), ("greedy", C.c_int32), ("reserved", C.c_int32),
                ("fc1_weight", C.c_void_p), ("fc1_bias", C.c_void_p), ("action_weight", C.c_void_p),
                ("action_bias", C.c_void_p), ("value_weight", C.c_void_p), ("value_bias", C.c_void_p)]


class BallenvA2CUpdate(C.Structure):
    """include/ballenv.h: BallenvA2CUpdate (ballenv_a2c_grads)."""
    _fields_ = [("n_inputs", C.c_int32), ("hidden", C.c_int32)] + [
        (n, C.c_void_p) for n in ("fc1_weight", "fc1_bias", "action_weight", "action_bias", "value_weight", "value_bias",
                                  "fc1_weight_grad", "fc1_bias_grad", "action_weight_grad", "action_bias_grad",
                                  "value_weight_grad", "value_bias_grad", "loss", "returns_stats", "policy_out")]


EXPORTS = (
    "ballenv_abi_version", "ballenv_last_error", "ballenv_config_default", "ballenv_state_bytes",
    "ballenv_create", "ballenv_destroy", "ballenv_state_ptrs", "ballenv_reset", "ballenv_step",
    "ballenv_step_many", "ballenv_observe", "ballenv_observe_features", "ballenv_observe_blocks", "ballenv_step_host", "ballenv_set_draw_tape", "ballenv_stats",
    "ballenv_stats_reset", "ballenv_error_flags", "ballenv_launch_count", "ballenv_selftest", "ballenv_kernel_variant",
    "ballenv_step_many_host", "ballenv_reset_fixed", "ballenv_state_written", "ballenv_observe_patches",
    "ballenv_rollout_policy", "ballenv_discounted_returns", "ballenv_a2c_workspace_bytes", "ballenv_a2c_grads",
)


def _bind(lib):
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    cfgp = C.POINTER(BallenvConfig)
    lib.ballenv_abi_version.restype = C.c_int
    lib.ballenv_last_error.restype = C.c_char_p
    lib.ballenv_config_default.argtypes = [cfgp, C.c_int]
    lib.ballenv_state_bytes.argtypes = [cfgp, i64]
    lib.ballenv_state_bytes.restype = i64
    lib.ballenv_create.argtypes = [cfgp, i64, i64, C.c_int, u64, vp, C.POINTER(vp)]
    lib.ballenv_destroy.argtypes = [vp]
    lib.ballenv_state_ptrs.argtypes = [vp, C.POINTER(BallenvStatePtrs)]
    lib.ballenv_state_written.argtypes = [vp, vp]
    lib.ballenv_reset.argtypes = [vp, vp, vp, vp]
    lib.ballenv_reset_fixed.argtypes = [vp, vp, C.c_double, C.c_double, vp, vp]
    lib.ballenv_step.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
    lib.ballenv_step_many.argtypes = [vp, vp, C.c_int, i32, vp, i32, vp, vp, vp]
    lib.ballenv_observe.argtypes = [vp, vp, vp]
    lib.ballenv_observe_features.argtypes = [vp, vp, vp]
    lib.ballenv_observe_blocks.argtypes = [vp, vp, vp]
    lib.ballenv_observe_patches.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.ballenv_rollout_policy.argtypes = [vp, C.POINTER(BallenvPolicyMLP), i32, vp, vp, vp, vp, vp, vp, vp]
    lib.ballenv_discounted_returns.argtypes = [vp, vp, vp, C.c_float, i32, i64, vp, vp]
    lib.ballenv_a2c_workspace_bytes.argtypes = [i32, i32, i64]
    lib.ballenv_a2c_workspace_bytes.restype = i64
    lib.ballenv_a2c_grads.argtypes = [C.POINTER(BallenvA2CUpdate), vp, vp, vp, i64, vp, i64, vp]
    lib.ballenv_step_host.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
    lib.ballenv_step_many_host.argtypes = [vp, vp, C.c_int, i32, vp, vp, vp, vp]
    lib.ballenv_set_draw_tape.argtypes = [vp, vp, i64, vp, i64, i32]
    lib.ballenv_stats.argtypes = [vp, vp, vp]
    lib.ballenv_stats_reset.argtypes = [vp, vp]
    lib.ballenv_error_flags.argtypes = [vp, vp, vp]
    lib.ballenv_launch_count.argtypes = [vp]
    lib.ballenv_launch_count.restype = i64
    lib.ballenv_kernel_variant.argtypes = [vp, C.c_int, i32]
    lib.ballenv_kernel_variant.restype = C.c_int
    lib.ballenv_selftest.argtypes = [C.c_int, i64, C.c_int, C.POINTER(i64)]
    lib.ballenv_selftest.restype = C.c_int
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("ballenv_abi_version",):
            fn.restype = C.c_int
    return lib


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gym_ballenv_b200: %s is missing - build it with `python __graft_entry__.py` or `python gym_ballenv_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = _bind(C.CDLL(LIB_PATH))
    if lib.ballenv_abi_version() != ABI_VERSION:
        raise ImportError("libballenv_b200.so ABI %d != binding ABI %d" % (lib.ballenv_abi_version(), ABI_VERSION))
    return lib


LIB = load()


class BallenvError(RuntimeError):
    pass


def check(rc):
    if rc < 0:
        raise BallenvError("ballenv error %d: %s" % (rc, LIB.ballenv_last_error().decode()))
    return rc
