/*
 * ballenv.h - C ABI of the B200-native batched gym-ballenv step()/reset()/observe hot path.
 *
 * This is the drop-in boundary: plain C, raw pointers and sizes, no C++/torch types.
 * The reference (ranok92/gym-ballenv) is pure Python and has no FFI; each entry point
 * below names the reference interface it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative BALLENV_E* code; the message for the
 *     last failure on the calling thread is ballenv_last_error().  Nothing throws.
 *   - nothing synchronises the device unless documented; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - "device pointer" arguments must live on the handle's device.
 *   - a handle is confined to one host thread / stream at a time; distinct handles are independent.
 *
 * State layout (struct-of-arrays, resident in HBM, owned by the handle or by a caller arena):
 *   per-env scalars are arrays of n_stride elements (one thread per environment reads them: a warp
 *   touches whole 128-byte lines); per-obstacle fields are environment-major rows [n_stride][K padded
 *   to a multiple of 4]: the rows of 32 consecutive environments are one contiguous slice per field, which
 *   the thread-per-environment kernels move as one bulk copy per warp and the block-of-roles kernels as
 *   consecutive 128-bit quads (one quad of four obstacles per thread).
 */
#ifndef BALLENV_H_
#define BALLENV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BALLENV_ABI_VERSION 5

#define BALLENV_MAX_DYNAMIC 64
#define BALLENV_MAX_GOALS 64
#define BALLENV_MAX_STATIC 1024
#define BALLENV_MAX_WINDOW 32

/* error codes */
#define BALLENV_OK 0
#define BALLENV_EINVAL (-1)   /* bad argument / config */
#define BALLENV_ECUDA (-2)    /* a CUDA call failed (see ballenv_last_error) */
#define BALLENV_ENOMEM (-3)
#define BALLENV_ESTATE (-4)   /* call not valid in the handle's current state (e.g. tape exhausted) */

/* BallenvConfig.ruleset */
#define BALLENV_RULESET_GYM 0     /* gym_ballenv/envs/ballenv_env.py : BallEnv (500x500, moving obstacles) */
#define BALLENV_RULESET_PYGAME 1  /* ballenv_pygame.py : createBoard (100x100, float coordinates) */

/* BallenvConfig.precision : type the positions are stored and compared in */
#define BALLENV_F32 0   /* production: fp32 positions, fp64 distance/reward arithmetic in registers (gym ruleset; the
                           pygame ruleset, whose coordinates are non-integral, always runs as BALLENV_F64:
                           BallenvStatePtrs.real_bytes tells) */
#define BALLENV_F64 1   /* parity mode: everything fp64, operation order of the reference */

/* BallenvConfig.obs_format : element type of the observation rows */
#define BALLENV_OBS_F32 0   /* float32 [n][4 + W*W]  - what prep_state4 returns (examples/ball_cnn_ac3.py:412) */
#define BALLENV_OBS_U8 1    /* uint8   [n][4 + W*W] */
#define BALLENV_OBS_BITS 2  /* uint32  [n][ceil((4 + W*W) / 32)], bit b of the row = element b */
/* (the 20-float feature vector of featureExtractor.py is a separate call: ballenv_observe_features) */

/* action_kind of ballenv_step */
#define BALLENV_ACT_INDEX_I64 0 /* int64 [n]  index into the agent move table of examples/ball_cnn_ac3.py:530 */
#define BALLENV_ACT_INDEX_I32 1
#define BALLENV_ACT_INDEX_U8 2
#define BALLENV_ACT_XY_F32 3    /* float32 [n][2] raw (dx, dy), what BallEnv.step(action) indexes (ballenv_env.py:247-248) */
#define BALLENV_ACT_XY_F64 4

/* bits of the per-env flags byte / of done */
#define BALLENV_FLAG_GOAL 1
#define BALLENV_FLAG_HIT 2
#define BALLENV_FLAG_TRUNCATED 4
#define BALLENV_FLAG_HIT_DYNAMIC 8   /* the first obstacle hit (list order) is a dynamic one */

/* device-side error bits (ballenv_error_flags) */
#define BALLENV_DEVERR_BAD_ACTION 1       /* action index outside [0, 9) : treated as (0, 0) */
#define BALLENV_DEVERR_TAPE_EXHAUSTED 2   /* a tape had no slot for a requested draw : Philox used instead */
#define BALLENV_DEVERR_RESET_STUCK 4      /* an obstacle placement was rejected 4096 times (the reference would spin) */

/* slots of the episode-statistics vector (doubles) */
#define BALLENV_STAT_EPISODES 0
#define BALLENV_STAT_RETURN_SUM 1
#define BALLENV_STAT_LENGTH_SUM 2
#define BALLENV_STAT_GOALS 3
#define BALLENV_STAT_HITS_STATIC 4
#define BALLENV_STAT_HITS_DYNAMIC 5
#define BALLENV_STAT_TIMEOUTS 6
#define BALLENV_STAT_STEPS 7
#define BALLENV_NUM_STATS 16

/*
 * Environment configuration; one config is shared by all environments of a handle.
 * Field names follow the argparse Namespace consumed by BallEnv.customize_environment
 * (gym_ballenv/envs/ballenv_env.py:87-109; defaults examples/ball_cnn_ac3.py:40-51).
 */
typedef struct BallenvConfig {
  int32_t abi_version;        /* BALLENV_ABI_VERSION */
  int32_t ruleset;            /* BALLENV_RULESET_* */
  int32_t window;             /* WINDOW of examples/ball_cnn_ac3.py:493 (1..32) */
  int32_t static_obstacles;   /* args.static_obstacles */
  int32_t dynamic_obstacles;  /* args.dynamic_obstacles (gym ruleset only) */
  int32_t n_goals;            /* len(args.obs_goal_position) ; >= dynamic_obstacles ; >= 2 distinct if any dynamic */
  int32_t time_step_for_change; /* args.time_step_for_change */
  int32_t rd_th_obs;          /* args.rd_th_obs */
  int32_t max_episode_steps;  /* TimeLimit of gym_ballenv/__init__.py:7 (1000) ; 0 = none */
  int32_t auto_reset;         /* 1: done envs are reset inside the step launch (vector wrapper) */
  int32_t precision;          /* BALLENV_F32 / BALLENV_F64 */
  int32_t obs_format;         /* BALLENV_OBS_* */
  double static_penalty;      /* args.static_penalty[1]  (ballenv_env.py:138,223) */
  double dynamic_penalty;     /* args.dynamic_penalty[1] (ballenv_env.py:158,223) */
  double agent_radius;            /* pygame ruleset ctor (ballenv_pygame.py:316) */
  double static_obstacle_radius;  /* pygame ruleset ctor */
  double obstacle_speed[BALLENV_MAX_DYNAMIC]; /* args.obstacle_speed */
  double obs_goal_x[BALLENV_MAX_GOALS];       /* args.obs_goal_position, parsed */
  double obs_goal_y[BALLENV_MAX_GOALS];
} BallenvConfig;

/* Device pointers of the SoA state (ballenv_state_ptrs).  Real = float or double per config.precision. */
typedef struct BallenvStatePtrs {
  int64_t n_envs;
  int64_t n_stride;     /* allocated environments per array (>= n_envs, multiple of 128) */
  int64_t static_stride;  /* elements per environment row of static_x / static_y  (static_obstacles rounded up to 4) */
  int64_t dynamic_stride; /* elements per environment row of dynamic_x / dynamic_y / dynamic_meta */
  int32_t real_bytes;   /* 4 or 8 */
  int32_t obs_row_elems; /* elements per observation row in the configured obs_format */
  void *agent_x, *agent_y;    /* Real [n_stride]            state[0] */
  void *goal_x, *goal_y;      /* Real [n_stride]            state[1] */
  double *dist;               /* [n_stride]                 state[2] (becomes old_dist of the next step) */
  double *total_distance;     /* [n_stride]                 self.total_distance */
  double *acc_reward;         /* [n_stride]                 self.total_reward_accumulated */
  int32_t *ep_len;            /* [n_stride]                 steps since reset (TimeLimit counter) */
  uint32_t *episode;          /* [n_stride]                 index of the current episode (reset-draw address) */
  uint32_t *tick;             /* [n_stride]                 steps since creation (step-draw address) */
  void *static_x, *static_y;  /* Real [n_stride][static_stride]    state[3 : 3 + Ks] of env e at row e */
  void *dynamic_x, *dynamic_y;/* Real [n_stride][dynamic_stride]   state[3 + Ks :] */
  uint32_t *dynamic_meta;     /* [n_stride][dynamic_stride]  curr_goal index | curr_counter << 8 */
  uint8_t *flags;             /* [n_stride]                 BALLENV_FLAG_* of the last step */
  double *stats;              /* [BALLENV_NUM_STATS]        episode statistics (all-reduce these across GPUs) */
  uint32_t *error_flags;      /* [1]                        BALLENV_DEVERR_* */
} BallenvStatePtrs;

typedef struct BallenvHandle BallenvHandle;
typedef void *ballenv_stream_t; /* cudaStream_t */

int ballenv_abi_version(void);
const char *ballenv_last_error(void);

/* Fill *cfg with the defaults of examples/ball_cnn_ac3.py:40-51 (gym) or createBoard() (pygame). */
int ballenv_config_default(BallenvConfig *cfg, int ruleset);

/* Bytes of device memory one handle needs for n_envs environments (to size a caller-owned arena). */
int64_t ballenv_state_bytes(const BallenvConfig *cfg, int64_t n_envs);

/*
 * Replaces: gym.make('gymball-v0') + env.unwrapped.customize_environment(args)
 *           (gym_ballenv/__init__.py:4-11, ballenv_env.py:43-109)  /  createBoard(...) (ballenv_pygame.py:316).
 * global_env_offset : id of environment 0 in the whole (multi-GPU) job; Philox streams are keyed by the
 *                     global id so trajectories do not depend on the sharding.
 * arena             : device memory of >= ballenv_state_bytes() bytes (256-byte aligned) or NULL to let the
 *                     library cudaMalloc it.
 */
int ballenv_create(const BallenvConfig *cfg, int64_t n_envs, int64_t global_env_offset, int device,
                   uint64_t seed, void *arena, BallenvHandle **out);
int ballenv_destroy(BallenvHandle *h);
int ballenv_state_ptrs(BallenvHandle *h, BallenvStatePtrs *out);
/*
 * Tell the library that the caller has written state through the pointers of ballenv_state_ptrs (state injection:
 * no reference counterpart - the reference's scripts assign env.state directly).  The production kernels take exact
 * shortcuts that hold for the integral coordinates the gym ruleset produces and skip the test for it while nothing
 * but resets and index-action steps has touched the state; this call re-validates the whole state on `stream` (one
 * small kernel, no synchronisation).  Writing non-integral coordinates WITHOUT calling it gives wrong observations.
 */
int ballenv_state_written(BallenvHandle *h, ballenv_stream_t stream);

/*
 * Replaces: BallEnv.reset() (ballenv_env.py:113-167) / createBoard.reset() (ballenv_pygame.py:460-513),
 * followed by prep_state4(state, WINDOW) (examples/ball_cnn_ac3.py:384-412).
 * mask : device uint8 [n] (non-zero = reset this env) or NULL for all.  obs_out : device rows or NULL.
 */
int ballenv_reset(BallenvHandle *h, const uint8_t *mask, void *obs_out, ballenv_stream_t stream);

/*
 * Replaces: createBoard.resetFixedstate() (ballenv_pygame.py:589-624; pygame ruleset only): a new episode with the
 * goal at a fixed point ((145, 120) in the reference) and the OBSTACLES KEPT - the agent is redrawn until it is at least
 * 50 from the goal and touches no obstacle; state[2] is the distance of the accepted attempt's first draw; the
 * accumulated reward restarts at 0.  mask / obs_out as for ballenv_reset.
 */
int ballenv_reset_fixed(BallenvHandle *h, const uint8_t *mask, double goal_x, double goal_y, void *obs_out,
                        ballenv_stream_t stream);

/*
 * Replaces: BallEnv.step(action) (ballenv_env.py:232-289) incl. move_obstacles (:323-353) and
 * calculate_reward (:200-229) / createBoard.step (ballenv_pygame.py:650-706), the TimeLimit wrapper,
 * and the prep_state4 call of the training loop (examples/ball_cnn_ac3.py:560) - one fused launch.
 * actions    : device, layout per action_kind
 * obs_out    : device rows in cfg.obs_format (post-reset observation for envs that finished), or NULL
 * reward_out : device float32 [n] (BALLENV_F32) or float64 [n] (BALLENV_F64), or NULL
 * done_out   : device uint8 [n] (goal | hit | truncated), or NULL
 */
int ballenv_step(BallenvHandle *h, const void *actions, int action_kind, void *obs_out, void *reward_out,
                 uint8_t *done_out, ballenv_stream_t stream);

/*
 * T consecutive steps in one call (open-loop rollouts: the actions of all T steps are known up front, i.e. the loop
 * of examples/ball_cnn_ac3.py:553-613 with pre-sampled actions).  actions [T][n], reward_out [T][n], done_out [T][n];
 * obs_out [T][n][row] if obs_all_steps != 0, else [n][row] holding the last step's observation.
 * With the production configuration (BALLENV_F32, gym ruleset, Philox draws, index actions, BALLENV_OBS_F32 rows)
 * this is ONE launch: every block keeps its environments on chip for all T steps, so only actions, observations,
 * rewards and dones touch device memory inside the loop.  Otherwise it is T launches of ballenv_step.  Results are
 * identical either way (tests/test_gpu_parity.py::test_rollout_kernel_matches_per_step_launches).
 */
int ballenv_step_many(BallenvHandle *h, const void *actions, int action_kind, int32_t n_steps, void *obs_out,
                      int32_t obs_all_steps, void *reward_out, uint8_t *done_out, ballenv_stream_t stream);

/*
 * Replaces: the policy-in-the-loop rollout of examples/ball_cnn_ac3.py:553-613 - per step prep_state4 (:560),
 * Policy(window) forward (:109-146), Categorical(probs).sample() with its .item() host round trip (:210-220) and
 * env.step(move_list[action]) (:588) - for all environments and n_steps steps in ONE launch: the environment's own
 * lane(s) evaluate softmax(action_head(relu(fc1(obs)))) from a shared-memory copy of the weights between two steps.
 * The parameters are the tensors of the torch module as they lie (nn.Linear layout, float32, device); the value
 * head is not needed to act (the update recomputes log-probabilities and values with autograd from the stored
 * observations and actions).
 * first_obs   : device float32 [n][row], the observation of the CURRENT state (what the last reset / step returned)
 * obs_out     : device float32 [n_steps][n][row], the observation after each step (post-reset where an episode ended)
 * actions_out : device int64   [n_steps][n], the action taken at each step (index into the agent move list, :530)
 * reward_out / done_out : [n_steps][n] float32 / uint8, or NULL
 * Sampling: word x of Philox4x32-10(counter = {global env id, tick, 0, stream 3}, key = seed), u = (word >> 8) 2^-24,
 * action = first j with u * sum(e) < e_0 + .. + e_j, e = exp(logit - max): Categorical by inverse CDF, reproducible and
 * independent of the sharding; greedy != 0 takes the first maximum instead.
 * Production configuration only (BALLENV_F32, gym ruleset, Philox draws, BALLENV_OBS_F32 rows) with WINDOW = 5 or 10:
 * the tuned instances (13 + 5, 8 + 24 obstacles) or the run-time-count form (any counts up to 64 obstacles, at least
 * one moving); BALLENV_ESTATE otherwise (step from the caller's policy with ballenv_step).  The block's copy of the
 * weights must fit shared memory (160 KB: hidden = 208 for WINDOW = 10 takes 99 KB).
 */
typedef struct BallenvPolicyMLP {
  int32_t n_inputs;            /* 4 + WINDOW^2 */
  int32_t hidden;              /* Policy.hidden_layer (128 for WINDOW = 5, 208 for 10); a multiple of 8 */
  int32_t greedy;              /* 0: sample, 1: argmax */
  int32_t reserved;
  const float *fc1_weight;     /* device [hidden][n_inputs] */
  const float *fc1_bias;       /* device [hidden] */
  const float *action_weight;  /* device [9][hidden] */
  const float *action_bias;    /* device [9] */
  const float *value_weight;   /* device [1][hidden], or NULL: only needed for policy_out */
  const float *value_bias;     /* device [1], or NULL */
} BallenvPolicyMLP;
/* policy_out : NULL, or device float32 [n_steps][n][10] - the 9 action probabilities and the value V(s) of the forward
 *              pass each step acted on: with unchanged weights this is exactly what the update's forward pass would
 *              recompute (ballenv_a2c_grads takes it as BallenvA2CUpdate.policy_out). */
int ballenv_rollout_policy(BallenvHandle *h, const BallenvPolicyMLP *policy, int32_t n_steps, const float *first_obs,
                           float *obs_out, int64_t *actions_out, void *reward_out, uint8_t *done_out, float *policy_out,
                           ballenv_stream_t stream);

/*
 * Replaces: the discounted-return loop of finish_episode (examples/ball_cnn_ac3.py:228-230: R = r + gamma * R, from the
 * end), for n trajectories of n_steps steps at once, restarted where an episode ended inside the slice:
 * out[t][e] = reward[t][e] + gamma * out[t + 1][e] * (1 - done[t][e]), with out[n_steps][e] = bootstrap[e] (or 0 when
 * bootstrap is NULL).  All device pointers on the CURRENT device ([n_steps][n] float32 / uint8, bootstrap [n]); float32
 * arithmetic rounded like the tensor expression.  No handle: it touches no environment state.
 */
int ballenv_discounted_returns(const float *reward, const uint8_t *done, const float *bootstrap, float gamma,
                               int32_t n_steps, int64_t n, float *out, ballenv_stream_t stream);

/*
 * Replaces: finish_episode's loss and its backward pass (examples/ball_cnn_ac3.py:233-244: policy loss
 * -log_prob(a) * (R - V.item()), value loss smooth_l1(V, R), summed over the steps, loss.backward()) for Policy(window)
 * (:109-146), over n_samples (observation, action, return) triples at once - hand-written forward + backward, no
 * activation ever leaves the chip.  All pointers device, float32, on the CURRENT device, in nn.Linear layout; the *_grad
 * arrays and loss[0] are OVERWRITTEN (not accumulated).  returns: the normalised discounted returns (:228-232), or the
 * raw ones together with returns_stats.
 * workspace: ballenv_a2c_workspace_bytes(n_inputs, hidden, n_samples) bytes of device memory.  The gradients are sums
 * in a fixed order (deterministic); they agree with autograd's to float32 rounding of the summation order.
 * n_inputs: 4 + WINDOW^2 with WINDOW = 5 or 10; hidden: a multiple of 4, at most 256.
 */
typedef struct BallenvA2CUpdate {
  int32_t n_inputs, hidden;
  const float *fc1_weight, *fc1_bias;         /* [hidden][n_inputs], [hidden] */
  const float *action_weight, *action_bias;   /* [9][hidden], [9] */
  const float *value_weight, *value_bias;     /* [1][hidden], [1] */
  float *fc1_weight_grad, *fc1_bias_grad, *action_weight_grad, *action_bias_grad, *value_weight_grad, *value_bias_grad;
  float *loss;                                /* [1] */
  const float *returns_stats;                 /* NULL: `returns` are normalised already; else device [2] = {mean, std + eps}
                                                 of the raw returns passed: (R - mean) / (std + eps) is formed in the kernel */
  const float *policy_out;                    /* NULL: the forward pass is recomputed from obs; else device [n_samples][10],
                                                 ballenv_rollout_policy's policy_out of the SAME weights: probabilities and
                                                 values are read instead (obs is still needed for the backward pass) */
} BallenvA2CUpdate;
int64_t ballenv_a2c_workspace_bytes(int32_t n_inputs, int32_t hidden, int64_t n_samples);
int ballenv_a2c_grads(const BallenvA2CUpdate *u, const float *obs /* [n_samples][n_inputs] */,
                      const int64_t *actions /* [n_samples] */, const float *returns /* [n_samples] */, int64_t n_samples,
                      void *workspace, int64_t workspace_bytes, ballenv_stream_t stream);

/* prep_state4 on the current state without stepping (examples/ball_cnn_ac3.py:384-412). */
int ballenv_observe(BallenvHandle *h, void *obs_out, ballenv_stream_t stream);

/*
 * Replaces: featureExtractor.featureExtractor(state, obstacle_list, agent_vel, agent_radius)
 * (featureExtractor.py:247-265), which createBoard calls after reset() and step() to fill self.sensor_readings
 * (ballenv_pygame.py:512, 674).  out : device float32 [n][20] (goal-distance bin, goal direction one-hot[4],
 * density[3], orientation x speed histogram[9], social forces[3]) of the CURRENT state.  Agent and obstacle
 * velocities are 0 as in the reference's only call sites; agent_radius comes from the config.
 */
int ballenv_observe_features(BallenvHandle *h, float *out, ballenv_stream_t stream);

/*
 * Replaces: prep_state2(state) of the REINFORCE / imitation / supervised scripts (examples/ball_env_reinforce.py:130-172
 * with block_to_arrpos; the same function in ball_env_imitate.py and test_model.py) - the legacy 29-float observation:
 * 4 goal-quadrant bits, then a 5 x 5 grid of obstacle counts in 20-pixel blocks around the agent (the agent's own
 * cell starts at 1).  out : device float32 [n][29] of the CURRENT state.
 */
int ballenv_observe_blocks(BallenvHandle *h, float *out, ballenv_stream_t stream);

/*
 * Replaces: extract_patch(state, width=100) of the pixel-policy scripts (examples/ball_cnn_reinforce.py:120-163; the
 * copy at examples/ball_cnn_ac3.py:248-307) - env.render(mode='rgb_array') (ballenv_env.py:357-386: white 500 x 500
 * frame, agent = black 30-gon of radius 5, goal = black quad, obstacles = 30-gons of radius 20, red if speed == 0
 * else green, drawn in that order), the frame padded with white, cropped to width x width around the agent, resized
 * with PIL (BICUBIC in ball_cnn_reinforce.py:124, BILINEAR in ball_cnn_ac3.py:255) to 40 x 40 and scaled to [0, 1].
 * out : device float32 (BALLENV_OBS_F32, value / 255) or uint8 (BALLENV_OBS_U8) [n][3][out_size][out_size] of the
 * CURRENT state.  gym ruleset only; width even, <= 128; out_size <= 64.  The resize is Pillow's 8-bit two-pass
 * resampling bit for bit; the frame restates the gym 0.10.9 viewer (pixel centre inside the polygon, no
 * anti-aliasing, coordinates rounded to integers) - parity for that half is unpinned: pyglet / OpenGL cannot run here.
 * Synchronises `stream` once per new (width, out_size, interp) combination (the resampling tables are built on the host
 * and uploaded); further calls with the same geometry only enqueue the kernel.
 */
#define BALLENV_INTERP_BILINEAR 0
#define BALLENV_INTERP_BICUBIC 1
int ballenv_observe_patches(BallenvHandle *h, void *out, int32_t width, int32_t out_size, int32_t interp,
                            int32_t out_format, ballenv_stream_t stream);

/*
 * Same as ballenv_step but with HOST buffers (pinned or pageable): copies actions host->device, steps,
 * copies obs/reward/done device->host and synchronises `stream` before returning.  This is the call the
 * end-to-end number of bench.py is measured through.
 */
int ballenv_step_host(BallenvHandle *h, const void *actions_host, int action_kind, void *obs_host,
                      void *reward_host, uint8_t *done_host, ballenv_stream_t stream);

/*
 * T consecutive ballenv_step_host calls as one call, for rollouts whose actions are known up front (open loop, or a
 * host policy that acts on observations a chunk old): actions_host [T][n], obs_host [T][n][row] (or NULL), reward_host
 * [T][n], done_host [T][n] (or NULL).  The steps run in chunks: per chunk one copy in, ONE launch of the rollout kernel
 * (as ballenv_step_many) and the copies out, on two internal copy streams with two device staging sets, so that the
 * copies of consecutive chunks overlap the kernels.  Every step's actions and results cross the bus; the state never
 * leaves the device.  Same results as T calls of ballenv_step_host; returns synchronised.  Pin the host buffers (and
 * allocate them on the GPU's NUMA node) for speed.
 */
int ballenv_step_many_host(BallenvHandle *h, const void *actions_host, int action_kind, int32_t n_steps, void *obs_host,
                           void *reward_host, uint8_t *done_host, ballenv_stream_t stream);

/*
 * Parity mode: replace Philox words by injected ones (copied to the device; NULL clears a tape).
 * step_tape  : host uint32 [n_steps][n][dynamic_obstacles][2]; step s of the tape answers the s-th
 *              ballenv_step call after this call.
 * reset_tape : host uint32 [n_episodes][n][4 + 2*attempts*static_obstacles + 2*dynamic_obstacles];
 *              row e answers the resets of episode index e (gym ruleset).
 */
int ballenv_set_draw_tape(BallenvHandle *h, const uint32_t *step_tape, int64_t n_steps,
                          const uint32_t *reset_tape, int64_t n_episodes, int32_t attempts);

/* Copy the statistics vector to host (synchronises `stream`). */
int ballenv_stats(BallenvHandle *h, double *out /* [BALLENV_NUM_STATS] host */, ballenv_stream_t stream);
int ballenv_stats_reset(BallenvHandle *h, ballenv_stream_t stream);
/* Read and clear the device error bits (synchronises `stream`). */
int ballenv_error_flags(BallenvHandle *h, uint32_t *out /* host */, ballenv_stream_t stream);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t ballenv_launch_count(BallenvHandle *h);
/* Which kernel a ballenv_step (n_steps = 1) or ballenv_step_many (n_steps > 1) call with this action kind and
 * observation rows requested launches (no reference counterpart; diagnostics, bench.py, tests):
 * BALLENV_KERNEL_GENERIC, _ROLES (block of roles, production specialisation), _LEAN (one lane per environment) or
 * _LEAN2 (a pair of lanes per environment). */
#define BALLENV_KERNEL_GENERIC 0
#define BALLENV_KERNEL_ROLES 1
#define BALLENV_KERNEL_LEAN 2
#define BALLENV_KERNEL_LEAN2 3
int ballenv_kernel_variant(BallenvHandle *h, int action_kind, int32_t n_steps);
/* Device self-tests of the arithmetic shortcuts the step kernel takes (no reference counterpart; run by the GPU
 * tests).  which = 0: the integer square root used for the distance to the goal when all coordinates are integral
 * (sqrt_int22) against sqrt() for every integer 0 <= s < arg; which = 1: the inline fp64 division of the progress
 * reward (the division's fast path without its out-of-line slow path) against `/` on arg pseudo-random pairs shaped
 * like the reward's operands.  *mismatches receives the number of differing results (synchronises the device). */
int ballenv_selftest(int which, int64_t arg, int device, int64_t *mismatches /* host */);

#ifdef __cplusplus
}
#endif
#endif /* BALLENV_H_ */
