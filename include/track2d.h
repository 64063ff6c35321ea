/*
 * track2d.h -- C ABI of libtrack2d.so: a batched, device-resident implementation of the reference's
 * gym-track2d environment (zfw1226/active_tracking_rl, envs/gym-track2d) for NVIDIA B200 (sm_100a).
 *
 * The reference has no native FFI; its boundary is the Python gym protocol.  Each entry point below
 * names the reference interface it stands in for (paths relative to the reference root, "envs/" =
 * envs/gym-track2d/gym_track2d/envs/).  The Python side (active_tracking_rl_b200/) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C: opaque handle, raw pointers, sizes; no torch / C++ types.
 *   - every function returns 0 on success or a negative T2D_E_* code; track2d_last_error() gives the
 *     message for the calling thread.
 *   - "_dev" pointers are device pointers on the handle's device, owned by the caller; `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).  Device entry points only enqueue
 *     work on `stream`; they never synchronise the host.
 *   - "_host" entry points take host pointers, do H2D -> kernels -> D2H on the handle's own stream and
 *     return after the results are in the host buffers.
 *   - one handle per device, driven from one host thread (like one gym env object per process).
 *   - agent 0 = tracker, agent 1 = target; positions are (row, col) on the 82x82 (Block/Empty) or
 *     81x81 (Maze) grid including the wall border; actions 0..3 = up, down, left, right
 *     (envs/track_1v1.py:276).
 */
#ifndef TRACK2D_H
#define TRACK2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRACK2D_ABI_VERSION 1

/* gym id grammar 'Track2D-{Map}{Obs}{Target}-v{level}' (gym_track2d/__init__.py:3-18) */
enum { T2D_MAP_BLOCK = 0, T2D_MAP_MAZE = 1, T2D_MAP_EMPTY = 2 };
enum { T2D_OBS_PARTIAL = 0, T2D_OBS_FULL = 1 };
enum { T2D_TARGET_ADV = 0, T2D_TARGET_PZR = 1, T2D_TARGET_FAR = 2, T2D_TARGET_NAV = 3, T2D_TARGET_RAM = 4, T2D_TARGET_RPF = 5 };

/* RNG behind map generation, spawn sampling and the scripted targets.
 *   PHILOX : counter-based Philox4x32-10 keyed by (seed, env); parallel sampling algorithms with the
 *            same distributions as the reference (throughput mode).
 *   NUMPY  : per-env MT19937 driven through numpy's legacy RandomState algorithms in exactly the
 *            reference's draw order; env e behaves like the reference after np.random.seed(seed + e)
 *            (with its no-argument np.random.seed() calls neutralised).  Bit-exact parity mode. */
enum { T2D_RNG_PHILOX = 0, T2D_RNG_NUMPY = 1 };

enum {
    T2D_FLAG_AUTO_RESET = 1, /* step(): envs that finish are reset in the same call and their obs replaced by the reset obs */
    T2D_FLAG_KEEP_F64 = 2,   /* keep the float64 rewards of the last step for track2d_get_rewards_f64 */
    T2D_FLAG_PLAN_AHEAD = 4  /* T2D_RNG_PHILOX + T2D_FLAG_AUTO_RESET handles only: prepare next-episode worlds (Maze maps, Nav targets) and Nav
                              * plans AHEAD of time on a side stream instead of inside step().  The results are identical either way
                              * (counter-based RNG; tests/test_gpu_fullsize.py); whether it is faster depends on what else runs on the
                              * device: kernels that fill the register file / shared memory of every SM (the learner's GEMMs) leave no
                              * room for the planners to run beside them.  track2d_set_nav is refused. */
};

enum {
    T2D_OK = 0,
    T2D_E_INVALID = -1,   /* bad argument */
    T2D_E_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    T2D_E_UNSUPPORTED = -3,
    T2D_E_STATE = -4      /* e.g. step before reset */
};

/* bits of the device status word (track2d_get_status) */
enum {
    T2D_STATUS_BAD_ACTION = 1,    /* an action outside 0..3 was seen (it was reduced mod 4) */
    T2D_STATUS_PLAN_OVERFLOW = 2, /* an A* plan exceeded TRACK2D_NAV_MAXPLAN */
    T2D_STATUS_ASTAR_REPLACE = 4, /* Frontier.replace condition observed (provably unreachable on a unit-cost grid) */
    T2D_STATUS_HEAP_OVERFLOW = 8
};

#define TRACK2D_FOV 13            /* 2 * pob_size + 1, envs/track_1v1.py:17 */
#define TRACK2D_NAV_MAXPLAN 1024  /* longest stored Navigator plan (actions) */
#define TRACK2D_RAM_MAXPLAN 9     /* RamAgent plans have 1..9 actions, envs/navigator.py:85,91 */

typedef struct track2d_env track2d_env;

typedef struct track2d_config {
    int32_t abi_version;       /* TRACK2D_ABI_VERSION */
    int32_t num_envs;          /* E >= 1 */
    int32_t map_type;          /* T2D_MAP_*     (kwargs of gym.make, gym_track2d/__init__.py:12-16) */
    int32_t obs_type;          /* T2D_OBS_* */
    int32_t target_mode;       /* T2D_TARGET_* */
    int32_t level;             /* 0 = random density, >0 = fixed (envs/track_1v1.py:220-229) */
    int32_t rng_mode;          /* T2D_RNG_* */
    int32_t device;            /* CUDA device ordinal */
    int32_t max_episode_steps; /* gym TimeLimit; 500 in every registered id (gym_track2d/__init__.py:17); <=0 disables */
    int32_t flags;             /* T2D_FLAG_* */
    uint64_t seed;
} track2d_config;

/* ---- lifetime ------------------------------------------------------------------------------- */

/* gym.make(id): Track1v1Env.__init__ (envs/track_1v1.py:14-69) for num_envs envs.  Allocates the
 * struct-of-arrays world state on cfg->device.  No map is drawn until the first reset. */
int track2d_create(const track2d_config *cfg, track2d_env **out);
/* env.close() */
int track2d_destroy(track2d_env *env);
const char *track2d_last_error(void);
int track2d_abi_version(void);
/* number of this library's kernels enqueued so far in this process (host-side count; bench.py's gpu_launches) */
uint64_t track2d_launch_count(void);

/* observation_space / define_observation (envs/track_1v1.py:252-262): cells per agent per env:
 * 169 (Partial) or H*W (Full).  An obs buffer holds num_envs * 2 * cells floats, laid out
 * [env][agent][row][col] -- the reference's (2, 1, h, w) per env, batched on a leading axis. */
int track2d_obs_cells(const track2d_env *env);
int track2d_num_envs(const track2d_env *env);
int track2d_map_height(const track2d_env *env);
int track2d_map_width(const track2d_env *env);

/* ---- device-resident API (caller-owned device buffers, stream-ordered, no host sync) --------- */

/* Track1v1Env.reset (envs/track_1v1.py:134-168) + TimeLimit.reset, for every env whose mask byte is
 * non-zero (mask_dev == NULL: all envs).  Writes the reset observations of those envs into obs_dev
 * (float32 [E][2][cells]; other envs' rows are left untouched).  obs_dev may be NULL. */
int track2d_reset(track2d_env *env, const uint8_t *mask_dev, float *obs_dev, void *stream);

/* Track1v1Env.step (envs/track_1v1.py:71-127) + gym 0.12.5 TimeLimit.step, for all envs.
 *   actions_dev : int32 [E][2]  (tracker, target); the target entry is ignored for Ram/Nav/RPF
 *   obs_dev     : float32 [E][2][cells], 16-byte aligned  (environment.py:146 casts to float32)
 *   reward_dev  : float32 [E][2]  (r_track, r_target; player_util.py:58 casts to float32)
 *   done_dev    : uint8 [E]       (far-counter done OR elapsed >= max_episode_steps)
 * With T2D_FLAG_AUTO_RESET, finished envs are then reset and their obs rows replaced. */
int track2d_step(track2d_env *env, const int32_t *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev, void *stream);

/* Same transition, observations written as uint8 (values 0,1,2,4) instead of float32 -- a lower
 * traffic path for consumers that convert on load.  Not the contract dtype; reported separately. */
int track2d_step_u8(track2d_env *env, const int32_t *actions_dev, uint8_t *obs_dev, float *reward_dev, uint8_t *done_dev, void *stream);
int track2d_reset_u8(track2d_env *env, const uint8_t *mask_dev, uint8_t *obs_dev, void *stream);

/* ---- host-buffer API (what a numpy-facing gym user calls; H2D/D2H inside) -------------------- */
int track2d_reset_host(track2d_env *env, const uint8_t *mask_host, float *obs_host);
int track2d_step_host(track2d_env *env, const int32_t *actions_host, float *obs_host, float *reward_host, uint8_t *done_host);
/* the same with uint8 observations (values 0,1,2,4): a quarter of the PCIe traffic; the consumer converts after upload */
int track2d_reset_host_u8(track2d_env *env, const uint8_t *mask_host, uint8_t *obs_host);
int track2d_step_host_u8(track2d_env *env, const int32_t *actions_host, uint8_t *obs_host, float *reward_host, uint8_t *done_host);

/* Makes `stream` wait for the work this handle still has in flight on its side stream (standby worlds, plans made ahead of time).
 * Needed before a CUDA-graph capture that contains track2d_step calls ends, and before reading state another way than through this
 * API; a no-op for handles without side work.  Never blocks the host. */
int track2d_join(track2d_env *env, void *stream);

/* The pipelined form of track2d_step_host[_u8]: same transition, but the observation D2H is issued as n_chunks (1..16) pieces of
 * consecutive envs and the call returns after ENQUEUEING.  track2d_host_chunk_wait(env, c) blocks until chunk c -- envs
 * [E c / n, E (c+1) / n) -- is in obs_host (reward_host / done_host are complete with chunk 0), so a consumer can process or re-upload
 * chunk c while later chunks are still crossing the bus (PCIe is full duplex).  Host buffers should be pinned.  obs_host is float32
 * (obs_is_u8 == 0, the reference's dtype, environment.py:146) or uint8. */
int track2d_step_host_begin(track2d_env *env, const int32_t *actions_host, void *obs_host, int32_t obs_is_u8, float *reward_host,
                            uint8_t *done_host, int32_t n_chunks);
int track2d_host_chunk_wait(track2d_env *env, int32_t chunk);
/* Device-side form of the same dependency: work enqueued on `stream` after this call starts once chunk `chunk` has landed in the host
 * buffers -- no host wake-up between a chunk's D2H and its consumer's H2D.  The host must not touch the buffers before it has
 * synchronised with `stream` (or called track2d_host_chunk_wait). */
int track2d_host_chunk_wait_stream(track2d_env *env, int32_t chunk, void *stream);

/* ---- state read-back / injection (host pointers, synchronous; parity tests and the gym shim) -- */

/* maps as uint8 [count][H][W], 1 = wall (Track1v1Env.maze) */
int track2d_get_maps(track2d_env *env, int32_t first, int32_t count, uint8_t *maze_host);
int track2d_set_maps(track2d_env *env, int32_t first, int32_t count, const uint8_t *maze_host);
/* positions int32 [count][2][2] (Track1v1Env.state); counters int32 [count][2] = (C_far, elapsed_steps) */
int track2d_get_agents(track2d_env *env, int32_t first, int32_t count, int32_t *pos_host, int32_t *counters_host);
int track2d_set_agents(track2d_env *env, int32_t first, int32_t count, const int32_t *pos_host, const int32_t *counters_host);
/* goal_states int32 [count][2][2] (envs/track_1v1.py:237) */
int track2d_get_goals(track2d_env *env, int32_t first, int32_t count, int32_t *goals_host);
/* RamAgent.plan_actions / a_i (envs/navigator.py:73-93): plan int32 [count][TRACK2D_RAM_MAXPLAN], len, idx int32 [count] */
int track2d_get_ram(track2d_env *env, int32_t first, int32_t count, int32_t *plan_host, int32_t *len_host, int32_t *idx_host);
int track2d_set_ram(track2d_env *env, int32_t first, int32_t count, const int32_t *plan_host, const int32_t *len_host, const int32_t *idx_host);
/* Navigator.plan_actions / a_i / goal_states (envs/navigator.py:5-63): plan int32 [count][TRACK2D_NAV_MAXPLAN] */
int track2d_get_nav(track2d_env *env, int32_t first, int32_t count, int32_t *plan_host, int32_t *len_host, int32_t *idx_host, int32_t *goal_host);
int track2d_set_nav(track2d_env *env, int32_t first, int32_t count, const int32_t *plan_host, const int32_t *len_host, const int32_t *idx_host, const int32_t *goal_host);
/* AstarSolver(env, goal).solve() + get_actions (envs/Astar_solver.py:102-149) called directly, for known-answer tests: plans from
 * start_host[i] to goal_host[i] ((row, col), int32 [count][2]) on the generator maze of env first + i.  len_host[i] = number of actions
 * or -1 when the goal is unreachable; plan_host (may be NULL) int32 [count][TRACK2D_NAV_MAXPLAN].  Overwrites the Navigator plan of
 * those envs.  Nav / RPF handles only; synchronous. */
int track2d_astar_solve(track2d_env *env, int32_t first, int32_t count, const int32_t *start_host, const int32_t *goal_host,
                        int32_t *plan_host, int32_t *len_host);
/* float64 rewards of the last step, [E][2] (needs T2D_FLAG_KEEP_F64): exactly what Track1v1Env.step returns */
int track2d_get_rewards_f64(track2d_env *env, int32_t first, int32_t count, double *rewards_host);
/* the action the target actually executed in the last step (Ram/Nav/RPF override), int32 [count] */
int track2d_get_target_actions(track2d_env *env, int32_t first, int32_t count, int32_t *actions_host);
/* T2D_RNG_NUMPY: np.random.seed(seed) for one env / read its MT19937 state (key[624], pos) */
int track2d_seed_env(track2d_env *env, int32_t index, uint32_t seed);
int track2d_get_rng_numpy(track2d_env *env, int32_t index, uint32_t *key_host, int32_t *pos_host);
/* Track1v1Env.__init__ draws one map it never uses (envs/track_1v1.py:45); in T2D_RNG_NUMPY mode the
 * gym shim calls this so that `np.random.seed(s); gym.make(id); reset()` stays stream-exact. */
int track2d_init_maze(track2d_env *env, const uint8_t *mask_dev, void *stream);
/* OR of T2D_STATUS_* seen so far (synchronises the handle's device work on `stream`) */
int track2d_get_status(track2d_env *env, uint32_t *status_host, void *stream);
/* number of envs reset by the last auto-reset step / episodes finished so far (synchronous) */
int track2d_get_counters(track2d_env *env, uint64_t *episodes_done, uint64_t *steps_done);

/* ---- learner kernels on the same stream (reference: shared_optim.py:122-175, player_util.py:157) */

/* clip_grad_norm_(params, max_norm) followed by SharedAdam.step (AMSGrad, eps added AFTER the sqrt,
 * old-style bias correction) over one flat fp32 parameter vector.  All pointers are device pointers
 * of n floats; `step` is the 1-based update count (state['step'] after the increment);
 * `grad_scale` multiplies the gradient first (1/world_size after a sum-allreduce).
 * `norm_scratch_dev` is TWO floats of scratch (squared global norm, step size).  If `step_dev` is not NULL it points to
 * a device-resident int64 update counter that the call increments and uses instead of `step` (the bias correction is then
 * computed on the device): that form can be captured in a CUDA graph and replayed. */
int track2d_sharedadam_step(float *param_dev, const float *grad_dev, float *exp_avg_dev, float *exp_avg_sq_dev,
                            float *max_exp_avg_sq_dev, int64_t n, int64_t step, double lr, double beta1, double beta2,
                            double eps, double max_grad_norm, double grad_scale, float *norm_scratch_dev, int64_t *step_dev,
                            void *stream);

/* The backward recursion of Agent.optimize (player_util.py:127-140) for all envs: n-step returns R_t and
 * GAE advantages, cut at episode ends.  rewards [T][E][2], done [T][E], values [T+1][E][2] (row T = the
 * bootstrap value V(s_T), ignored where done[T-1]); outputs returns, gae [T][E][2].  Device pointers. */
int track2d_gae_returns(const float *rewards_dev, const uint8_t *done_dev, const float *values_dev, float *returns_dev,
                        float *gae_dev, int32_t T, int64_t E, double gamma, double tau, void *stream);

/* CNN_maze's convolution stack (perception.py:71-72,86-88) fused: conv1 3x3/s2/p1 (1->16) + ReLU + conv2 3x3/s2/p1
 * (16->32) + ReLU.  x [n_images][13][13] float32 -> y2 [n_images][512] in (channel, row, col) order, i.e. what
 * `x.view(1, -1)` (perception.py:89) flattens.  Device pointers; w1 [16][1][3][3], w2 [32][16][3][3]. */
int track2d_maze_conv_forward(const float *x_dev, int64_t n_images, const float *w1_dev, const float *b1_dev, const float *w2_dev,
                              const float *b2_dev, float *y2_dev, void *stream);
/* its backward: given gy2 = dL/dy2, ACCUMULATES (+=) dL/dw1, db1, dw2, db2 into the given buffers (zero them first for a
 * fresh gradient).  The conv1 activations are recomputed from x; the observation gets no gradient. */
int track2d_maze_conv_backward(const float *x_dev, const float *y2_dev, const float *gy2_dev, int64_t n_images, const float *w1_dev,
                               const float *b1_dev, const float *w2_dev, float *dw1_dev, float *db1_dev, float *dw2_dev, float *db2_dev,
                               void *stream);

/* the same two kernels with the image source made explicit: x is float32 (x_is_u8 == 0) or uint8 (the env's lossless observation
 * encoding, cell values 0, 1, 2, 4: track2d_step_u8), consecutive images x_stride ELEMENTS apart (169 = packed; 338 = one agent's
 * image out of an [E][2][13][13] observation buffer). */
int track2d_maze_conv_forward_ex(const void *x_dev, int32_t x_is_u8, int64_t x_stride, int64_t n_images, const float *w1_dev, const float *b1_dev,
                                 const float *w2_dev, const float *b2_dev, float *y2_dev, void *stream);
int track2d_maze_conv_backward_ex(const void *x_dev, int32_t x_is_u8, int64_t x_stride, const float *y2_dev, const float *gy2_dev, int64_t n_images,
                                  const float *w1_dev, const float *b1_dev, const float *w2_dev, float *dw1_dev, float *db1_dev, float *dw2_dev,
                                  float *db2_dev, void *stream);

/* float32-accurate GEMM on the tcgen05 tensor cores (3xTF32 split, fp32 accumulation in TMEM) for the policy's Linear /
 * LSTMCell layers (reference: nn.Linear / nn.LSTMCell in model.py:116-127,175-182, perception.py:73 -- fp32 on the CPU):
 *     D[m][n] = act( sum_k A(m,k) * B(n,k) + bias[n] ),  m < M, n < N, k < K;  D row-major with leading dimension ldd.
 * A(m,k) is read at a_dev[m*lda + k] (a_mn_major == 0, "K-major") or a_dev[k*lda + m] (a_mn_major == 1); B(n,k) likewise at
 * b_dev[n*ldb + k] or b_dev[k*ldb + n].  So y = x W^T is (A=x K-major, B=W K-major), dx = dy W is (A=dy K-major, B=W MN-major)
 * and dW = dy^T x is (A=dy MN-major, B=x MN-major, K = batch) with no transposed copies.  bias_dev may be NULL; relu != 0 applies
 * max(., 0).  All pointers are device pointers, 16-byte aligned; contiguous extents and leading dimensions are multiples of 4.
 * When few output tiles exist the reduction is split across SMs: pass a workspace of track2d_gemm_workspace_floats(M, N, K)
 * floats (0 = none needed).  The partial sums are combined in a fixed order (bit-reproducible). */
int64_t track2d_gemm_workspace_floats(int64_t M, int64_t N, int64_t K);
int track2d_gemm_tf32x3(const float *a_dev, int a_mn_major, int64_t lda, const float *b_dev, int b_mn_major, int64_t ldb,
                        float *d_dev, int64_t ldd, int64_t M, int64_t N, int64_t K, const float *bias_dev, int relu,
                        float *workspace_dev, int64_t workspace_floats, void *stream);

/* The pointwise half of nn.LSTMCell (model.py:116,176) for E rows, hidden size H (128): gates = igates + hgates + b_ih + b_hh
 * in PyTorch's [i | f | g | o] order -> hy, cy [E][H] and the activated gates act [E][4H] (kept for the backward).  cx may be a
 * strided view (row stride cx_stride floats).  The backward takes dL/dhy, dL/dcy (either may be NULL = zero; row strides in
 * floats) and writes dgates [E][4H] (the gradient of igates AND of hgates), dcx [E][H] and dbias [4H] (the gradient of b_ih and
 * of b_hh: the column sums of dgates, combined in a fixed order through a workspace of
 * track2d_lstm_bias_workspace_floats(E, H) floats).  Device pointers, 16-byte aligned. */
int64_t track2d_lstm_bias_workspace_floats(int64_t E, int32_t H);
int track2d_lstm_cell_forward(const float *igates_dev, const float *hgates_dev, const float *b_ih_dev, const float *b_hh_dev,
                              const float *cx_dev, int64_t cx_stride, float *hy_dev, float *cy_dev, float *act_dev, int64_t E, int32_t H,
                              void *stream);
int track2d_lstm_cell_backward(const float *dhy_dev, int64_t dhy_stride, const float *dcy_dev, int64_t dcy_stride, const float *cx_dev,
                               int64_t cx_stride, const float *cy_dev, const float *act_dev, float *dgates_dev, float *dcx_dev,
                               float *dbias_dev, float *workspace_dev, int64_t workspace_floats, int64_t E, int32_t H, void *stream);

/* out[n] = sum_m x[m*ld + n] (bias gradient of a Linear layer: the reference's autograd sums dy over the batch).  N a power of
 * two in [4, 1024]; workspace of track2d_colsum_workspace_floats(M, N) floats; fixed summation order. */
int64_t track2d_colsum_workspace_floats(int64_t M, int32_t N);
int track2d_colsum(const float *x_dev, int64_t ld, int64_t M, int32_t N, float *out_dev, float *workspace_dev, int64_t workspace_floats,
                   void *stream);

/* ---- the recurrent half of the policy step and the A3C loss, fused (csrc/track2d_a3c.cu) ------------------------------------
 * Reference: model.py:41-50 (sample_action), :116-127 / :175-209 (LSTMCell + actor / critic / reward_aux heads),
 * player_util.py:108-161 (Agent.optimize).  Hidden size 128.  "Packed heads" = an [8][128] weight / [8] bias / [E][8] output
 * whose rows are: 4 actor logits, the critic value, the TAT's reward prediction (or 0), 2 x padding.  All pointers are device
 * pointers, 16-byte aligned unless they address single columns. */

/* One policy step after the gate GEMM (gates_dev [E][512] = [feature | h] [W_ih | W_hh]^T, no bias): LSTM cell -> packed heads ->
 * softmax / log-softmax / entropy -> action.  Writes act_dev [E][512] (activated gates, for the backward; may be NULL), c_next_dev
 * [E][128], h_out_dev [E][128] (may be NULL), the same h into h_next_dev rows of stride h_next_ld floats (the recurrent input slot
 * of the next step's GEMM; may be NULL), out8_dev [E][8] (may be NULL).  action_dev / forced_dev / value_dev / logp_dev /
 * entropy_dev address ONE COLUMN of [E][2] arrays (element stride 2).  The action is forced_dev's when given, the argmax when
 * `greedy` (player_util.py:69-82 action_test; logp_all_dev [E][4] then receives the log-probabilities, model.py:45-46), otherwise
 * a multinomial sample drawn with Philox4x32-10 keyed by `seed` at counter (first_row + row, rng_stream, *rng_step_dev): a batch
 * processed in slices (pointers advanced to the slice, first_row = its first row) draws the same actions as the whole batch. */
int track2d_lstm_heads_forward(const float *gates_dev, const float *b_ih_dev, const float *b_hh_dev, const float *c_prev_dev, float *act_dev,
                               float *c_next_dev, float *h_out_dev, float *h_next_dev, int64_t h_next_ld, const float *w_head_dev,
                               const float *b_head_dev, float *out8_dev, int32_t *action_dev, const int32_t *forced_dev, float *value_dev,
                               float *logp_dev, float *entropy_dev, float *logp_all_dev, const uint64_t *rng_step_dev, uint64_t seed,
                               uint32_t rng_stream, int32_t greedy, int64_t E, int64_t first_row, void *stream);
/* After env.step: envs whose done byte is set start the next step from zero recurrent state (train.py:73-74 -> Agent.reset):
 * zeroes their rows of h0 / h1 (row stride h_ld) and c0 / c1 ([E][128]); eps_len += 1 or = 0 (player_util.py:63,101); advances
 * the sampling counter.  Any pointer may be NULL. */
int track2d_policy_post_step(const uint8_t *done_dev, float *h0_dev, float *h1_dev, int64_t h_ld, float *c0_dev, float *c1_dev,
                             int32_t *eps_len_dev, uint64_t *rng_step_dev, int64_t E, void *stream);
/* out[e][j] = x[e][j] + w[j][action[e]] + b[j], j < N: TAT's fc_action_tracker applied to the one-hot tracker action
 * (model.py:198-199); w [N][4]; row strides ld / out_ld floats; action_dev = one column of an int32 [E][2] array. */
int track2d_embed_add(const float *x_dev, int64_t ld, float *out_dev, int64_t out_ld, const float *w_dev, const float *b_dev,
                      const int32_t *action_dev, int32_t N, int64_t E, void *stream);
/* Agent.optimize's recursions and loss gradients for all envs (player_util.py:117-145): from the packed head outputs of both
 * agents (out8_*_dev [T+1][E][8], slot T = the bootstrap forward), actions [T][E][2], rewards [T][E][2], done [T][E]:
 * n-step returns and GAE cut at episode ends, then dL/d(out8) -> dout8_*_dev [T][E][8] for L = scale * sum_e sum_t of the
 * reference's per-step loss terms of the agents with train_k != 0 (+ the aux L1 when use_aux == 2; use_aux == 1 only reports it, 0 = the
 * target has no reward_aux head), and the per-env statistics
 * stats_dev [7][E] = policy_loss 0/1, value_loss 0/1, entropy 0/1, pred_loss.  returns_dev / gae_dev [T][E][2] optional. */
int track2d_a3c_loss_grad(const float *out8_0_dev, const float *out8_1_dev, float *dout8_0_dev, float *dout8_1_dev, const int32_t *actions_dev,
                          const float *rewards_dev, const uint8_t *done_dev, float *stats_dev, float *returns_dev, float *gae_dev, int32_t T,
                          int64_t E, double gamma, double tau, double w_ent0, double w_ent1, double scale, int32_t train0, int32_t train1,
                          int32_t use_aux, void *stream);
/* Backward of track2d_lstm_heads_forward for one step of the BPTT sweep: dgates_dev [E][512] from dout8_dev [E][8] (through the
 * packed head weights) plus, when dh_rec_dev != NULL, the recurrent gradients dh_rec_dev [E][128] / dc_dev [E][128] of step t+1,
 * both dropped where done_dev[e] (the episode ended at this step).  dc_dev is overwritten with dL/dc_prev. */
int track2d_lstm_heads_backward(const float *dout8_dev, const float *w_head_dev, const float *dh_rec_dev, float *dc_dev, const uint8_t *done_dev,
                                const float *act_dev, const float *c_prev_dev, float *dgates_dev, int64_t E, void *stream);
/* In place dy *= (y > 0) over [M][N] (y rows y_ld floats apart; ReLU backward of the encoder fc, perception.py:91) and, in the same pass, dbias = column sums
 * of the result; with group_dev (one column of an int32 [M][2] array, values 0..3) also the gradients of a Linear(4, N) whose
 * one-hot-selected output was added after the ReLU: gw_dev [N][4] = per-group column sums of the incoming dy, gb_dev [N] = their
 * total.  Fixed summation order; workspace of track2d_relu_backward_workspace_floats(M, N) floats. */
int64_t track2d_relu_backward_workspace_floats(int64_t M, int32_t N);
int track2d_relu_backward_groupsum(float *dy_dev, const float *y_dev, int64_t y_ld, const int32_t *group_dev, int64_t M, int32_t N,
                                   float *dbias_dev, float *gw_dev, float *gb_dev, float *workspace_dev, int64_t workspace_floats, void *stream);

/* ---- gradient all-reduce over NVLink / NVSwitch peer memory (csrc/track2d_peer.cu) -------------------------------------------
 * Reference: main.py:102-116 + utils.py:36-44 (ensure_shared_grads): the W workers' gradients meet in one shared model.  Here every
 * GPU applies the same update to the sum of all ranks' flat fp32 gradients; this is that sum, as plain kernels on the caller's
 * stream (CUDA-graph capturable), reading the peers' copies directly: grad[i] = sum over ranks 0..world-1, in that order, of the
 * peers' grad[i] -- the same bits on every rank.  One process per GPU: create, exchange the 64-byte handles (any transport),
 * connect, then call track2d_peer_allreduce once per update on EVERY rank.  Waits for a peer are bounded; a peer that never arrives is
 * reported by track2d_peer_status (0 = ok) instead of hanging the device.  world <= 8. */
typedef struct track2d_peer track2d_peer;
int track2d_peer_create(int32_t rank, int32_t world, int64_t n_floats, int32_t device, track2d_peer **out);
int track2d_peer_handle(track2d_peer *p, uint8_t *handle64_out);                 /* cudaIpcGetMemHandle of this rank's segment */
int track2d_peer_connect(track2d_peer *p, const uint8_t *handles_world_x_64);    /* all ranks' handles, rank-major */
int track2d_peer_segment(track2d_peer *p, void **segment_out);                   /* same-process peers: the raw segment pointer ... */
int track2d_peer_connect_local(track2d_peer *p, void *const *segments_world);    /* ... and a connect that takes them (tests) */
int track2d_peer_allreduce(track2d_peer *p, float *grad_dev, void *stream);      /* in place; grad_dev 16-byte aligned, n_floats long */
/* The exchange FUSED with the update it feeds (one kernel sums the ranks' gradients and applies SharedAdam.step, shared_optim.py:122-175):
 * bit for bit track2d_peer_allreduce followed by track2d_sharedadam_step with max_grad_norm = 0 and the device-resident counter
 * (clipping needs the global norm of the whole sum first -- use the two calls then).  grad_dev ends up holding the sum. */
int track2d_peer_sharedadam_step(track2d_peer *p, float *param, float *grad_dev, float *exp_avg, float *exp_avg_sq, float *max_exp_avg_sq,
                                 double lr, double beta1, double beta2, double eps, double grad_scale, float *norm_scratch,
                                 int64_t *step_dev, void *stream);
int track2d_peer_status(track2d_peer *p, uint64_t *status_out);
void track2d_peer_destroy(track2d_peer *p);

#ifdef __cplusplus
}
#endif
#endif /* TRACK2D_H */
