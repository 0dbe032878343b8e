/*
 * wab_b200.h — C ABI of the B200-native batched Wolves-and-Bushes simulator.
 *
 * The reference (johnmatthewtennant/wab-gym) has no FFI / plugin layer: its only boundary is the
 * Python gym surface of WolvesAndBushesEnv (/root/reference/wab_env.py:103-342). This header is the
 * boundary a binding for that surface calls into; each entry point names the reference interface it
 * replaces. Plain pointers and sizes only — no torch types. All `d_*` pointers are DEVICE pointers
 * owned by the caller (16-byte aligned for `d_grids`), all work is enqueued on the caller's
 * cudaStream_t (passed as void* so C callers need no CUDA headers), nothing synchronises unless
 * stated. One handle <-> one device <-> one stream at a time; handles are independent.
 *
 * Return value of every int function: 0 on success, a WAB_E_* code otherwise; the message is
 * available from wab_last_error() (thread-local).
 */
#ifndef WAB_B200_H
#define WAB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAB_ABI_VERSION 1
#define WAB_MAX_ACTIONS 8

enum {
    WAB_OK = 0,
    WAB_E_NULL = 1,        /* null pointer argument                                              */
    WAB_E_CONFIG = 2,      /* invalid option values (reference: ValueError, wab_env.py:147-148)  */
    WAB_E_UNSUPPORTED = 3, /* valid for the reference but outside what the kernels implement     */
    WAB_E_CUDA = 4,        /* CUDA runtime error (message carries cudaGetErrorString)            */
    WAB_E_NO_DEVICE = 5    /* no CUDA device: there is no CPU fallback                           */
};

/* Food arithmetic. The reference keeps food as float64 (wab_env.py:307-322, :452). WAB_FOOD_INT is
 * an integer counter in units of 1/turns_to_empty_food, legal only when the host has proven it
 * equivalent for every reachable state (wab_gym_b200/config.py: prove_integer_food). */
enum { WAB_FOOD_F64 = 0, WAB_FOOD_INT = 1 };

/* Rule constants — the POD image of `default_game_options` (wab_env.py:11-39) plus the action table
 * (wab_env.py:149-182) and host-precomputed integer thresholds for the keyed draws. */
typedef struct WabConfig {
    int32_t abi_version;          /* WAB_ABI_VERSION                                               */
    int32_t width, height;        /* viewport (odd; kernels implement 11x11)      wab_env.py:25-26 */
    int32_t max_turns;            /*                                              wab_env.py:23    */
    int32_t wolf_spawn_margin;    /* kernels implement 1                          wab_env.py:34    */
    int32_t n_actions;            /* 5 or 6                                       wab_env.py:149-191 */
    int8_t action_dx[WAB_MAX_ACTIONS];
    int8_t action_dy[WAB_MAX_ACTIONS];
    int8_t action_role[WAB_MAX_ACTIONS]; /* -1 = keep role (NaN in the reference) wab_env.py:257-258 */
    uint8_t lookout_only;         /*                                              wab_env.py:19, :302 */
    uint8_t restrict_view;        /*                                              wab_env.py:20, :351 */
    uint8_t wolves;               /*                                              wab_env.py:37    */
    uint8_t wolves_can_move;      /*                                              wab_env.py:38    */
    uint8_t god_mode;             /*                                              wab_env.py:292   */
    int8_t starting_role;         /* -1 = random                                  wab_env.py:598-599 */
    uint8_t food_mode;            /* WAB_FOOD_F64 / WAB_FOOD_INT                                   */
    uint8_t auto_reset;           /* 1: done envs are reset inside step() and their post-reset
                                     observation is returned (VecEnv); 0: reference behaviour,
                                     stepping after done keeps returning done     wab_env.py:328-340 */
    int32_t food_int_start;       /* INT mode: starting_food * turns_to_empty_food                 */
    int32_t food_int_inc;         /* INT mode: turns_to_empty_food / turns_to_fill_food            */
    int32_t food_int_max;         /* INT mode: turns_to_empty_food (clip upper bound)              */
    int32_t wolf_cap;             /* wolf slots per env (<= 64); overflow is counted, see stats    */
    int32_t log_cap;              /* depletion-log slots per env (<= 255); overflow is counted     */
    double food_start;            /* F64 mode: starting_food; < 0 = random        wab_env.py:596-597 */
    double food_inc;              /* 1 / turns_to_fill_food                       wab_env.py:307-309 */
    double food_dec;              /* 1 / turns_to_empty_food                      wab_env.py:316   */
    double food_obs_scale;        /* turns_to_empty_food                          wab_env.py:452   */
    /* binomial-first draws (oracle/keyed_rng.py): K = #{k < 32 : v >= cdf[k]} events among n cells, v a 64-bit
       draw; cdf[k] = min(ceil(BinomialCDF(n, chance/2; k) * 2^64), 2^64 - 1)                                  */
    uint64_t spawn_cdf[32];       /* n = 48 ring cells per step                   wab_env.py:571-574 */
    uint64_t init_cdf[32];        /* n = 121 window cells per reset               wab_env.py:588-591 */
    uint64_t thr_keep;            /* wolf kept iff word >= thr (U > despawn)      wab_env.py:262-264 */
    float reward_table[8];        /* [ate*4 + outcome], outcome 0 alive, 1 finished, 2 starved,
                                     3 killed; the f32 image of the fp64 sums     wab_env.py:328-340 */
    uint32_t mask_lookout[4];     /* 121-bit blind-spot masks, bit = i*11 + j     wab_env.py:109-139 */
    uint32_t mask_gatherer[4];
} WabConfig;

/* Observation batch: the first six elements of the reference's tuple (wab_env.py:374-385) for N
 * envs. The seventh (view_mask) is a function of role and options only. */
typedef struct WabObs {
    uint8_t *d_grids;  /* [N][3][11][11] u8: wolves, bushes, ostriches; cell [i][j] = [5-dx][5-dy]. May be NULL in
                        * wab_vec_reset / wab_vec_step / wab_vec_step_many while a feature buffer is bound
                        * (wab_vec_bind_features): FEATURES-ONLY stepping — the 28 PragmaticObsWrapper bytes per env, the
                        * scalars, reward, done and info are written, the one-hot grids are not materialised. */
    uint8_t *d_food;   /* [N] ceil(food * turns_to_empty_food)                     wab_env.py:452  */
    uint8_t *d_role;   /* [N]                                                      wab_env.py:390  */
    uint8_t *d_status; /* [N] 0 alive, 1 starved, 2 killed                         wab_env.py:387  */
} WabObs;

/* info byte written by step (optional): bits 0-1 outcome (index into reward_table), bit 2 ate,
 * bit 3 invalid action, bits 4-5 status before auto-reset. */
#define WAB_INFO_OUTCOME(b) ((b) & 3)
#define WAB_INFO_ATE(b) (((b) >> 2) & 1)
#define WAB_INFO_BAD_ACTION(b) (((b) >> 3) & 1)
#define WAB_INFO_FINAL_STATUS(b) (((b) >> 4) & 3)

/* indices into the int64[8] vector of wab_vec_stats */
enum {
    WAB_STAT_EPISODES = 0, WAB_STAT_STEPS = 1, WAB_STAT_FINISHED = 2, WAB_STAT_STARVED = 3,
    WAB_STAT_KILLED = 4, WAB_STAT_EATS = 5, WAB_STAT_BAD_ACTIONS = 6, WAB_STAT_OVERFLOWS = 7
};

typedef struct WabVec WabVec;

/* Replaces WolvesAndBushesEnv.__init__ (wab_env.py:106-186) for n_envs environments with global ids
 * env_id_base .. env_id_base + n_envs - 1 (keys depend on the global id only, so results do not
 * depend on how a batch is sharded over GPUs). bush_thr[k-1] = least 32-bit draw whose bush value
 * round(U**bush_power * max_berries) is >= k (wab_env.py:631-635). Allocates state on `device`.
 * Does NOT reset (the reference constructor does, :186): call wab_vec_reset. */
int wab_vec_create(const WabConfig *cfg, const uint32_t *bush_thr, int32_t n_bush_thr, int64_t n_envs,
                   uint64_t seed, uint64_t env_id_base, int32_t device, WabVec **out);

/* Replaces reset() (wab_env.py:231-248). d_mask: NULL = all envs, else u8[N], nonzero = reset. Each
 * reset env starts its next episode (first reset -> episode 0). Observations are written for EVERY
 * env: the others get the observation their last step returned again — including the bush under the
 * ostrich if that step ate it empty, which the reference's observation still shows because its frame
 * (wab_env.py:266) predates the eat (:300-313) and is only rebuilt by the next step. */
int wab_vec_reset(WabVec *h, const uint8_t *d_mask, WabObs obs, void *stream);

/* Replaces step(action) (wab_env.py:250-342) for all envs. d_actions u8[N]; outputs: obs, d_reward
 * f32[N], d_done u8[N], d_info u8[N] (may be NULL). */
int wab_vec_step(WabVec *h, const uint8_t *d_actions, WabObs obs, float *d_reward, uint8_t *d_done,
                 uint8_t *d_info, void *stream);

/* T lockstep steps in ONE launch: d_actions u8[T][N]; every step's results are materialised:
 * grids [T][N][3][11][11], food/role/status/done/info [T][N], reward [T][N] (info may be NULL). */
int wab_vec_step_many(WabVec *h, int32_t n_steps, const uint8_t *d_actions, WabObs obs, float *d_reward,
                      uint8_t *d_done, uint8_t *d_info, void *stream);

/* Same contract as wab_vec_step with HOST buffers (pageable or pinned): copies actions in, steps,
 * copies every output back, and synchronises the stream before returning. h_info may be NULL. */
int wab_vec_step_host(WabVec *h, const uint8_t *h_actions, uint8_t *h_grids, uint8_t *h_food,
                      uint8_t *h_role, uint8_t *h_status, float *h_reward, uint8_t *h_done,
                      uint8_t *h_info, void *stream);
/* Same step with ONE contiguous host block for every output (one device-to-host transfer instead of
 * seven): wab_vec_host_block_layout gives the byte offsets of grids, food, role, status, reward,
 * done, info inside the block (each 256-byte aligned) and its total size. With PINNED buffers and a batch
 * of at most 16,384 envs the kernel reads the actions from, and writes the block into, host memory directly
 * (no staging copy); either way the block is complete when the call returns. */
int wab_vec_host_block_layout(const WabVec *h, int64_t *offsets7, int64_t *total_bytes);
int wab_vec_step_host_packed(WabVec *h, const uint8_t *h_actions, uint8_t *h_block, void *stream);
/* The host-buffer entry points remember, by host address, what they resolved for the caller's buffers (the device
 * aliases of pinned memory, the captured copy-step-copy graph). Call this before freeing or unregistering a buffer that
 * was passed to them: a later allocation at the same address must not inherit a stale alias. */
int wab_vec_forget_host_buffers(WabVec *h);
int wab_vec_reset_host(WabVec *h, uint8_t *h_grids, uint8_t *h_food, uint8_t *h_role, uint8_t *h_status,
                       void *stream);

/* Episode statistics accumulated on the device since create (or the last call with clear != 0):
 * copies int64[8] (WAB_STAT_*) to h_out8 and synchronises the stream. d_out8 variant: enqueue a
 * device-to-device copy only (for an NCCL all-reduce by the caller), no sync. */
int wab_vec_stats(WabVec *h, int64_t *h_out8, int32_t clear, void *stream);
int wab_vec_stats_device(WabVec *h, int64_t *d_out8, void *stream);

/* Hidden state for differential tests (synchronises). Any pointer may be NULL.
 * x,y i32[N]; food f64[N] (INT mode: counter / turns_to_empty_food); role,status,turn i32[N];
 * episode i64[N]; n_wolves i32[N]; wolves_xy i32[N][wolf_cap][2]; bush_mask u32[N][4];
 * n_log i32[N]; log i32[N][log_cap][3] (x, y, eats). */
int wab_vec_export_state(WabVec *h, int32_t *x, int32_t *y, double *food, int32_t *role, int32_t *status,
                         int32_t *turn, int64_t *episode, int32_t *n_wolves, int32_t *wolves_xy,
                         uint32_t *bush_mask, int32_t *n_log, int32_t *log_xyc, void *stream);

int64_t wab_vec_num_envs(const WabVec *h);
/* Lanes cooperating on one env (1, 4, 8, 16 or 32), chosen at create from the batch size so that a
 * small batch still covers every SM; results do not depend on it. */
int wab_vec_lanes_per_env(const WabVec *h);
/* 1 when a wab_vec_step_many launch of n_steps steps runs as the two-warp pipeline (wab_step_pipe_kernel: one warp of a
 * pair runs the rules, the other publishes observations and scalars), 0 when it runs wab_step_kernel. The pipeline serves
 * lanes-per-env batches of up to 8 rule warps per SM and launches of >= 4 steps; WAB_PIPE=0 / 2 in the environment
 * forces never / always. Results are identical. */
int wab_vec_step_many_pipelined(const WabVec *h, int32_t n_steps);
/* Which kernels serve this handle: 0 = the specialised 11 x 11 / spawn-margin-1 kernels (a 121-bit window sliding in
 * registers), 1 = the warp-per-env kernels for every other odd viewport up to 31 x 31 and margins 1, 2 (wab_generic.cuh;
 * WAB_GENERIC=1 in the environment forces them onto the default geometry, for tests). Results are identical. */
int wab_vec_kernel_kind(const WabVec *h);
void wab_vec_destroy(WabVec *h);

/* ---- next row of the path: the reference's PragmaticObsWrapper (wab_env.py:670-824), the observation
 * its actor-critic actually consumes (actor_critic.py:42, :188). 28 bytes per env:
 * nearest_wolf[4] second_wolf[4] n_wolves[4] nearest_bush[4] second_bush[4] n_bushes[4] (each
 * [up, right, down, left]) standing_on_bush food role status. */
#define WAB_FEATURE_BYTES 28
/* Bind (or unbind with NULL) a device buffer u8[N][28] ([T][N][28] for step_many): every later
 * reset/step also writes the features of the observation it returns, fused into the same kernel. */
int wab_vec_bind_features(WabVec *h, uint8_t *d_features);
/* The same features for an arbitrary observation batch already in HBM (replaces
 * PragmaticObsWrapper.observation, wab_env.py:726-761). */
int wab_pragmatic_features(const uint8_t *d_grids, const uint8_t *d_food, const uint8_t *d_role,
                           const uint8_t *d_status, int64_t n, uint8_t *d_features, void *stream);
/* gym.spaces.flatten of the wrapper observation (actor_critic.py:188): f32[n_rows][wab_vec_flat_dim]
 * one-hot vectors (449 for default options) including the role's view mask. */
int wab_vec_flatten_features(WabVec *h, const uint8_t *d_features, int64_t n_rows, float *d_out, void *stream);
int wab_vec_flat_dim(const WabVec *h);
/* The policy input of actor_critic.py:188-189 in one pass: the same one-hot rows plus noise_scale * U[0,1) per element
 * (the reference adds np.random.rand(...) / 100), written as f32 (out_bf16 = 0) or bf16 (1). The draws are keyed by
 * the handle's seed, the element index and the 64-bit value at d_counter (device memory, may be NULL = 0), which
 * the caller advances between calls — so a captured CUDA graph gets fresh noise on every replay. */
int wab_vec_flatten_features_noisy(WabVec *h, const uint8_t *d_features, int64_t n_rows, void *d_out,
                                   int32_t out_bf16, float noise_scale, const uint64_t *d_counter, void *stream);

/* Categorical(probs).sample() of the reference's select_action (actor_critic.py:117-120) for n rows of n_actions <= 8
 * probabilities (f32, or bf16 when probs_bf16 = 1), one u8 action per row, on the device: inverse CDF on one
 * uniform per row keyed by (seed, row, *d_counter) — d_counter as in wab_vec_flatten_features_noisy. */
int wab_sample_categorical(const void *d_probs, int32_t probs_bf16, int64_t n, int32_t n_actions, uint64_t seed,
                           const uint64_t *d_counter, uint8_t *d_actions, void *stream);

/* ---- the reference's egocentric observation family (wab_env.py:637-667, WolvesAndBushesEnvEgoCentric :930-958):
 * for the five squares the ostrich can reach next (up, right, down, left, stay) the proximity
 * clip(max_distance - taxicab distance, 0, max_distance), max_distance = 11, of the nearest wolf and of the nearest
 * bush with food among every cell seen this episode. wab_vec_enable_ego (before the first reset) makes the step
 * kernels keep each episode's position history (4 bytes per env-step); wab_vec_ego_proximities writes
 * d_out10 u8[N][10] = wolves[5], bushes[5] for the state left by the last reset/step. Envs stepped past max_turns
 * without a reset keep only their first max_turns positions. */
int wab_vec_enable_ego(WabVec *h);
int wab_vec_ego_proximities(WabVec *h, uint8_t *d_out10, void *stream);

/* The first layer of the reference's Policy on the tensor cores (actor_critic.py:59, :88-90, :188-189), fp32 accuracy:
 * d_out f32[n_rows][128] = leaky_relu(affine1(flatten(obs) + noise_scale * U[0,1))) straight from the 28 feature bytes per
 * row (d_features as in wab_vec_flatten_features) — the 449-wide input is generated inside the kernel, with the SAME keyed
 * noise as wab_vec_flatten_features_noisy for the same *d_counter, and never written to memory. One tcgen05 kernel
 * (bf16 x 3 operand splits, fp32 accumulation in tensor memory; the result equals the fp32 GEMM to rounding).
 * wab_policy_affine1_prepare packs affine1.weight f32[128][in_dim] (in_dim = wab_vec_flat_dim, at most 512) into
 * d_packed (wab_policy_affine1_packed_bytes() bytes, 16-byte aligned) — once per weight update; d_bias f32[128]. */
int64_t wab_policy_affine1_packed_bytes(void);
int wab_policy_affine1_prepare(const float *d_weight, int32_t in_dim, void *d_packed, void *stream);
int wab_policy_affine1(WabVec *h, const uint8_t *d_features, int64_t n_rows, const void *d_packed, const float *d_bias,
                       float noise_scale, float leaky_slope, const uint64_t *d_counter, float *d_out, void *stream);

/* The whole trunk of the reference's Policy.forward (actor_critic.py:88-92) in the same tcgen05 kernel, fp32 accuracy:
 * d_z3 f32[n_rows][128] = affine3(leaky_relu(affine2(leaky_relu(affine1(flatten(obs) + noise))))) — the PRE-activation
 * output of affine3, which is what wab_policy_tail takes. The hidden activations never leave the SM: accumulators ->
 * registers (bias, leaky-ReLU, three-way bf16 split) -> shared memory as the next layer's operand. d_packed1 / d_bias1 as
 * for wab_policy_affine1; d_packed2 = wab_policy_linear_prepare(affine2.weight f32[150][128]), d_packed3 =
 * wab_policy_linear_prepare(affine3.weight f32[128][150]) (wab_policy_linear_packed_bytes(n_out, n_in) bytes each,
 * 16-byte aligned; once per weight update); hidden2 must be 150. */
int64_t wab_policy_linear_packed_bytes(int32_t n_out, int32_t n_in);
int wab_policy_linear_prepare(const float *d_weight, int32_t n_out, int32_t n_in, void *d_packed, void *stream);
int wab_policy_trunk(WabVec *h, const uint8_t *d_features, int64_t n_rows, const void *d_packed1, const float *d_bias1,
                     const void *d_packed2, const float *d_bias2, int32_t hidden2, const void *d_packed3,
                     const float *d_bias3, float noise_scale, float leaky_slope, const uint64_t *d_counter, float *d_z3,
                     void *stream);

/* Policy.forward + select_action (actor_critic.py:84-97, :108-125) for n_rows environments as ONE kernel: wab_policy_trunk
 * followed, in the same launch, by what wab_policy_tail does (x = clamp(leaky_relu(z3), lo, hi), both heads, softmax, one
 * action per row by inverse CDF on a uniform keyed by (sample_seed, row, *d_counter)). Arguments as for those two calls;
 * d_value, d_probs, d_logp and d_z3 may be NULL. */
int wab_policy_forward(WabVec *h, const uint8_t *d_features, int64_t n_rows, const void *d_packed1, const float *d_bias1,
                       const void *d_packed2, const float *d_bias2, int32_t hidden2, const void *d_packed3,
                       const float *d_bias3, const float *d_w_heads, const float *d_b_heads, int32_t n_actions,
                       float noise_scale, float leaky_slope, float clamp_lo, float clamp_hi, uint64_t sample_seed,
                       const uint64_t *d_counter, uint8_t *d_actions, float *d_value, float *d_probs, float *d_logp,
                       float *d_z3, void *stream);

/* The tail of the reference's Policy.forward + select_action (actor_critic.py:84-97, :108-125) for n_rows rows in one
 * pass, fp32: d_z3 f32[n_rows][128] is the PRE-activation output of affine3; x = clamp(leaky_relu(z3), lo, hi);
 * logits = W[0..A) x + b, value = W[A] x + b[A] (d_w_heads f32[A + 1][128] = action_head.weight stacked on
 * value_head.weight, d_b_heads f32[A + 1]); probs = softmax(logits); one action per row by inverse CDF on a uniform
 * keyed by (seed, row, *d_counter). Outputs: d_actions u8[n_rows]; d_value f32[n_rows], d_probs f32[n_rows][A],
 * d_logp f32[n_rows] (log-probability of the sampled action) — each may be NULL. */
int wab_policy_tail(const float *d_z3, int32_t hidden, const float *d_w_heads, const float *d_b_heads, int64_t n_rows,
                    int32_t n_actions, float leaky_slope, float clamp_lo, float clamp_hi, uint64_t seed,
                    const uint64_t *d_counter, uint8_t *d_actions, float *d_value, float *d_probs, float *d_logp,
                    void *stream);

/* ---- Environment 2.0 ("/root/reference/Environment 2.0"): a toroidal W x H world of ostriches, wolves and
 * bushes per environment. One call = one world turn: every entity, in id order (ostriches, wolves, bushes),
 * observes and acts exactly as the reference driver loop does (Env2Tests.py:46-88: get_obs(i), take_action(i)). */
typedef struct Wab2Config {
    int32_t abi_version;
    int32_t width, height;                       /* World(width, height), World.py:141                              */
    int32_t n_ostriches, n_wolves, n_bushes;     /* create_ostriches / wolves / bushes, WAB_Environment2.py:61-110  */
    int32_t lookout_view_radius, gatherer_view_radius, wolf_view_radius;   /* WAB_Environment2.py:35-36, :49       */
    int32_t window_radius;                       /* R of the (2R+1)^2 observation window (>= the largest radius)    */
    int32_t starting_role;                       /* :19                                                             */
    int32_t ostrich_starting_food;               /* :32 (integer-valued in the reference's arithmetic)              */
    int32_t wolf_starting_food, wolf_food_for_eating_ostrich;   /* :43-44                                           */
    int32_t food_per_bush, food_given_per_turn;  /* :28-29, Bush.take_food Bush.py:31-39                            */
} Wab2Config;
typedef struct Wab2World Wab2World;
/* World.__init__ + create_* for n_envs worlds (keyed spawn positions). Synchronises. */
int wab2_create(const Wab2Config *cfg, int64_t n_envs, uint64_t seed, uint64_t env_id_base, int32_t device, Wab2World **out);
/* reset_environment (WAB_Environment2.py:113-118) of every world. */
int wab2_reset(Wab2World *h, void *stream);
/* One world turn. d_actions u8[A][N] is ENTITY-MAJOR, A = n_ostriches + n_wolves (bushes act with 0). Outputs per
 * acting entity, in the handle's output layout (wab2_output_layout): 0 = entity-major, d_planes u8[A][N][3][2R+1][2R+1]
 * (ostriches, wolves, bushes listed by get_observations at [dx+R][dy+R]; may be NULL), d_internal i32[A][N][5]
 * (x, y, food, role | is_running, status; may be NULL), d_reward f32[A][N], d_done u8[A][N] — the worlds of a warp
 * are contiguous; 1 = world-major, the same arrays with the first two dimensions exchanged ([N][A]...) — the A
 * windows of a world are one contiguous run for the warp that owns it. The observation of entity i is taken right
 * before it acts. */
int wab2_turn(Wab2World *h, const uint8_t *d_actions, uint8_t *d_planes, int32_t *d_internal, float *d_reward,
              uint8_t *d_done, void *stream);
/* Hidden state for tests: out9 i32[N][E][9] = type, x, y, table X, table Y, Visible, food, role, status. Synchronises. */
int wab2_export_state(Wab2World *h, int32_t *out9, int32_t *turn, void *stream);
/* The inverse of wab2_export_state (tests: the reference's own known-answer worlds, World_tests.py:5-88, are built
 * from explicit positions): in9 i32[N][E][9] in the export layout (the type column must match the handle's entity
 * order: ostriches, wolves, bushes); turn i32[N] or NULL (unchanged). Synchronises. */
int wab2_import_state(Wab2World *h, const int32_t *in9, const int32_t *turn, void *stream);
/* Which kernel serves this handle: 0 = wab2_turn_kernel (one thread per world), 1 = wab2_grid_turn_kernel (one
 * warp per world with occupancy planes, for worlds every observation window fits once). Results are identical. */
int wab2_kernel_kind(const Wab2World *h);
/* Layout of wab2_turn's outputs for this handle: 0 = entity-major [A][N]..., 1 = world-major [N][A]... (the
 * warp-per-world kernel). */
int wab2_output_layout(const Wab2World *h);
void wab2_destroy(Wab2World *h);

/* Raw Philox4x32-10 on the device for n counters (cross-checks the RNG contract). d_ctr u32[n][4],
 * d_out u32[n][4]. */
int wab_philox_device(const uint32_t *d_ctr, uint32_t key0, uint32_t key1, int64_t n, uint32_t *d_out,
                      void *stream);

const char *wab_last_error(void);
int wab_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
