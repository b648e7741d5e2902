/*
 * cm3env.h - C ABI of the B200 batched environment stepper (libcm3env.so).
 *
 * The reference (011235813/cm3) has no FFI: its "plugin API" for this path is two Python
 * object protocols, Checkers.reset/step (env/checkers.py:265,228) and
 * MultiAgentEnv.reset/step + the scenario callbacks (multiagent/environment.py:125,81;
 * multiagent/scenarios/multi-goal_spread.py:19-154).  The Python facades in cm3_b200/
 * keep those protocols; everything they compute goes through the entry points below,
 * which are what a ctypes / cgo / JNI binding on the reference side would bind
 * (INTEGRATION.md shows the ctypes stub).
 *
 * Conventions
 *  - plain C types only; every function returns a cm3_status (0 = ok, < 0 = error) and
 *    leaves a message retrievable with cm3_last_error() (thread local);
 *  - the library owns no HBM: state, action and output buffers are allocated by the
 *    caller (the facades hold them in torch tensors) and passed as raw device pointers;
 *    they must outlive the stream work.  A handle holds only configuration;
 *  - all device work is enqueued on the caller's stream (cudaStream_t passed as void*,
 *    NULL = legacy default stream); nothing synchronises except the *_host entry points;
 *  - there is NO CPU implementation behind this ABI: without a CUDA device every compute
 *    entry point fails with CM3_ERR_NO_DEVICE;
 *  - a handle is used by one host thread at a time; handles are independent.
 *
 * Batched layout ("B" = num_envs on this device, "N" = n_agents, Real = float or double
 * as selected by cfg.real):
 *  every output field is its own dense array [B][...] (or [T][B][...] for rollouts), so a
 *  learner can consume a field directly as a batch tensor.
 */
#ifndef CM3ENV_H
#define CM3ENV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM3_ABI_VERSION 2
#define CM3_MAX_AGENTS 8
#define CM3_MAX_DST 8 /* destination buffer sets of *_rollout_gather (GPUs of one NVSwitch node) */

typedef enum {
    CM3_OK = 0,
    CM3_ERR_BAD_ARG = -1,     /* NULL handle / pointer, bad enum, bad mask */
    CM3_ERR_BAD_SHAPE = -2,   /* geometry the reference itself rejects (checkers.py:16-17) */
    CM3_ERR_CUDA = -3,        /* a CUDA runtime call failed; message has the CUDA error */
    CM3_ERR_UNSUPPORTED = -4, /* geometry / agent count outside what the kernels address */
    CM3_ERR_NO_DEVICE = -5,   /* no CUDA device: there is no CPU fallback */
    CM3_ERR_NCCL = -6         /* cm3_comm_*: NCCL missing or an NCCL call failed */
} cm3_status;

typedef enum { CM3_REAL_F32 = 0, CM3_REAL_F64 = 1 } cm3_real;
/* Element type of the two bulky Checkers outputs (grid, obs_self_t), whose values are always in
 * {-1, 0, +1}: CM3_TILE_REAL writes them as Real like the reference's float arrays, CM3_TILE_I8
 * as signed bytes - the same numbers in a quarter of the bytes (HBM, NVLink and PCIe traffic). */
/* CM3_TILE_U2 packs them 2 bits per cell (0 -> 0, 1 -> +1, 3 -> -1), least significant cell first, in
 * 32-bit words - 1/16 of the float bytes:
 *   obs_self_t [B][N][W][RW] uint32   one window row (cell (dc, ch) at bits 2 (3 dc + ch)) per
 *                                     RW = ceil(6 W / 32) words, W = 2 n_obs + 1
 *   grid       [B][R][GW]    uint32   one grid row (cell (j, ch) at bits 2 (2 j + ch)), 8 cells per word,
 *                                     GW = ceil((n_columns + 1) / 8)
 * (cm3_b200/tiles.py holds the decoders.) */
typedef enum { CM3_TILE_REAL = 0, CM3_TILE_I8 = 1, CM3_TILE_U2 = 2 } cm3_tile;

int cm3_abi_version(void);
const char *cm3_last_error(void);
/* number of visible CUDA devices (0 when the driver is absent) */
int cm3_device_count(int *count);
/* waits for the work enqueued on `stream` (cudaStreamSynchronize): for bindings without a CUDA
 * runtime of their own.  Together with pinned host memory - which the device addresses directly
 * under unified addressing - this is the low-latency B = 1 path: pass pinned host pointers as
 * `actions` and `outs` of cm3_*_step and call cm3_stream_synchronize; the kernel reads the actions
 * from and writes the outputs to host memory itself, with no copy calls on the host path. */
int cm3_stream_synchronize(void *stream);

/* ------------------------------------------------------------------ Checkers */

/* Mirrors Checkers.__init__(n_rows, n_columns, n_obs, agents_r, agents_c, n_agents,
 * max_steps) - env/checkers.py:5-35.  agents_r / agents_c are BEFORE the n_obs expansion,
 * exactly as the reference takes them. */
typedef struct {
    int32_t n_rows, n_columns, n_obs, n_agents, max_steps;
    int32_t agents_r[CM3_MAX_AGENTS];
    int32_t agents_c[CM3_MAX_AGENTS];
    int32_t num_envs;      /* B on this device */
    int32_t real;          /* cm3_real of the float outputs */
    int32_t device;        /* CUDA ordinal */
    int32_t tile;          /* cm3_tile of grid / obs_self_t (CM3_TILE_I8 / CM3_TILE_U2 need real == F32) */
    int64_t env_id_offset; /* global id of local env 0 (keys the Philox streams, so results
                              do not depend on how the batch is sharded over GPUs) */
    int32_t random_goal;   /* n_agents == 1 only: 1 = an in-kernel episode reset (auto_reset) draws
                              the new episode's goal uniformly from {0, 1} like the trainer does
                              before every episode (alg/train_offpolicy.py:291-296), from Philox keyed
                              by (seed; global env id, step); 0 = the goal is kept */
    int32_t reserved;
} cm3_checkers_config;

/* Any board the reference's constructor accepts (n_rows odd, n_columns even, checkers.py:16-17) is
 * accepted as DATA as long as the bitboards can address it: n_rows * n_columns <= 64,
 * n_columns + 2 * n_obs + 1 <= 32, 1 <= n_obs <= 3, n_rows + 2 * n_obs <= 16, n_agents <= 8.  The
 * boards of the reference's configs (3 x 8 and 3 x 16, n_obs 2) additionally have kernels with the
 * geometry compiled in. */

/* Compact per-env state (device pointers, caller-owned):
 *   remaining[b]  bit i*n_columns+j set <=> valid-grid cell (i,j) still holds its reward
 *                 (the green/orange colour is a function of (i+j) parity, checkers.py:54-63)
 *   agents[b][n]  r | c<<8 | n_green<<16 | n_orange<<24, (r,c) in expanded coordinates
 *   meta[b]       steps (bits 0-23) | goal index of agent n at bit 24+n */
typedef struct {
    uint64_t *remaining; /* [B]    */
    uint32_t *agents;    /* [B][N] */
    uint32_t *meta;      /* [B]    */
    uint32_t *sync;      /* [2][cm3_checkers_tiles(h)] launch-chaining words, zero-initialised by the
                            caller and otherwise opaque; NULL = launches are ordered by the stream
                            alone and *_step_chained is refused (see cm3_checkers_step_chained) */
} cm3_checkers_state;

/* Outputs of reset/step (env/checkers.py:262,291), dense per field.  For rollouts every
 * pointer addresses [T][B][...].  Any pointer may be NULL: that field is not written. */
typedef struct {
    void *grid;          /* [B][n_rows][n_columns+1][2]   Tile  get_valid_grid  :66-76  */
    void *vec;           /* [B][N][4]                      Real  get_global_state :89-93 */
    void *obs_others;    /* [B][N][2*max(N-1,1)]           Real  :143-151 */
    void *obs_self_t;    /* [B][N][2*n_obs+1][2*n_obs+1][3] Tile get_obs :97-109 */
    void *obs_self_v;    /* [B][N][4]                      Real  :137-139 */
    void *reward;        /* [B]                            Real  np.sum(local_rewards) :243 */
    void *local_rewards; /* [B][N]                         Real  :232-237 */
    uint8_t *done;       /* [B]                                  :246-260 */
    uint8_t *goal_idx;   /* [B][N] goal index of every agent in the episode the written observations
                            belong to (the trainers carry `goals` beside every transition,
                            train_offpolicy.py:291-298; with cfg.random_goal it changes at in-kernel
                            resets).  Optional like every field. */
} cm3_checkers_outputs;

typedef struct cm3_checkers_s *cm3_checkers_t;

int cm3_checkers_create(const cm3_checkers_config *cfg, cm3_checkers_t *out);
int cm3_checkers_destroy(cm3_checkers_t h);
/* number of env tiles (one warp each) a launch of this handle is cut into: the length of one row of
 * cm3_checkers_state.sync */
int cm3_checkers_tiles(cm3_checkers_t h, int32_t *tiles);

/* Checkers.reset(goals) for the envs selected by env_mask (NULL = all).  goal_idx is
 * [B][N] uint8 on the device, goal_idx[b][n] = argmax(goals[n]) in {0,1} (checkers.py:235);
 * NULL = every selected env keeps the goals it already has (zeroed state: goal 0 for everyone).
 * Observations of ALL envs are written to outs (fresh for reset envs, current otherwise);
 * reward / local_rewards are not touched, done is written as 0 for the reset envs. */
int cm3_checkers_reset(cm3_checkers_t h, const cm3_checkers_state *st, const uint8_t *goal_idx,
                       const uint8_t *env_mask, const cm3_checkers_outputs *outs, void *stream);

/* Checkers.step(actions): actions [B][N] int8 on the device.  Values outside 0..4 are
 * legal and earn the -0.1 penalty (checkers.py:184-186). */
int cm3_checkers_step(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                      const cm3_checkers_outputs *outs, void *stream);

/* T fused steps in one launch; state stays in registers between steps.
 *   actions      [T][B][N] int8 device, or NULL: uniform actions in {0..4} are drawn on the
 *                device from Philox4x32-10 keyed by (seed; global env id, t0 + t)
 *   actions_out  [T][B][N] int8 device or NULL: the actions that were applied
 *   auto_reset   0: reference behaviour (keeps stepping past done)
 *                1: an env whose step returned done is reset inside the kernel (same goals,
 *                   or a fresh goal when cfg.random_goal is set)
 *                   and the observations written for that step are those of the fresh
 *                   episode; reward / done still describe the terminal transition
 *   outs         [T][B][...] per field */
int cm3_checkers_rollout(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                         uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset,
                         int8_t *actions_out, const cm3_checkers_outputs *outs, void *stream);

/* One step that may OVERLAP the previous launch on this state (policy-free inner loops, e.g. a
 * pre-generated or double-buffered action stream; alg/train_onpolicy.py:302-350 with the policy's
 * actions already in HBM).  Consecutive launches on one state depend on each other tile by tile -
 * tile i of step k + 1 needs tile i of step k - and that is all a chained launch waits for (ticket
 * words in st->sync, acquire / release), instead of the completion of the whole previous grid as
 * stream order would have it: the store phase of step k overlaps the compute phase of step k + 1.
 * Contract: (1) st->sync non-NULL; (2) `actions` was complete in memory before the PREVIOUS launch
 * on this stream was enqueued, or is written by a kernel that sits between the two step launches
 * in the stream (that kernel then orders everything, as usual); (3) `outs` does not alias the
 * outs of the previous TWO launches (use a ring of >= 3 slots): a launch releases its tile as soon
 * as the compact state is stored, possibly before its own outputs have left the SM, so the next
 * launch's outputs - and the one after that, which only waits for the next launch's state - may
 * be on their way at the same time.  auto_reset as in cm3_checkers_rollout
 * (seed / t0 key the goal redraw when cfg.random_goal is set). */
int cm3_checkers_step_chained(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                              uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_checkers_outputs *outs, void *stream);

/* Fused rollout + all-gather.  Same computation as cm3_checkers_rollout, but every output element
 * is stored to n_dst destination buffer sets instead of one.  Each set is laid out
 * [T][dst_B][...] per field and this handle's B envs occupy rows [dst_env0, dst_env0 + B).
 * With dsts[] = the rollout buffers of all ranks of a node (peer device pointers obtained through
 * CUDA IPC / symmetric memory; NVLink 5 + NVSwitch gives every GPU a direct store path to every
 * peer) and dst_env0 = this rank's first global env id, the gathered [T][total_envs][...] batch
 * materialises on every GPU straight from the step kernel's stores: the observation tile is
 * expanded once in shared memory and leaves the SM as n_dst TMA bulk stores, with no staging pass
 * through local HBM and no separate collective.  The reference has no counterpart (one env per
 * process, alg/train_multiprocess.py:31-43); it stands where a data-parallel learner would
 * all-gather its workers' rollouts.  All sets must have the same NULL fields.  Making the stores
 * visible to the peers (a barrier after the stream work) is the caller's job. */
int cm3_checkers_rollout_gather(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                                uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset,
                                int8_t *actions_out, int32_t n_dst, const cm3_checkers_outputs *dsts,
                                int64_t dst_B, int64_t dst_env0, void *stream);

/* Host-buffer form of step: copies actions_host -> actions_dev, steps, copies every
 * non-NULL field of outs_host back from the matching field of outs_dev and waits for the
 * stream.  This is the call a host-side binding uses when its buffers live in host memory
 * (pinned memory makes the copies asynchronous with respect to other streams). */
int cm3_checkers_step_host(cm3_checkers_t h, const cm3_checkers_state *st,
                           const int8_t *actions_host, int8_t *actions_dev,
                           const cm3_checkers_outputs *outs_dev,
                           const cm3_checkers_outputs *outs_host, void *stream);

/* State transfer for bindings without a tensor library (checkpoint / resume, injecting a
 * reference state for parity): copies the compact state between the caller's device arrays `dev`
 * and host arrays `host` of the same shapes.  NULL fields of `host` are skipped.  Enqueued on
 * `stream`; get_state waits for the stream so that the host arrays are filled on return.
 * The reference keeps no such thing (its state is the Python object itself, env/checkers.py:26-35);
 * this replaces pickling an env. */
int cm3_checkers_get_state(cm3_checkers_t h, const cm3_checkers_state *dev, const cm3_checkers_state *host,
                           void *stream);
int cm3_checkers_set_state(cm3_checkers_t h, const cm3_checkers_state *dev, const cm3_checkers_state *host,
                           void *stream);

/* step_host for a caller that keeps every field of outs_dev inside ONE device allocation
 * [dev_block, dev_block + block_bytes) and wants the same bytes, at the same offsets, in
 * host_block: the outputs travel as one device-to-host copy instead of one per field (the call is
 * PCIe-bound; per-copy set-up time is pure overhead).  Fields of outs_dev outside the block are
 * rejected with CM3_ERR_BAD_ARG. */
int cm3_checkers_step_host_packed(cm3_checkers_t h, const cm3_checkers_state *st,
                                  const int8_t *actions_host, int8_t *actions_dev,
                                  const cm3_checkers_outputs *outs_dev, const void *dev_block,
                                  void *host_block, size_t block_bytes, void *stream);

/* Host-buffer rollout, double buffered: T steps whose actions come from host memory and whose
 * outputs all land in host memory, with the device-to-host copy of step t overlapping the kernel of
 * step t + 1.  Per step t: copy actions_host[t] ([B][N]) to actions_dev[t & 1], step (one launch,
 * in-kernel episode reset if auto_reset), then copy the packed output block dev_blocks[t & 1]
 * (block_bytes; outs_dev[t & 1] are the field pointers inside it, see *_step_host_packed) to
 * host_blocks + t * host_stride on a second stream owned by the handle.  Returns after the last copy
 * has landed.  Pinned host memory makes the copies asynchronous.  This is the loop a host-side
 * trainer with a pre-drawn action sequence runs (the trainers' pre-training phase draws
 * np.random.randint actions, alg/train_onpolicy.py:305-307), at the PCIe rate of the outputs instead
 * of copy + kernel + copy in series. */
int cm3_checkers_rollout_host(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions_host,
                              int8_t *actions_dev, int32_t T, uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_checkers_outputs *outs_dev, const void *const *dev_blocks,
                              void *host_blocks, size_t block_bytes, size_t host_stride, void *stream);

/* ------------------------------------------------------------------ Particle */

/* World / scenario constants (defaults = the reference's, filled by
 * cm3_particle_default_config) and the multi-goal_spread presets. */
typedef struct {
    int32_t n_agents;  /* agents == landmarks, multi-goal_spread.py:36-37 */
    int32_t max_steps; /* MultiAgentEnv(max_steps=...), environment.py:16 */
    int32_t num_envs;
    int32_t real;      /* cm3_real of state AND outputs (F64 = free-running parity mode) */
    int32_t device;
    int32_t reserved;
    int64_t env_id_offset;
    double dt;             /* core.py:94   */
    double damping;        /* core.py:96   */
    double contact_force;  /* core.py:98   */
    double contact_margin; /* core.py:99   */
    double agent_size;     /* multi-goal_spread.py:47 */
    double mass;           /* core.py:47-51 */
    double sensitivity;    /* environment.py:211 */
    double reach_thresh;   /* multi-goal_spread.py:126 */
    /* reset presets, multi-goal_spread.py:29-35 and alg/config_particle_*.json */
    double agents_x[CM3_MAX_AGENTS], agents_y[CM3_MAX_AGENTS];
    double landmarks_x[CM3_MAX_AGENTS], landmarks_y[CM3_MAX_AGENTS];
    double initial_std;
    double prob_random;
    /* float kernels only: an agent pair whose contact force is provably smaller than this (in force
     * units; the action force is `sensitivity` = 5) is not evaluated.  The reference evaluates the
     * softplus penetration for every pair at every distance (core.py:143-155); its tail decays as
     * contact_force * contact_margin * exp(-(dist - dist_min) / contact_margin), i.e. below 1e-9
     * beyond dist_min + 18.4 contact_margin - four orders of magnitude under float32 resolution of
     * the quantities it is added to.  Default 0 (cm3_particle_default_config) = skip a pair only where
     * its force is exactly zero in float arithmetic, which keeps the float kernel identical to its own
     * literal evaluation; 1e-9 measured within noise of that on every workload (warps, not lanes,
     * skip the evaluation: profiles/r02c_ab.txt).  The double kernels always use the exact criterion. */
    double contact_cutoff;
} cm3_particle_config;

void cm3_particle_default_config(cm3_particle_config *cfg, int32_t n_agents, int32_t max_steps);

/* State (device, caller-owned).  sv rows are (vel.x, vel.y, pos.x, pos.y) - the same
 * layout as the reference's global_state rows (environment.py:113-116). */
typedef struct {
    void *sv;            /* [B][N][4] Real */
    void *landmarks;     /* [B][N][2] Real  landmark.state.p_pos */
    int32_t *steps;      /* [B]  env.steps */
    int32_t *collisions; /* [B]  scenario.collisions (episode total, double counted :135-137) */
    uint8_t *reached;    /* [B]  bit n = agents[n].reached */
    uint32_t *sync;      /* [2][cm3_particle_tiles(h)], see cm3_checkers_state.sync */
} cm3_particle_state;

typedef struct {
    void *global_state; /* [B][N][4]              Real environment.py:113-116 */
    void *obs_others;   /* [B][N][4*max(N-1,1)]   Real multi-goal_spread.py:146-154 */
    void *obs_self;     /* [B][N][4]              Real */
    void *reward;       /* [B]                    Real np.sum(reward_n), environment.py:107 */
    void *reward_n;     /* [B][N]                 Real multi-goal_spread.py:121-138 */
    uint8_t *done;      /* [B]                         environment.py:118-121 */
    int32_t *collisions; /* [B] scenario.collisions after this step, BEFORE an in-kernel reset zeroes
                            it: at a step with done = 1 it is the finished episode's total, which is
                            what the trainer's good / bad episode split reads
                            (alg/train_onpolicy.py:356).  Optional like every field. */
    uint8_t *reached;    /* [B] bit n = agents[n].reached after this step (multi-goal_spread.py:126-129),
                            what Scenario.done(agent) returns (:140-143); before an in-kernel reset
                            clears it.  Optional. */
} cm3_particle_outputs;

typedef struct cm3_particle_s *cm3_particle_t;

int cm3_particle_create(const cm3_particle_config *cfg, cm3_particle_t *out);
int cm3_particle_destroy(cm3_particle_t h);
int cm3_particle_tiles(cm3_particle_t h, int32_t *tiles);

/* MultiAgentEnv.reset() for the envs selected by env_mask (NULL = all).
 *   init_pos / init_landmarks  [B][N][2] Real device: the state reset_world() produced on the
 *       host (parity protocol: inject the reference's draws).  If NULL the device draws them:
 *       multi-goal_spread.py:75-91 re-expressed on Philox4x32-10 keyed by
 *       (seed; global env id, reset_counter) - u < prob_random -> uniform(-1,1) agents and
 *       landmarks, else presets + N(0, initial_std) on the agents.
 * Velocities, steps, collisions and reached are zeroed; observations of all envs are written
 * to outs (reward fields untouched, done = 0 for the reset envs). */
int cm3_particle_reset(cm3_particle_t h, const cm3_particle_state *st, const void *init_pos,
                       const void *init_landmarks, const uint8_t *env_mask, uint64_t seed,
                       int64_t reset_counter, const cm3_particle_outputs *outs, void *stream);

/* MultiAgentEnv.step(action_n): actions [B][N] int8 device; values outside 1..4 apply no
 * force (environment.py:197-200). */
int cm3_particle_step(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                      const cm3_particle_outputs *outs, void *stream);

/* T fused steps; same contract as cm3_checkers_rollout.  With auto_reset the fresh episode is
 * drawn as in cm3_particle_reset, keyed by (seed; global env id, t0 + t + 1) in a counter domain of
 * its own (explicit resets and in-kernel resets never share a Philox counter). */
int cm3_particle_rollout(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                         uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset,
                         int8_t *actions_out, const cm3_particle_outputs *outs, void *stream);

/* see cm3_checkers_step_chained; seed / t0 key the in-kernel reset draws */
int cm3_particle_step_chained(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                              uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_particle_outputs *outs, void *stream);

/* Fused rollout + all-gather; see cm3_checkers_rollout_gather. */
int cm3_particle_rollout_gather(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                                uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset,
                                int8_t *actions_out, int32_t n_dst, const cm3_particle_outputs *dsts,
                                int64_t dst_B, int64_t dst_env0, void *stream);

int cm3_particle_step_host(cm3_particle_t h, const cm3_particle_state *st,
                           const int8_t *actions_host, int8_t *actions_dev,
                           const cm3_particle_outputs *outs_dev,
                           const cm3_particle_outputs *outs_host, void *stream);

/* see cm3_checkers_get_state / cm3_checkers_set_state (world.agents[i].state, landmarks, env.steps,
 * scenario.collisions, agent.reached of the reference objects) */
int cm3_particle_get_state(cm3_particle_t h, const cm3_particle_state *dev, const cm3_particle_state *host,
                           void *stream);
int cm3_particle_set_state(cm3_particle_t h, const cm3_particle_state *dev, const cm3_particle_state *host,
                           void *stream);

/* see cm3_checkers_step_host_packed */
int cm3_particle_step_host_packed(cm3_particle_t h, const cm3_particle_state *st,
                                  const int8_t *actions_host, int8_t *actions_dev,
                                  const cm3_particle_outputs *outs_dev, const void *dev_block,
                                  void *host_block, size_t block_bytes, void *stream);

/* see cm3_checkers_rollout_host */
int cm3_particle_rollout_host(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions_host,
                              int8_t *actions_dev, int32_t T, uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_particle_outputs *outs_dev, const void *const *dev_blocks,
                              void *host_blocks, size_t block_bytes, size_t host_stride, void *stream);

/* ------------------------------------------------------------------ rollout exchange without torch */

/* The all-gather of rollout buffers (BASELINE configs[3]) for bindings that have no
 * torch.distributed: one process per GPU, NCCL underneath (resolved with dlopen at first use, so the
 * library has no link-time NCCL dependency).  Rank 0 calls cm3_comm_unique_id and hands the 128 bytes
 * to every other rank by its own means; all ranks then call cm3_comm_init with the same id.
 * cm3_comm_allgather: every rank contributes `bytes_per_rank` bytes from `send`; `recv` receives
 * world * bytes_per_rank bytes, rank r's contribution at offset r * bytes_per_rank (ncclAllGather on
 * `stream`, asynchronous).  The reference has no counterpart (alg/train_multiprocess.py:31-43 runs
 * independent seeds).  The fused alternative - the step kernel storing straight into the peers'
 * buffers - is cm3_*_rollout_gather. */
typedef struct cm3_comm_s *cm3_comm_t;
int cm3_comm_unique_id(uint8_t *id /* [128] */);
int cm3_comm_init(const uint8_t *id /* [128] */, int32_t rank, int32_t world, int32_t device, cm3_comm_t *out);
int cm3_comm_allgather(cm3_comm_t c, const void *send, void *recv, size_t bytes_per_rank, void *stream);
int cm3_comm_destroy(cm3_comm_t c);

#ifdef __cplusplus
}
#endif
#endif /* CM3ENV_H */
