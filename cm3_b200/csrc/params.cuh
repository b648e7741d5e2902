// params.cuh - kernel parameter blocks shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cm3env.h"

namespace cm3 {

constexpr int kCkMaxTR = 16;   // n_rows + 2*n_obs
constexpr int kCkMaxTC = 40;   // n_columns + 2*n_obs + 1
constexpr int kCkMaxCnt = 34;  // n_rows*n_columns/2 + 1

constexpr int kMaxDst = CM3_MAX_DST;  // destination buffer sets of a fused rollout + all-gather

enum CkMode { kCkStep = 0, kCkReset = 1 };
enum PtMode { kPtStep = 0, kPtReset = 1 };

struct CkOut {
    char *grid, *vec, *obs_others, *obs_self_t, *obs_self_v, *reward, *local_rewards;
    uint8_t *done, *goal_idx;
};
struct PtOut {
    char *global_state, *obs_others, *obs_self, *reward, *reward_n;
    uint8_t *done;
    int32_t *collisions;
    uint8_t *reached;
};

struct CkParams {
    uint64_t *remaining;
    uint32_t *agents;
    uint32_t *meta;
    uint32_t *sync;  // per-tile launch-chaining words (common.cuh: TileTicket), may be NULL
    const int8_t *actions;
    const uint8_t *goal_idx;
    const uint8_t *env_mask;
    int8_t *actions_out;
    // Output buffer sets.  out[0] is the only one for reset / step / rollout (out_B = B,
    // out_env0 = 0); rollout_gather stores every element to all n_dst sets, each laid out
    // [T][out_B][...] with this shard at env offset out_env0.  All sets have the same NULL fields.
    CkOut out[kMaxDst];
    int n_dst;
    long long out_B, out_env0;
    int B, T, max_steps, mode, auto_reset;
    int chained;      // 1: wait for this tile's predecessor only (TileTicket) instead of griddepcontrol.wait
    int tpb;          // tiles per block of a chained single-step launch (0 / 1: one; checkers.cu: launch_ck)
    int early;        // when the compact state of the launch's LAST step is written back and the tile published
                      // (params.cuh: chain_early_mode): 0 after the step's outputs, 1 state before / publish after,
                      // 2 both before the outputs are assembled
    int tile0;        // first tile of this launch (a chained step is issued as several partial grids)
    int random_goal;  // N == 1: in-kernel resets redraw the goal (train_offpolicy.py:291-296)
    int R, C, O;      // board geometry for the kernels that take it as data (checkers.cu: DynGeo)
    unsigned long long color_mask[2];  // bitboard of the green / orange cells (checkers.py:54-63)
    unsigned long long seed;
    long long t0, env_id_offset;
    int start_r[CM3_MAX_AGENTS], start_c[CM3_MAX_AGENTS];  // expanded coordinates
    // (r - total_rows/2.0)/total_rows etc., evaluated on the host in float64 exactly as
    // checkers.py:120-121,139 write them, so the device never divides
    double norm_row[kCkMaxTR], norm_col[kCkMaxTC], norm_cnt[kCkMaxCnt];
};

// World constants in the arithmetic type of the kernel.  far2: squared centre distance beyond
// which the contact force is exactly +-0 in that arithmetic; near2: beyond which two agents cannot
// be in collision (particle.cu).
template <typename Real> struct PtConsts {
    Real dt, keep, contact_force, contact_margin, dist_min, mass, sensitivity, neg_reach, far2, near2;
};

// tensor maps of the three tiled particle outputs (particle.cu: swizzled staging + UTMASTG)
struct PtTensorMaps { CUtensorMap gs, os, oo; };

struct PtParams {
    char *sv, *landmarks;
    int32_t *steps, *collisions;
    uint8_t *reached;
    uint32_t *sync;  // see CkParams
    const int8_t *actions;
    int8_t *actions_out;
    const char *init_pos, *init_landmarks;
    const uint8_t *env_mask;
    PtOut out[kMaxDst];  // see CkParams
    int n_dst;
    long long out_B, out_env0;
    int B, T, max_steps, mode, auto_reset;
    int chained, tile0, early, tpb;  // see CkParams
    unsigned long long seed;
    long long t0, env_id_offset, reset_counter;
    double dt, damping, contact_force, contact_margin, dist_min, mass, sensitivity, reach_thresh;
    double dist_min2;  // dist_min * dist_min (particle.cu: Contact<float>)
    double agents_x[CM3_MAX_AGENTS], agents_y[CM3_MAX_AGENTS];
    double landmarks_x[CM3_MAX_AGENTS], landmarks_y[CM3_MAX_AGENTS];
    double initial_std, prob_random;
    PtConsts<float> kf;   // the constants above rounded to the kernel's Real on the host
    PtConsts<double> kd;
    int tma;              // 1: tiles staged swizzled and stored through tm (set by particle_launch)
    PtTensorMaps tm;
};

bool checkers_geometry_supported(int R, int C, int O, int N);
int checkers_tile_envs(int N);  // envs per warp tile of the Checkers kernels
bool dyn_geometry_forced();    // CM3_CK_DYNAMIC=1: always take the geometry-as-data kernels (tests)
int checkers_launch(int R, int C, int O, int N, int real, int tile, const CkParams &p, cudaStream_t stream);
int checkers_launch_f32_i8(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream);
int checkers_launch_f32_u2(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream);
int checkers_launch_f32(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream);
int checkers_launch_f64(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream);
int particle_launch(int N, int real, const PtParams &p, cudaStream_t stream);
int particle_pair_launch(int real, const PtParams &p, cudaStream_t stream);  // particle_pair.cu: N = 2, one lane per agent
constexpr int kDuoNotMine = 1;  // particle_duo_launch: "not my case", as opposed to a cm3_status
int particle_duo_launch(int N, int real, const PtParams &p, cudaStream_t stream);  // particle_duo.cu: N <= 2, two envs per thread
bool duo_enabled();             // CM3_PT_DUO=1 sends the common launch of one- and two-agent envs through the two-envs-per-thread kernel (experiment, off)
bool pair_enabled();            // CM3_PT_PAIR=1 sends two-agent envs through the one-lane-per-agent kernel (experiment, off)
int particle_tile_envs(int N);  // envs per warp tile (the granularity of cm3_particle_state.sync)
bool full_enabled();  // particle kernel: the specialised all-outputs / whole-tiles instantiation (CM3_PT_FULL=0 disables)
bool tma_enabled();  // swizzled tiles + tensor-map stores in the particle kernel (CM3_TMA=0 disables)
bool pdl_enabled();  // programmatic dependent launch between consecutive step launches (CM3_PDL=0 disables)
// Partial grids of a chained single-step launch (experiment, default 1 = off).  All blocks of one
// particle launch are resident at once and in the same phase (compute, then store), and blocks of
// the next launch only get the slots the previous ones free; the idea was that `parts` grids over
// disjoint tile ranges would let the stores of one part overlap the compute of the next.  Measured
// slower for every workload at 2, 3 and 4 parts (profiles/r02c_ab.txt: the extra launches cost more
// than the overlap returns); CM3_CHAIN_PARTS=<n> re-enables it.
int chain_parts(int ntiles);
// Where a chained launch writes its compact state back and publishes its tile (CkParams::early; -1 = decided
// at launch by chain_early_mode).  The next step's block of the same tile needs the state, not the outputs:
// storing the state and releasing the ticket BEFORE the observation tiles are assembled and stored takes the
// whole output phase - and the memory barrier of the release, which otherwise queues behind this block's own
// tile stores - off the tile-to-tile dependency chain.  It pays exactly when the launch is ONE resident wave:
// there the step time is the chain (every tile's next block is already resident, waiting for its ticket), e.g.
// at 65 536 envs PA3 0.687 -> 0.745, PM2 0.487 -> 0.54, PA4 0.755 -> 0.778, and at 32 768 envs CK2 0.734 -> 0.866,
// CK1 0.523 -> 0.691 (profiles/r02p_ab.txt, r02u_ab.txt).  With more blocks than resident slots the next block of
// a tile is mostly waiting for a SLOT, which frees when a block has stored its outputs, and the early barrier only
// delays those stores: CK2 0.896 -> 0.878 at 65 536 envs (1.7 waves), 0.781 -> 0.753 at 49 152, PA4 0.804 -> 0.785 at
// 131 072.  Rule: 2 (state and release before the outputs) when blocks <= resident slots of the kernel, else 0;
// CM3_CHAIN_EARLY=0|1|2 overrides (1 = state before, release after: within noise of 0).
int chain_early_mode(const void *kern, int threads, int smem, int nblocks);
// Tiles per block of a chained single-step launch of `nblocks` one-tile blocks: as many as it takes to make the launch
// ONE resident wave - ceil(nblocks / resident slots), at most 4 - so that the early release applies and fewer blocks
// queue for a slot; 1 when it fits anyway.  A block steps its tiles one after the other (a loop around the tile body).
// CM3_CHAIN_TPB=1..4 forces a value (4 where 2 would do is much slower: the serial chain inside a block).
int chain_tiles_per_block(const void *kern, int threads, int smem, int nblocks);
// The same for the particle kernels.  Their headline launch already is one wave (2048 blocks on 2368 slots), so the rule
// above gives 1 there; CM3_PT_TPB=<n> forces n tiles per block whenever the launch has at least 2 n blocks (experiment:
// half as many blocks would let two consecutive launches be resident together).
int particle_chain_tpb(const void *kern, int threads, int smem, int nblocks);

}  // namespace cm3
