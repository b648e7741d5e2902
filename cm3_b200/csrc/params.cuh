// params.cuh - kernel parameter blocks shared by the kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cm3env.h"

namespace cm3 {

constexpr int kCkMaxTR = 16;   // n_rows + 2*n_obs
constexpr int kCkMaxTC = 40;   // n_columns + 2*n_obs + 1
constexpr int kCkMaxCnt = 34;  // n_rows*n_columns/2 + 1

enum CkMode { kCkStep = 0, kCkReset = 1 };
enum PtMode { kPtStep = 0, kPtReset = 1 };

struct CkParams {
    uint64_t *remaining;
    uint32_t *agents;
    uint32_t *meta;
    const int8_t *actions;
    const uint8_t *goal_idx;
    const uint8_t *env_mask;
    int8_t *actions_out;
    char *grid, *vec, *obs_others, *obs_self_t, *obs_self_v, *reward, *local_rewards;
    uint8_t *done;
    int B, T, max_steps, mode, auto_reset;
    unsigned long long seed;
    long long t0, env_id_offset;
    int start_r[CM3_MAX_AGENTS], start_c[CM3_MAX_AGENTS];  // expanded coordinates
    // (r - total_rows/2.0)/total_rows etc., evaluated on the host in float64 exactly as
    // checkers.py:120-121,139 write them, so the device never divides
    double norm_row[kCkMaxTR], norm_col[kCkMaxTC], norm_cnt[kCkMaxCnt];
};

struct PtParams {
    char *sv, *landmarks;
    int32_t *steps, *collisions;
    uint8_t *reached;
    const int8_t *actions;
    int8_t *actions_out;
    const char *init_pos, *init_landmarks;
    const uint8_t *env_mask;
    char *global_state, *obs_others, *obs_self, *reward, *reward_n;
    uint8_t *done;
    int B, T, max_steps, mode, auto_reset;
    unsigned long long seed;
    long long t0, env_id_offset, reset_counter;
    double dt, damping, contact_force, contact_margin, dist_min, mass, sensitivity, reach_thresh;
    double agents_x[CM3_MAX_AGENTS], agents_y[CM3_MAX_AGENTS];
    double landmarks_x[CM3_MAX_AGENTS], landmarks_y[CM3_MAX_AGENTS];
    double initial_std, prob_random;
};

bool checkers_geometry_supported(int R, int C, int O, int N);
int checkers_launch(int R, int C, int O, int N, int real, const CkParams &p, cudaStream_t stream);
int particle_launch(int N, int real, const PtParams &p, cudaStream_t stream);

}  // namespace cm3
