// particle.cu - fused cooperative-navigation (multi-goal_spread) reset/step/rollout kernel, sm_100a.
//
// What it computes, for B env instances at once (reference file:line):
//   _set_action                       multiagent/environment.py:177-225
//   World.step / apply_action_force / apply_environment_force / get_collision_force /
//   integrate_state                   multiagent/core.py:117-196
//   Scenario.observation / reward / is_collision / done / reset_world
//                                     multiagent/scenarios/multi-goal_spread.py:65-154
//   MultiAgentEnv.step tail / reset   multiagent/environment.py:95-149
//
// Design:
//  * one THREAD per env, a warp per 32 consecutive envs, one warp per block.  All N <= 8 agents
//    of an env live in the registers of one thread, so the all-pairs contact force, the collision
//    penalty and the env-level reductions are plain register arithmetic: every pair is evaluated
//    once (in the reference's pair order, core.py:145-154), there are no shuffles, and the
//    65 536-env headline batch is 2048 warps = 14 per SM - one resident wave.  (Round-1's first
//    version used one lane per (env, agent); ncu showed it latency-bound: 55 warps per SM in 2.3
//    waves at 80 registers, every pair evaluated twice, 44 shuffles per step.)
//  * the observation rows are written into a per-warp shared-memory tile that holds the bytes of
//    the output arrays for those 32 envs and leaves the SM as one TMA store per field: full-line
//    HBM writes, like the Checkers kernel.  global_state and obs_self are the same bytes
//    (environment.py:113-116 vs multi-goal_spread.py:147), so one staged tile feeds both stores.
//    A thread writes its env's record, so consecutive lanes are a whole record (16 N .. 48 N bytes)
//    apart - a 4- to 8-way bank conflict per 16-byte store in a linear tile (ncu, round 1: 75 % of
//    the shared-store wavefronts were conflicts).  The tile is therefore laid out in the TMA
//    swizzle pattern (32/64/128-byte XOR swizzle chosen per record size, PtGeom::sw_bits) and
//    stored with cp.async.bulk.tensor through a tensor map (SASS: UTMASTG) that un-swizzles it on
//    the way out: conflict-free staging, linear global layout.  Ragged batches (B % 32 != 0) and
//    the multi-destination gather use the linear tile + plain bulk stores (UBLKCP).  The tensor-map
//    stores carry an L2 evict-first policy (write-once stream); for N <= 2 the staging set is
//    double buffered (PtGeom::kStages).
//    Rewards / done flags go straight from registers (consecutive threads -> consecutive words);
//  * arithmetic follows the reference operation by operation with round-to-nearest intrinsics
//    (no FMA contraction, IEEE sqrt/div, no fast-math), in float (throughput mode) or double
//    (free-running parity mode);
//  * T steps can be fused in one launch with state in registers, actions from Philox or streamed
//    one step ahead through shared memory (common.cuh: ActionStream) and in-kernel episode reset,
//    like the Checkers kernel;
//  * the common launch (all outputs, whole tiles, unit mass) has its own instantiation without the
//    run-time checks of the general one (template parameter FULL).
#include "particle_math.cuh"

namespace cm3 {

// Measured and dropped (round 2, profiles/r02d_ab.txt): float kernels with at most two agents storing
// their observation records straight from registers (16-byte streaming stores, no staging, no TMA, no
// drain wait - see emit).  The records of a warp are contiguous, but every store instruction fills
// only half of each 32-byte sector it touches: PM2 fused 0.63 -> 0.42 at 65 536 envs, 0.91 -> 0.60 at
// 262 144.  Full-line TMA stores it is.  CM3_PT_DIRECT=1 rebuilds the variant.
#ifndef CM3_PT_DIRECT
#define CM3_PT_DIRECT 0
#endif
template <int N> constexpr bool kDirectStores = (CM3_PT_DIRECT != 0) && N <= 2;

// GATHER = false: one destination set (reset / step / rollout); GATHER = true: rollout_gather,
// every element goes to n_dst destination sets (peer GPUs), linear tiles + plain bulk stores.
//
// FULL = true: the common case compiled without its run-time checks - every output field
// requested, every tile whole (B % 32 == 0), unit mass, tensor-map stores.  The kernel is bound by
// the latency of a warp's dependent instruction chain (3.5 warps per scheduler at the 65 536-env
// batch), so the ~25 uniform branches per step that the general kernel spends on optional outputs
// and ragged tiles are worth removing (ncu, round 1: 13 % of the stall samples were branch
// resolution and instruction fetch).
//
// Measured and dropped (profiles/r01m): a warp-specialised variant - a physics warp handing the
// (vel, pos) records of each step to an emitter warp through shared memory and named barriers - was
// 20-40 % SLOWER at 65 536 envs: two warps per tile cap the kernel at 72 registers for a single
// resident wave, and the physics warp lost more to spills and serialisation than the emitter took
// off its chain.
// Measured and dropped (round 2, profiles/r02b_ab.txt): asking ptxas for 20 / 24 resident blocks per
// SM (95 / 79 registers instead of 128 / 118, no spills) so that blocks of two chained single-step
// launches fit an SM together.  The shorter register budget lengthens the dependent chains:
// fused PA4 0.92 -> 0.77, PA3 0.87 -> 0.61, and even the chained per-step launches lost (0.78 -> 0.74).
// CM3_PT_MINB=<n> rebuilds that variant.
#ifndef CM3_PT_MINB
#define CM3_PT_MINB 1
#endif
__host__ __device__ constexpr int pt_min_blocks(int) { return CM3_PT_MINB; }

template <int N, typename Real, bool GATHER, bool FULL>
__global__ void __launch_bounds__(kWarp, pt_min_blocks(N)) particle_kernel(const __grid_constant__ PtParams p) {
    using Op = RealOps<Real>;
    using Gm = PtGeom<N, Real>;
    constexpr int NO = Gm::NO, LO = Gm::LO;
    constexpr uint32_t RS = (uint32_t)sizeof(Real);

    const int lane = threadIdx.x;
    const bool leader = elect_one();  // issues, commits and waits for this warp's TMA stores (common.cuh)
    // A block normally owns one tile.  A chained single-step launch may give every block p.tpb (2..4) tiles instead, one
    // grid stride apart, stepped one after the other (launch_pt; as in checkers.cu).
    const int tile_first = p.tile0 + (int)blockIdx.x;
    const int tile_stride = (int)gridDim.x;
    const int n_my_tiles = p.tpb > 1 ? p.tpb : 1;
    // launch chaining: the tickets of ALL my tiles are taken BEFORE the next grid may be scheduled (common.cuh)
    // (0xFFFFFFFF = not taken: neutral in the AND below, which must depend on every atomic that WAS issued)
    uint32_t mine0 = 0xFFFFFFFFu, mine1 = 0xFFFFFFFFu, mine2 = 0xFFFFFFFFu, mine3 = 0xFFFFFFFFu;
    if (p.sync != nullptr) {
        const int t1 = tile_first + tile_stride, t2 = t1 + tile_stride, t3 = t2 + tile_stride;
        const bool h0 = tile_first * kWarp < p.B, h1 = n_my_tiles > 1 && t1 * kWarp < p.B, h2 = n_my_tiles > 2 && t2 * kWarp < p.B,
                   h3 = n_my_tiles > 3 && t3 * kWarp < p.B;
        if (lane == 0) {  // all atomics in flight together, then one broadcast each
            if (h0) mine0 = atomicAdd(p.sync + 2 * (size_t)tile_first, 1u);
            if (h1) mine1 = atomicAdd(p.sync + 2 * (size_t)t1, 1u);
            if (h2) mine2 = atomicAdd(p.sync + 2 * (size_t)t2, 1u);
            if (h3) mine3 = atomicAdd(p.sync + 2 * (size_t)t3, 1u);
        }
        mine0 = __shfl_sync(0xFFFFFFFFu, mine0, 0);
        if (n_my_tiles > 1) {
            mine1 = __shfl_sync(0xFFFFFFFFu, mine1, 0);
            mine2 = __shfl_sync(0xFFFFFFFFu, mine2, 0);
            mine3 = __shfl_sync(0xFFFFFFFFu, mine3, 0);
        }
    } else {
        mine0 = 0;  // no chaining words: trigger at once
    }
    // the trigger is issued under a branch on the returned tickets: every atomic has been performed at L2 before any
    // block of the next grid can take its own
    if ((mine0 & mine1 & mine2 & mine3) != 0xFFFFFFFFu) pdl_launch_dependents();  // the next step's grid may become resident while this one drains

    extern __shared__ unsigned char smem_dyn[];
    unsigned char *stage_base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);

    const bool reset_mode = !FULL && p.mode == kPtReset;
#pragma unroll 1
    for (int my = 0; my < n_my_tiles; ++my) {
    const int tile = tile_first + my * tile_stride;
    const int env0 = tile * kWarp;
    if (env0 >= p.B) return;  // warp-uniform (one warp per block); later tiles of mine lie further out
    TileTicket ticket;
    ticket.on = p.sync != nullptr;
    ticket.w = p.sync + 2 * (size_t)tile;
    ticket.mine = my == 0 ? mine0 : my == 1 ? mine1 : my == 2 ? mine2 : mine3;
    const int env = env0 + lane;
    const int nenv = FULL ? kWarp : min(kWarp, p.B - env0);
    const bool valid = FULL || lane < nenv;
    const size_t B = (size_t)p.B;

    // world constants, rounded to Real once on the host (no per-thread F2F conversions)
    const PtConsts<Real> &K = pt_consts<Real>(p);
    const Real dt = K.dt, keep = K.keep, cf = K.contact_force, km = K.contact_margin, dist_min = K.dist_min;
    const Real mass = K.mass, sens = K.sensitivity, neg_reach = K.neg_reach;
    // squared distances beyond which a pair provably contributes exactly nothing (see below)
    const Real far2 = K.far2, near2 = K.near2;
    const bool unit_mass = FULL || (mass == (Real)1);  // x / 1 == x exactly: skip the IEEE division

    // ---- where the outputs go
    const PtOut &o0 = p.out[0];
    const bool has_gs = FULL || o0.global_state != nullptr, has_os = FULL || o0.obs_self != nullptr;
    const bool has_oo = FULL || o0.obs_others != nullptr;
    const bool has_rn = FULL || o0.reward_n != nullptr, has_rw = FULL || o0.reward != nullptr, has_dn = FULL || o0.done != nullptr;
    const bool tma = FULL || (!GATHER && p.tma != 0);           // swizzled tiles + tensor-map stores
    const uint32_t mask_row = tma ? Gm::kRowMask : 0u, mask_oo = tma ? Gm::kOthMask : 0u;
    if (tma && leader) {
        if (has_oo) tma_prefetch_map(&p.tm.oo);
        if (has_gs) tma_prefetch_map(&p.tm.gs);
        if (has_os) tma_prefetch_map(&p.tm.os);
    }
    const size_t OB = (size_t)p.out_B, oe0 = (size_t)p.out_env0;
    // per-thread output cursors of slot t = 0, advanced by one [out_B] slice per step
    Real *rn_ptr = reinterpret_cast<Real *>(o0.reward_n) + (oe0 + env) * N;
    Real *rw_ptr = reinterpret_cast<Real *>(o0.reward) + (oe0 + env);
    uint8_t *dn_ptr = o0.done + (oe0 + env);
    const bool has_cl = o0.collisions != nullptr;  // optional in every instantiation
    int32_t *cl_ptr = o0.collisions + (oe0 + env);
    const bool has_rc = o0.reached != nullptr;
    uint8_t *rc_ptr = o0.reached + (oe0 + env);
    int tile_idx = (int)((oe0 + env0) / kWarp);  // tensor-map tile coordinate (tma: out_B, out_env0 % 32 == 0)
    const int tiles_per_slot = (int)(OB / kWarp);

    // the state this tile's previous launch wrote is visible from here on
    if (p.chained) ticket.wait(lane); else pdl_wait();

    // ---- state of my env: all agents in registers
    Real vx[N], vy[N], px[N], py[N], lx[N], ly[N];
    int steps = 0, collisions = 0;
    uint32_t reached = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) { vx[i] = 0; vy[i] = 0; px[i] = (Real)(2 * i); py[i] = 0; lx[i] = 0; ly[i] = 0; }
    if (valid) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ld4<Real>(reinterpret_cast<const Real *>(p.sv) + ((size_t)env * N + i) * 4, vx[i], vy[i], px[i], py[i]);
            ld2<Real>(reinterpret_cast<const Real *>(p.landmarks) + ((size_t)env * N + i) * 2, lx[i], ly[i]);
        }
        steps = __ldcg(p.steps + env);
        collisions = __ldcg(p.collisions + env);
        reached = __ldcg(p.reached + env);
    }

    // multi-goal_spread.py:65-93 on Philox (or the injected host draws)
    auto reset_state = [&](unsigned long long counter, uint32_t tag) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (reset_mode && p.init_pos != nullptr) {
                ld2<Real>(reinterpret_cast<const Real *>(p.init_pos) + ((size_t)env * N + i) * 2, px[i], py[i]);
                ld2<Real>(reinterpret_cast<const Real *>(p.init_landmarks) + ((size_t)env * N + i) * 2, lx[i], ly[i]);
            } else {
                const ResetDraw<Real> d = draw_reset<Real>(p, (unsigned long long)(p.env_id_offset + env), counter, tag, i);
                px[i] = d.px; py[i] = d.py; lx[i] = d.lx; ly[i] = d.ly;
            }
            vx[i] = 0; vy[i] = 0;
        }
        steps = 0; collisions = 0; reached = 0;  // :84-86, :93; environment.py:148
    };

    bool pending = false;  // async stores whose shared-memory source may still be in flight

    // action rows: streamed through shared memory (multi-step launches on whole tiles, see
    // ActionStream), else loaded directly at the top of each step
    ActionStream<N> acts;
    acts.init(stage_base + Gm::kActOff, reset_mode ? nullptr : p.actions, p.B, env0, kWarp, nenv == kWarp, p.T, lane);
    uint32_t act_word = acts.on ? acts.begin(lane) : 0u;

    // observations of the current state -> outputs of slot t (multi-goal_spread.py:145-154,
    // environment.py:113-116)
    auto emit = [&](int t) {
        // next step's action word: in flight from here to the end of this phase (ActionStream)
        const uint32_t act_loaded = acts.on ? acts.load(t + 1) : 0u;
        if (acts.on) acts.prefetch(t + 3);
        if constexpr (FULL && kDirectStores<N> && sizeof(Real) == 4) {
            // N <= 2: an env's record is 16 N / 16 N max(N-1,1) bytes, consecutive lanes write
            // consecutive records, so the tile is a contiguous 512 N bytes per field either way.  Plain
            // 16-byte streaming stores from registers (two per field and lane for N = 2) reach the same
            // lines without the staging stores, the proxy fence, the three TMA issues and - the point -
            // without ever waiting for a staging tile to drain: with these small records a step is
            // short compared with the drain, and the wait was most of the step.
            const size_t rec = (size_t)t * OB + oe0 + env;
            float4 *gs = reinterpret_cast<float4 *>(reinterpret_cast<Real *>(o0.global_state) + rec * (N * 4));
            float4 *os = reinterpret_cast<float4 *>(reinterpret_cast<Real *>(o0.obs_self) + rec * (N * 4));
            float4 *oo = reinterpret_cast<float4 *>(reinterpret_cast<Real *>(o0.obs_others) + rec * (N * LO));
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float4 row = make_float4((float)vx[i], (float)vy[i], (float)px[i], (float)py[i]);
                __stcs(gs + i, row);
                __stcs(os + i, row);
                const int j = (N == 1) ? 0 : 1 - i;   // the one other agent (N = 1: the agent itself, :148-153)
                __stcs(oo + i, make_float4((float)Op::sub(vx[j], vx[i]), (float)Op::sub(vy[j], vy[i]),
                                           (float)Op::sub(px[j], px[i]), (float)Op::sub(py[j], py[i])));
            }
            if (acts.on) act_word = acts.hand_over(t, act_loaded, lane);
            return;
        }
        // this step's staging set; with two sets only the stores of step t - 2 must have left it
        unsigned char *stage_row = stage_base + (Gm::kStages == 2 ? (t & 1) * Gm::kSetBytes : 0);
        unsigned char *stage_oo = stage_row + Gm::kOthOff;
        if (pending) {
            if (leader) {
                if (Gm::kStages == 2) bulk_wait_read_but_one(); else bulk_wait_read();
            }
        }
        __syncwarp();
        if (valid) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                // [vel, pos] of agent i: global_state row and obs_self alike
                if (has_gs || has_os) stage4<Real>(stage_row, (uint32_t)(lane * N + i) * 4u * RS, mask_row, vx[i], vy[i], px[i], py[i]);
                if (has_oo) {
                    const uint32_t oo = (uint32_t)(lane * N + i) * (uint32_t)LO * RS;
                    if (N == 1) {  // "others" is the agent itself, :148-153
                        stage4<Real>(stage_oo, oo, mask_oo, Op::sub(vx[0], vx[0]), Op::sub(vy[0], vy[0]),
                                     Op::sub(px[0], px[0]), Op::sub(py[0], py[0]));
                    } else {
#pragma unroll
                        for (int k = 0; k < NO; ++k) {
                            const int j = k + (k >= i ? 1 : 0);  // compile-time after unrolling
                            stage4<Real>(stage_oo, oo + 4u * k * RS, mask_oo, Op::sub(vx[j], vx[i]), Op::sub(vy[j], vy[i]),
                                         Op::sub(px[j], px[i]), Op::sub(py[j], py[i]));
                        }
                    }
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        pending = false;
        if (tma) {
            // whole tiles only (B % 32 == 0): one tensor-map store per field, un-swizzling on the way out
            if (leader) {
                const int tile = tile_idx + t * tiles_per_slot;
#ifndef CM3_NO_L2_HINT_TMA  // write-once stream: evict-first (see common.cuh, L2 cache policies)
                const uint64_t pol = l2_policy_evict_first();
                if (has_oo) tma_store_2d_hint(&p.tm.oo, stage_oo, 0, tile * Gm::kOthRows, pol);
                if (has_gs) tma_store_2d_hint(&p.tm.gs, stage_row, 0, tile * Gm::kRowRows, pol);
                if (has_os) tma_store_2d_hint(&p.tm.os, stage_row, 0, tile * Gm::kRowRows, pol);
#else
                if (has_oo) tma_store_2d(&p.tm.oo, stage_oo, 0, tile * Gm::kOthRows);
                if (has_gs) tma_store_2d(&p.tm.gs, stage_row, 0, tile * Gm::kRowRows);
                if (has_os) tma_store_2d(&p.tm.os, stage_row, 0, tile * Gm::kRowRows);
#endif
                bulk_commit();
            }
            pending = true;
            if (acts.on) act_word = acts.hand_over(t, act_loaded, lane);
            return;
        }
        const size_t row0 = (size_t)t * OB + oe0 + env0;
        // linear tiles, n_dst bulk stores each: with rollout_gather the destinations are the
        // rollout buffers of every GPU of the node (peer memory over NVLink)
        auto put = [&](char *PtOut::*field, const unsigned char *stage_b, int per_env) {
            if (o0.*field == nullptr) return;
            const Real *stage = reinterpret_cast<const Real *>(stage_b);
            const uint32_t bytes = (uint32_t)(nenv * per_env * sizeof(Real));
            const int nd = GATHER ? p.n_dst : 1;
            for (int d = 0; d < nd; ++d) {
                Real *g = reinterpret_cast<Real *>(p.out[d].*field) + row0 * (size_t)per_env;
                if (nenv == kWarp && (reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
                    if (leader) bulk_store(g, stage, bytes);
                    pending = true;
                } else {
                    for (int idx = lane; idx < nenv * per_env; idx += kWarp) g[idx] = stage[idx];
                }
            }
        };
        put(&PtOut::obs_others, stage_oo, N * LO);
        put(&PtOut::global_state, stage_row, N * 4);
        put(&PtOut::obs_self, stage_row, N * 4);
        if (leader) bulk_commit();
        if (acts.on) act_word = acts.hand_over(t, act_loaded, lane);
    };

    // compact state back to HBM
    auto store_state = [&]() {
        if (valid) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                st4<Real>(reinterpret_cast<Real *>(p.sv) + ((size_t)env * N + i) * 4, vx[i], vy[i], px[i], py[i]);
                if (reset_mode || p.auto_reset) {  // landmarks only change on a reset
                    Real *lm = reinterpret_cast<Real *>(p.landmarks) + ((size_t)env * N + i) * 2;
                    lm[0] = lx[i]; lm[1] = ly[i];
                }
            }
            p.steps[env] = steps;
            p.collisions[env] = collisions;
            p.reached[env] = (uint8_t)reached;
        }
    };

    const int T_eff = reset_mode ? 1 : p.T;
    for (int t = 0; t < T_eff; ++t) {
        bool sel = false;
        if (reset_mode) {
            sel = valid && (p.env_mask == nullptr || p.env_mask[env] != 0);
            if (sel) reset_state((unsigned long long)p.reset_counter, kTagReset);
        } else {
            // ---- actions -> control forces (environment.py:194-214, core.py:134-140)
            int act[N];
            if (p.actions != nullptr) {
                if constexpr (N <= 4) {
                    uint32_t w = act_word;  // picked up from the stream during the previous emit
                    if (!acts.on && valid) w = load_actions_packed<N>(p.actions + ((size_t)t * B + env) * N);
#pragma unroll
                    for (int i = 0; i < N; ++i) act[i] = unpack_action(w, i);
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) act[i] = valid ? (int)p.actions[((size_t)t * B + env) * N + i] : 0;
                }
            } else {
                // one Philox block of 4 words per 4 agents
#pragma unroll
                for (int i0 = 0; i0 < N; i0 += 4) {
                    const Philox4 w = philox_action_words(p.seed, (uint64_t)(p.env_id_offset + env), (uint64_t)(p.t0 + t), i0 / 4);
#pragma unroll
                    for (int i = i0; i < N && i < i0 + 4; ++i) act[i] = action_from_word(philox_word(w, i - i0), 5);
                }
            }
            if (p.actions_out != nullptr && valid) {
#pragma unroll
                for (int i = 0; i < N; ++i) p.actions_out[((size_t)t * B + env) * N + i] = (int8_t)act[i];
            }
            Real fx[N], fy[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const Real ux = (act[i] == 1) ? (Real)-1 : (act[i] == 2) ? (Real)1 : (Real)0;
                const Real uy = (act[i] == 3) ? (Real)-1 : (act[i] == 4) ? (Real)1 : (Real)0;
                fx[i] = Op::mul(ux, sens); fy[i] = Op::mul(uy, sens);
            }

            // ---- contact forces over the agent pairs a < b in the reference's order
            // (core.py:143-155, 180-196; landmarks do not collide, multi-goal_spread.py:55).
            // The reference evaluates the softplus penetration for EVERY pair at every distance;
            // beyond dist_min + ~110 k (float) / ~760 k (double) exp() underflows to exactly 0, so
            // pen == 0, the force is +-0 and "F + p_force" returns p_force bit for bit (p_force is
            // never -0: it starts as +0 or +-sensitivity).  Those pairs are skipped: same bits,
            // none of the double-precision sqrt / div / exp / log1p instructions.
            // All squared distances first (independent arithmetic), then ONE branch for the warp's
            // common case "no pair anywhere near contact".
            constexpr int NPAIR = N * (N - 1) / 2;
            if (NPAIR > 0) {
                Real d2c[NPAIR > 0 ? NPAIR : 1];
                bool any_near = false;
                int q = 0;
#pragma unroll
                for (int a = 0; a < N; ++a) {
#pragma unroll
                    for (int b = a + 1; b < N; ++b) {
                        const Real ex = Op::sub(px[a], px[b]), ey = Op::sub(py[a], py[b]);
                        d2c[q] = Op::add(Op::mul(ex, ex), Op::mul(ey, ey));
                        any_near = any_near || !(d2c[q] > far2);  // near, coincident or NaN
                        ++q;
                    }
                }
                if (any_near) {
                    q = 0;
#pragma unroll
                    for (int a = 0; a < N; ++a) {
#pragma unroll
                        for (int b = a + 1; b < N; ++b) {
                            if (!(d2c[q] > far2)) {  // the literal evaluation
                                Force2<Real> F;
                                if constexpr (N <= 2) F = contact_force_inl<Real>(px[a], py[a], px[b], py[b], p);
                                else F = contact_force_ool<Real>(px[a], py[a], px[b], py[b], p);
                                fx[a] = Op::add(F.x, fx[a]); fy[a] = Op::add(F.y, fy[a]);    // f_a + p_force[a]
                                fx[b] = Op::add(-F.x, fx[b]); fy[b] = Op::add(-F.y, fy[b]);  // f_b = -force
                            }
                            ++q;
                        }
                    }
                }
            }
            // ---- integrate (core.py:158-169)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (!unit_mass) { fx[i] = Op::div(fx[i], mass); fy[i] = Op::div(fy[i], mass); }
                vx[i] = Op::mul(vx[i], keep); vy[i] = Op::mul(vy[i], keep);
                vx[i] = Op::add(vx[i], Op::mul(fx[i], dt)); vy[i] = Op::add(vy[i], Op::mul(fy[i], dt));
                px[i] = Op::add(px[i], Op::mul(vx[i], dt)); py[i] = Op::add(py[i], Op::mul(vy[i], dt));
            }
            steps += 1;  // environment.py:93

            // ---- rewards (multi-goal_spread.py:121-138)
            Real rew[N];
            uint32_t reach_bits = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const Real tx = Op::sub(px[i], lx[i]), ty = Op::sub(py[i], ly[i]);
                rew[i] = Op::sub((Real)0, Op::sqrt(Op::add(Op::mul(tx, tx), Op::mul(ty, ty))));
                reach_bits |= (rew[i] >= neg_reach ? 1u : 0u) << i;  // :126-129
            }
            int hits = 0;
            if (NPAIR > 0) {
                Real d2r[NPAIR > 0 ? NPAIR : 1];
                bool any_close = false;
                int q = 0;
#pragma unroll
                for (int a = 0; a < N; ++a) {
#pragma unroll
                    for (int b = a + 1; b < N; ++b) {
                        const Real dx = Op::sub(px[a], px[b]), dy = Op::sub(py[a], py[b]);
                        d2r[q] = Op::add(Op::mul(dx, dx), Op::mul(dy, dy));
                        // sqrt is monotonic: d2 > near2 = (1.01 dist_min)^2 cannot round below dist_min
                        any_close = any_close || !(d2r[q] > near2);
                        ++q;
                    }
                }
                if (any_close) {
                    q = 0;
#pragma unroll
                    for (int a = 0; a < N; ++a) {
#pragma unroll
                        for (int b = a + 1; b < N; ++b) {
                            // is_collision, :114-118 - the penalty and the counter hit both agents
                            if (!(d2r[q] > near2) && Op::sqrt(d2r[q]) < dist_min) {
                                rew[a] = Op::sub(rew[a], (Real)1); rew[b] = Op::sub(rew[b], (Real)1);
                                hits += 2;
                            }
                            ++q;
                        }
                    }
                }
            }
            Real total = rew[0];
#pragma unroll
            for (int i = 1; i < N; ++i) total = Op::add(total, rew[i]);  // np.sum, environment.py:107
            collisions += hits;
            reached = reach_bits;
            const bool done = (steps == p.max_steps) || (reach_bits == (1u << N) - 1u);  // environment.py:118
            if (valid) {
                const int nd = GATHER ? p.n_dst : 1;
                for (int d = 0; d < nd; ++d) {
                    // destination d: same cursors, rebased from set 0 to set d (GATHER only)
                    Real *rn = rn_ptr, *rw = rw_ptr;
                    uint8_t *dn = dn_ptr;
                    int32_t *cl = cl_ptr;
                    uint8_t *rc = rc_ptr;
                    if (GATHER && d > 0) {
                        rc = p.out[d].reached + (rc_ptr - o0.reached);
                        rn = reinterpret_cast<Real *>(p.out[d].reward_n + (reinterpret_cast<char *>(rn_ptr) - o0.reward_n));
                        rw = reinterpret_cast<Real *>(p.out[d].reward + (reinterpret_cast<char *>(rw_ptr) - o0.reward));
                        dn = p.out[d].done + (dn_ptr - o0.done);
                        cl = p.out[d].collisions + (cl_ptr - o0.collisions);
                    }
                    if (has_rn) {
                        if (N % 4 == 0) {
#pragma unroll
                            for (int i = 0; i + 3 < N; i += 4) st4<Real>(rn + i, rew[i], rew[i + 1 < N ? i + 1 : 0], rew[i + 2 < N ? i + 2 : 0], rew[i + 3 < N ? i + 3 : 0]);
                        } else {
#pragma unroll
                            for (int i = 0; i < N; ++i) rn[i] = rew[i];
                        }
                    }
                    if (has_rw) *rw = total;
                    if (has_dn) *dn = done ? 1 : 0;
                    if (has_cl) *cl = collisions;  // the episode's count so far, before a reset zeroes it
                    if (has_rc) *rc = (uint8_t)reach_bits;
                }
            }
            rn_ptr += OB * N; rw_ptr += OB; dn_ptr += OB; cl_ptr += OB; rc_ptr += OB;
            if (p.auto_reset && done) reset_state((unsigned long long)(p.t0 + t + 1), kTagAutoReset);
        }
        // The state is final here: the next launch of this tile needs IT, not the observations.  p.early (chained
        // launches, params.cuh: chain_early_mode) writes it back - and, at 2, releases the tile - before the
        // observation tiles of the last step are assembled and stored.
        if (p.early != 0 && t == T_eff - 1) {
            // (no warp barrier needed here, unlike checkers.cu: a lane loads and stores the state of ITS env only)
            store_state();
            if (p.early == 2) ticket.publish(lane);
        }
        emit(t);
        if (sel && o0.done != nullptr) o0.done[oe0 + env] = 0;  // np.any(done_n), environment.py:149
    }

    if (p.early == 0 || T_eff < 1) store_state();
    if (p.early != 2 || T_eff < 1) ticket.publish(lane);  // this tile's next launch may go ahead
    // shared memory must outlive the async reads (and be free for my next tile); the global writes themselves complete with the grid
    if (pending && leader) bulk_wait_read();
    __syncwarp();
    }  // my tiles
}

// ------------------------------------------------------------------------ host side

// A [T][out_B] output field seen as a 2-D byte tensor of `width`-byte rows; one box = the tile of
// 32 consecutive envs.  The swizzle mode is the one the kernel staged the tile in.
bool encode_tile_map(CUtensorMap *tm, void *base, size_t total_bytes, int sw_bits, int width, int box_rows) {
    static const PFN_cuTensorMapEncodeTiled_v12000 encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            fn = nullptr;
        }
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }();
    if (!encode || (reinterpret_cast<uintptr_t>(base) & 15u) != 0 || total_bytes % (size_t)width != 0) return false;
    const cuuint64_t rows = total_bytes / (size_t)width;
    if (rows == 0 || rows > 0x7FFFFFFFull) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)width, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)width};
    const cuuint32_t box[2] = {(cuuint32_t)width, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1, 1};
    const CUtensorMapSwizzle sw = sw_bits == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : sw_bits == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : sw_bits == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int N, typename Real, bool GATHER, bool FULL>
static int launch_pt(const PtParams &p, cudaStream_t stream) {
    using Gm = PtGeom<N, Real>;
    auto kern = particle_kernel<N, Real, GATHER, FULL>;
    constexpr int kSmem = Gm::kSmemBytes;
    static std::atomic<uint64_t> attr_done{0};
    CM3_CUDA(ensure_smem_attr(kern, attr_done));
    const int nblocks = (p.B + kWarp - 1) / kWarp;
    // multi-step launches: equal waves (common.cuh: balance_waves)
    const int smem_launch = (p.mode == kPtStep && p.T > 1) ? balance_waves((const void *)kern, kWarp, kSmem, nblocks) : kSmem;
    const int parts = p.chained ? chain_parts(nblocks) : 1;
    // chained single-step launches may step several tiles per block (params.cuh: chain_tiles_per_block; CM3_PT_TPB)
    const int tpb = (p.chained && p.T == 1 && parts == 1) ? particle_chain_tpb((const void *)kern, kWarp, smem_launch, nblocks) : 1;
    const int grid_blocks = (nblocks + tpb - 1) / tpb;
    const int early = (p.chained && p.early < 0) ? chain_early_mode((const void *)kern, kWarp, smem_launch, grid_blocks) : (p.early < 0 ? 0 : p.early);
    for (int i = 0; i < parts; ++i) {  // disjoint tile ranges; one grid unless chained (params.cuh: chain_parts)
        PtParams q = p;
        q.early = early;
        q.tpb = tpb;
        q.tile0 = (int)((long long)grid_blocks * i / parts);
        const int n = (int)((long long)grid_blocks * (i + 1) / parts) - q.tile0;
        if (n > 0) CM3_CUDA(launch_kernel(kern, n, kWarp, smem_launch, stream, pdl_enabled(), q));
    }
    return CM3_OK;
}

// tensor maps for a single-destination launch over whole tiles; p.tma says whether they were made
template <int N, typename Real>
static void make_tensor_maps(PtParams &p) {
    using Gm = PtGeom<N, Real>;
    p.tma = 0;
    if (Gm::kTmaOk && p.n_dst == 1 && tma_enabled() && p.B % kWarp == 0 && p.out_B % kWarp == 0 && p.out_env0 % kWarp == 0) {
        const int T = (p.mode == kPtReset) ? 1 : p.T;
        const size_t envs = (size_t)T * (size_t)p.out_B;
        const PtOut &o = p.out[0];
        bool ok = true;
        if (o.global_state) ok = ok && encode_tile_map(&p.tm.gs, o.global_state, envs * N * 4 * sizeof(Real), Gm::kRowSw, Gm::kRowW, Gm::kRowRows);
        if (o.obs_self) ok = ok && encode_tile_map(&p.tm.os, o.obs_self, envs * N * 4 * sizeof(Real), Gm::kRowSw, Gm::kRowW, Gm::kRowRows);
        if (o.obs_others) ok = ok && encode_tile_map(&p.tm.oo, o.obs_others, envs * N * Gm::LO * sizeof(Real), Gm::kOthSw, Gm::kOthW, Gm::kOthRows);
        p.tma = ok ? 1 : 0;
    }
}

template <int N, typename Real>
static int dispatch_pt(const PtParams &p0, cudaStream_t stream) {
    if (p0.n_dst > 1) return launch_pt<N, Real, true, false>(p0, stream);
    PtParams p = p0;
    make_tensor_maps<N, Real>(p);
    const PtOut &o = p.out[0];
    // FULL: the step / rollout launch with nothing optional left to test at run time
    const bool full = p.tma && p.mode == kPtStep && p.mass == 1.0 && o.global_state && o.obs_self && o.obs_others &&
                      o.reward && o.reward_n && o.done && full_enabled();
    if constexpr (PtGeom<N, Real>::kTmaOk) {
        if (full) return launch_pt<N, Real, false, true>(p, stream);
    }
    return launch_pt<N, Real, false, false>(p, stream);
}

int particle_launch(int N, int real, const PtParams &p, cudaStream_t stream) {
    if (N == 2 && pair_enabled()) return particle_pair_launch(real, p, stream);
    if (N <= 2 && duo_enabled()) {  // the common launch of small envs: two envs per thread (particle_duo.cu)
        const int rc = particle_duo_launch(N, real, p, stream);
        if (rc != kDuoNotMine) return rc;
    }
#define CASE(n) \
    case n: return real == CM3_REAL_F64 ? dispatch_pt<n, double>(p, stream) : dispatch_pt<n, float>(p, stream);
    switch (N) {
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        default: break;
    }
#undef CASE
    set_error("no compiled particle kernel for n_agents=%d (1..%d supported)", N, CM3_MAX_AGENTS);
    return CM3_ERR_UNSUPPORTED;
}

}  // namespace cm3
