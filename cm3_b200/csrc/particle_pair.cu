// particle_pair.cu - EXPERIMENT, off by default (CM3_PT_PAIR=1 enables it): the two-agent
// cooperative-navigation env (the reference's "merge" scenario, alg/config_particle_stage2_merge.json)
// with ONE LANE PER AGENT, a warp owning 16 envs.
//
// The idea.  particle.cu keeps all agents of an env in one thread, which is the right trade for
// N >= 3 (every pair once, no shuffles) - but the batch then offers only B threads: at 65 536 envs
// that is 2048 warps, 14 per SM, 3.5 per scheduler, and a two-agent step is a single dependent chain.
// ncu on that kernel at this batch: 58 % of the issue slots used, the top stall is the fixed-latency
// dependency wait (0.63 of the HBM roofline at 65 536 envs, 0.91 at 262 144).  Splitting an env over
// two lanes doubles the independent instruction streams: each lane integrates ITS agent, evaluates
// the pair force redundantly from both positions (in the reference's operand order, agent 0 minus
// agent 1, core.py:186, so both lanes get the same bits and agent 1 applies -F, core.py:153-154) and
// trades the four state words with its partner through one round of shuffles per step.  Rows of
// consecutive lanes are consecutive records of the output arrays, so the staging tile is linear and
// conflict-free by construction (no swizzle, plain bulk stores).
//
// The measurement (profiles/r02h_ab.txt, r02h_pm2_pair_ncu.txt).  Bit-identical to the one-thread-per-
// env kernel (tests/test_gpu_round2.py::test_pair_kernel_is_bit_identical...), and SLOWER on the fused
// rollout: 0.535 vs 0.635 of the roofline at 65 536 envs, 0.62 vs 0.90 at 262 144.  A warp-step costs
// 407 instructions here against 423 there - what a step spends per WARP (action stream, staging fence,
// three TMA issues, cursors, the contact evaluation, which a warp takes whenever any lane needs it) is
// untouched by giving a lane half the agents, so twice the warps are twice the instructions and the
// kernel turns issue-bound (74 % issue-active, "not selected" the second stall).  Chained single-step
// launches do gain (0.44 vs 0.38).  Kept as a validated variant; the default path for N = 2 stays
// particle.cu.
//
// What it computes is particle.cu's contract for N = 2, operation by operation (same helpers:
// particle_math.cuh): _set_action (environment.py:177-225), World.step (core.py:117-196),
// Scenario.reward / observation / done / reset_world (multi-goal_spread.py:65-154), the
// MultiAgentEnv.step tail (environment.py:95-123), T-step rollouts with the action stream or Philox,
// in-kernel episode reset, chained launches, the multi-destination gather.
#include "particle_math.cuh"

namespace cm3 {

constexpr int kPairEnvs = kWarp / 2;  // envs per warp
// The point of the mapping is residency: 65 536 envs are 4096 one-warp blocks, 27.7 per SM, so the
// kernel must fit 28 blocks per SM - at most 73 registers per thread (one agent per lane needs fewer).
#ifndef CM3_PAIR_MINB
#define CM3_PAIR_MINB 28
#endif
constexpr int kPairMinBlocks = CM3_PAIR_MINB;

template <typename Real>
struct PairGeom {
    static constexpr int kRowBytes = kWarp * 4 * (int)sizeof(Real);  // one [vel, pos] row (or one "other" row) per lane
    static constexpr int kSetBytes = 2 * kRowBytes;                   // row tile + others tile
    static constexpr int kStages = 2;                                  // double buffered: a step is short compared with the drain
    static constexpr int kActOff = kStages * kSetBytes;
    static constexpr int kSmemBytes = kActOff + ActionStream<2>::kSmemBytes;
};

// FULL: every reference output requested, whole tiles, unit mass, step / rollout mode, one destination.
template <typename Real, bool GATHER, bool FULL>
__global__ void __launch_bounds__(kWarp, kPairMinBlocks) particle_pair_kernel(const __grid_constant__ PtParams p) {
    using Op = RealOps<Real>;
    using Gm = PairGeom<Real>;
    constexpr int N = 2;
    constexpr uint32_t RS = (uint32_t)sizeof(Real);

    const int lane = threadIdx.x;
    const bool leader = elect_one();
    const int tile = p.tile0 + (int)blockIdx.x;
    TileTicket ticket;
    ticket.take(p.sync, tile, lane);
    if (ticket.mine != 0xFFFFFFFFu) pdl_launch_dependents();

    extern __shared__ __align__(128) unsigned char smem_pair[];

    const int sub = lane & 1, e = lane >> 1;  // my agent, my env inside the tile
    const bool reset_mode = !FULL && p.mode == kPtReset;
    const int env0 = tile * kPairEnvs;
    const int env = env0 + e;
    const int nenv = FULL ? kPairEnvs : min(kPairEnvs, p.B - env0);
    const bool valid = FULL || e < nenv;
    const size_t B = (size_t)p.B;

    const PtConsts<Real> &K = pt_consts<Real>(p);
    const Real dt = K.dt, keep = K.keep, dist_min = K.dist_min, mass = K.mass, sens = K.sensitivity, neg_reach = K.neg_reach;
    const Real far2 = K.far2, near2 = K.near2;
    const bool unit_mass = FULL || (mass == (Real)1);

    const PtOut &o0 = p.out[0];
    const bool has_gs = FULL || o0.global_state != nullptr, has_os = FULL || o0.obs_self != nullptr;
    const bool has_oo = FULL || o0.obs_others != nullptr;
    const bool has_rn = FULL || o0.reward_n != nullptr, has_rw = FULL || o0.reward != nullptr, has_dn = FULL || o0.done != nullptr;
    const bool has_cl = o0.collisions != nullptr, has_rc = o0.reached != nullptr;
    const size_t OB = (size_t)p.out_B, oe0 = (size_t)p.out_env0;
    Real *rn_ptr = reinterpret_cast<Real *>(o0.reward_n) + (oe0 + env) * N + sub;
    Real *rw_ptr = reinterpret_cast<Real *>(o0.reward) + (oe0 + env);
    uint8_t *dn_ptr = o0.done + (oe0 + env);
    int32_t *cl_ptr = o0.collisions + (oe0 + env);
    uint8_t *rc_ptr = o0.reached + (oe0 + env);

    if (p.chained) ticket.wait(lane); else pdl_wait();

    // ---- my agent's state, my partner's state, the env's counters
    Real vx = 0, vy = 0, px = (Real)(2 * sub), py = 0, lx = 0, ly = 0;
    Real ovx = 0, ovy = 0, opx = (Real)(2 - 2 * sub), opy = 0;
    int steps = 0, collisions = 0;
    uint32_t reached = 0;
    if (valid) {
        const Real *sv = reinterpret_cast<const Real *>(p.sv) + (size_t)env * N * 4;
        ld4<Real>(sv + sub * 4, vx, vy, px, py);
        ld4<Real>(sv + (1 - sub) * 4, ovx, ovy, opx, opy);
        ld2<Real>(reinterpret_cast<const Real *>(p.landmarks) + ((size_t)env * N + sub) * 2, lx, ly);
        steps = __ldcg(p.steps + env);
        collisions = __ldcg(p.collisions + env);
        reached = __ldcg(p.reached + env);
    }

    // multi-goal_spread.py:65-93.  Both lanes of an env take the same decision (it is keyed by the env)
    // and each draws BOTH agents' positions, so a reset needs no exchange between the lanes.
    auto reset_state = [&](unsigned long long counter, uint32_t tag) {
        if (reset_mode && p.init_pos != nullptr) {
            const Real *ip = reinterpret_cast<const Real *>(p.init_pos) + (size_t)env * N * 2;
            ld2<Real>(ip + sub * 2, px, py);
            ld2<Real>(ip + (1 - sub) * 2, opx, opy);
            ld2<Real>(reinterpret_cast<const Real *>(p.init_landmarks) + ((size_t)env * N + sub) * 2, lx, ly);
        } else {
            const unsigned long long genv = (unsigned long long)(p.env_id_offset + env);
            const ResetDraw<Real> mine = draw_reset<Real>(p, genv, counter, tag, sub);
            const ResetDraw<Real> other = draw_reset<Real>(p, genv, counter, tag, 1 - sub);
            px = mine.px; py = mine.py; lx = mine.lx; ly = mine.ly;
            opx = other.px; opy = other.py;
        }
        vx = 0; vy = 0; ovx = 0; ovy = 0;
        steps = 0; collisions = 0; reached = 0;
    };

    bool pending = false;
    ActionStream<N> acts;
    acts.init(smem_pair + Gm::kActOff, reset_mode ? nullptr : p.actions, p.B, env0, kPairEnvs, nenv == kPairEnvs, p.T, lane);
    uint32_t act_word = acts.on ? acts.begin(e) : 0u;

    // observations of the current state -> outputs of slot t: lane L writes record L of the tile
    auto emit = [&](int t) {
        const uint32_t act_loaded = acts.on ? acts.load(t + 1) : 0u;
        if (acts.on) acts.prefetch(t + 3);
        unsigned char *stage_row = smem_pair + (t & 1) * Gm::kSetBytes;
        unsigned char *stage_oo = stage_row + Gm::kRowBytes;
        if (pending && leader) bulk_wait_read_but_one();   // the set staged two steps ago has left
        __syncwarp();
        if (valid) {
            if (has_gs || has_os) stage4<Real>(stage_row, (uint32_t)lane * 4u * RS, 0u, vx, vy, px, py);
            if (has_oo) stage4<Real>(stage_oo, (uint32_t)lane * 4u * RS, 0u, Op::sub(ovx, vx), Op::sub(ovy, vy),
                                     Op::sub(opx, px), Op::sub(opy, py));   // multi-goal_spread.py:148-153
        }
        fence_proxy_async();
        __syncwarp();
        pending = false;
        const size_t row0 = ((size_t)t * OB + oe0 + env0) * (size_t)(N * 4);   // in Reals, same for all three fields
        const uint32_t bytes = (uint32_t)(nenv * N * 4 * sizeof(Real));
        const int nd = GATHER ? p.n_dst : 1;
        auto put = [&](char *PtOut::*field, const unsigned char *stage_b, bool wanted) {
            if (!wanted) return;
            const Real *stage = reinterpret_cast<const Real *>(stage_b);
            for (int d = 0; d < nd; ++d) {
                Real *g = reinterpret_cast<Real *>(p.out[d].*field) + row0;
                if (FULL || (nenv == kPairEnvs && (reinterpret_cast<uintptr_t>(g) & 15u) == 0)) {
                    if (leader) bulk_store(g, stage, bytes);
                    pending = true;
                } else {
                    for (int idx = lane; idx < nenv * N * 4; idx += kWarp) g[idx] = stage[idx];
                }
            }
        };
        put(&PtOut::obs_others, stage_oo, has_oo);
        put(&PtOut::global_state, stage_row, has_gs);
        put(&PtOut::obs_self, stage_row, has_os);
        if (leader) bulk_commit();
        if (acts.on) act_word = acts.hand_over(t, act_loaded, e);
    };

    const int T_eff = reset_mode ? 1 : p.T;
    for (int t = 0; t < T_eff; ++t) {
        bool sel = false;
        if (reset_mode) {
            sel = valid && (p.env_mask == nullptr || p.env_mask[env] != 0);
            if (sel) reset_state((unsigned long long)p.reset_counter, kTagReset);
        } else {
            // ---- my action -> my control force (environment.py:194-214, core.py:134-140)
            int act;
            if (p.actions != nullptr) {
                uint32_t w = act_word;
                if (!acts.on && valid) w = load_actions_packed<N>(p.actions + ((size_t)t * B + env) * N);
                act = unpack_action(w, sub);
            } else {
                const Philox4 w = philox_action_words(p.seed, (uint64_t)(p.env_id_offset + env), (uint64_t)(p.t0 + t));
                act = action_from_word(sub ? w.y : w.x, 5);
            }
            if (p.actions_out != nullptr && valid) p.actions_out[((size_t)t * B + env) * N + sub] = (int8_t)act;
            const Real ux = (act == 1) ? (Real)-1 : (act == 2) ? (Real)1 : (Real)0;
            const Real uy = (act == 3) ? (Real)-1 : (act == 4) ? (Real)1 : (Real)0;
            Real fx = Op::mul(ux, sens), fy = Op::mul(uy, sens);

            // ---- the one agent pair (core.py:143-155, 180-196): a = agent 0, b = agent 1 on BOTH lanes
            const Real ax = sub ? opx : px, ay = sub ? opy : py, bx = sub ? px : opx, by = sub ? py : opy;
            {
                const Real ex = Op::sub(ax, bx), ey = Op::sub(ay, by);
                const Real d2 = Op::add(Op::mul(ex, ex), Op::mul(ey, ey));
                if (!(d2 > far2)) {  // near, coincident or NaN: the literal evaluation (see particle.cu for the exact-zero skip)
                    const Force2<Real> F = contact_force_inl<Real>(ax, ay, bx, by, p);
                    fx = Op::add(sub ? -F.x : F.x, fx);   // f_a + p_force[a];  f_b = -force
                    fy = Op::add(sub ? -F.y : F.y, fy);
                }
            }
            // ---- integrate my agent (core.py:158-169), then trade the new state with my partner
            if (!unit_mass) { fx = Op::div(fx, mass); fy = Op::div(fy, mass); }
            vx = Op::mul(vx, keep); vy = Op::mul(vy, keep);
            vx = Op::add(vx, Op::mul(fx, dt)); vy = Op::add(vy, Op::mul(fy, dt));
            px = Op::add(px, Op::mul(vx, dt)); py = Op::add(py, Op::mul(vy, dt));
            ovx = __shfl_xor_sync(0xFFFFFFFFu, vx, 1); ovy = __shfl_xor_sync(0xFFFFFFFFu, vy, 1);
            opx = __shfl_xor_sync(0xFFFFFFFFu, px, 1); opy = __shfl_xor_sync(0xFFFFFFFFu, py, 1);
            steps += 1;  // environment.py:93

            // ---- my reward (multi-goal_spread.py:121-138)
            const Real tx = Op::sub(px, lx), ty = Op::sub(py, ly);
            Real rew = Op::sub((Real)0, Op::sqrt(Op::add(Op::mul(tx, tx), Op::mul(ty, ty))));
            const uint32_t reach_me = rew >= neg_reach ? 1u : 0u;  // :126-129
            int hits = 0;
            {
                const Real cx = Op::sub(sub ? opx : px, sub ? px : opx), cy = Op::sub(sub ? opy : py, sub ? py : opy);
                const Real d2r = Op::add(Op::mul(cx, cx), Op::mul(cy, cy));
                if (!(d2r > near2) && Op::sqrt(d2r) < dist_min) {  // is_collision, :114-118: both agents pay, both count
                    rew = Op::sub(rew, (Real)1);
                    hits = 2;
                }
            }
            const Real rew_o = __shfl_xor_sync(0xFFFFFFFFu, rew, 1);
            const uint32_t reach_o = __shfl_xor_sync(0xFFFFFFFFu, reach_me, 1);
            const Real total = sub ? Op::add(rew_o, rew) : Op::add(rew, rew_o);   // np.sum in agent order, environment.py:107
            const uint32_t reach_bits = sub ? (reach_o | (reach_me << 1)) : (reach_me | (reach_o << 1));
            collisions += hits;
            reached = reach_bits;
            const bool done = (steps == p.max_steps) || (reach_bits == 3u);  // environment.py:118
            if (valid) {
                const int nd = GATHER ? p.n_dst : 1;
                for (int d = 0; d < nd; ++d) {
                    Real *rn = rn_ptr, *rw = rw_ptr;
                    uint8_t *dn = dn_ptr, *rc = rc_ptr;
                    int32_t *cl = cl_ptr;
                    if (GATHER && d > 0) {
                        rn = reinterpret_cast<Real *>(p.out[d].reward_n + (reinterpret_cast<char *>(rn_ptr) - o0.reward_n));
                        rw = reinterpret_cast<Real *>(p.out[d].reward + (reinterpret_cast<char *>(rw_ptr) - o0.reward));
                        dn = p.out[d].done + (dn_ptr - o0.done);
                        cl = p.out[d].collisions + (cl_ptr - o0.collisions);
                        rc = p.out[d].reached + (rc_ptr - o0.reached);
                    }
                    if (has_rn) *rn = rew;   // consecutive lanes -> consecutive words
                    if (sub == 0) {
                        if (has_rw) *rw = total;
                        if (has_dn) *dn = done ? 1 : 0;
                        if (has_cl) *cl = collisions;
                        if (has_rc) *rc = (uint8_t)reach_bits;
                    }
                }
            }
            rn_ptr += OB * N; rw_ptr += OB; dn_ptr += OB; cl_ptr += OB; rc_ptr += OB;
            if (p.auto_reset && done) reset_state((unsigned long long)(p.t0 + t + 1), kTagAutoReset);
        }
        emit(t);
        if (sel && sub == 0 && o0.done != nullptr) o0.done[oe0 + env] = 0;  // np.any(done_n), environment.py:149
    }

    if (valid) {
        st4<Real>(reinterpret_cast<Real *>(p.sv) + ((size_t)env * N + sub) * 4, vx, vy, px, py);
        if (reset_mode || p.auto_reset) {
            Real *lm = reinterpret_cast<Real *>(p.landmarks) + ((size_t)env * N + sub) * 2;
            lm[0] = lx; lm[1] = ly;
        }
        if (sub == 0) {
            p.steps[env] = steps;
            p.collisions[env] = collisions;
            p.reached[env] = (uint8_t)reached;
        }
    }
    ticket.publish(lane);
    if (pending && leader) bulk_wait_read();
}

// ------------------------------------------------------------------------ host side

template <typename Real, bool GATHER, bool FULL>
static int launch_pair(const PtParams &p, cudaStream_t stream) {
    auto kern = particle_pair_kernel<Real, GATHER, FULL>;
    constexpr int kSmem = PairGeom<Real>::kSmemBytes;
    static std::atomic<uint64_t> attr_done{0};
    CM3_CUDA(ensure_smem_attr(kern, attr_done));
    const int nblocks = (p.B + kPairEnvs - 1) / kPairEnvs;
    const int smem_launch = (p.mode == kPtStep && p.T > 1) ? balance_waves((const void *)kern, kWarp, kSmem, nblocks) : kSmem;
    const int parts = p.chained ? chain_parts(nblocks) : 1;
    for (int i = 0; i < parts; ++i) {
        PtParams q = p;
        q.tile0 = (int)((long long)nblocks * i / parts);
        const int n = (int)((long long)nblocks * (i + 1) / parts) - q.tile0;
        if (n > 0) CM3_CUDA(launch_kernel(kern, n, kWarp, smem_launch, stream, pdl_enabled(), q));
    }
    return CM3_OK;
}

template <typename Real>
static int dispatch_pair(const PtParams &p, cudaStream_t stream) {
    if (p.n_dst > 1) return launch_pair<Real, true, false>(p, stream);
    const PtOut &o = p.out[0];
    auto aligned = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool full = p.mode == kPtStep && p.mass == 1.0 && p.B % kPairEnvs == 0 && p.out_B % kPairEnvs == 0 &&
                      p.out_env0 % kPairEnvs == 0 && o.global_state && o.obs_self && o.obs_others && o.reward && o.reward_n &&
                      o.done && aligned(o.global_state) && aligned(o.obs_self) && aligned(o.obs_others) && full_enabled();
    return full ? launch_pair<Real, false, true>(p, stream) : launch_pair<Real, false, false>(p, stream);
}

int particle_pair_launch(int real, const PtParams &p, cudaStream_t stream) {
    return real == CM3_REAL_F64 ? dispatch_pair<double>(p, stream) : dispatch_pair<float>(p, stream);
}

}  // namespace cm3
