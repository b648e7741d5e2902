// particle_duo.cu - EXPERIMENT, off by default (CM3_PT_DUO=1 enables it): the common launch of the one-
// and two-agent particle envs (stage 1 and "merge", alg/config_particle_stage{1,2_merge}.json) with
// TWO ENVS PER THREAD, a warp owning 64 envs.
//
// The idea.  particle.cu at 65 536 envs x 2 agents is not bandwidth-bound (0.63 of the HBM roofline;
// 0.90 at 262 144): ncu shows ~420 instructions per warp-step on 3.5 warps per scheduler, 58 % of the
// issue slots used, the stall being the fixed-latency dependency wait.  Most of those instructions are
// what a step costs per WARP, not per env - the action stream hand-over, the staging fence, three TMA
// issues with their operand loops, output cursors, loop control - so giving a lane LESS work (one
// lane per agent, particle_pair.cu) doubled the instruction count and lost.  This is the mirror image:
// a lane steps two envs (env0 + lane and env0 + 32 + lane), phase by phase so that the two dependent
// chains can interleave in the one instruction stream, and the per-warp overhead is paid once per 64
// envs: the two 32-env sub-tiles are contiguous in every output array, so they are staged side by
// side and leave as ONE tensor-map store per field.
//
// The measurement (profiles/r02k_ab.txt, r02l_ab.txt).  Bit-identical to particle.cu
// (tests/test_gpu_round2.py::test_duo_kernel_is_bit_identical...), ~30 % fewer instructions per env -
// and slower: 0.48 vs 0.63 of the roofline at 65 536 envs (0.88 vs 0.91 at 262 144), env by env or
// phase by phase alike.  Half the warps (7 per SM) is what decides it: the contact evaluation, the
// IEEE square roots and divisions are branchy library sequences that ptxas does not interleave across
// the two envs, so a warp still runs two chains back to back where two warps ran them side by side.
// With particle_pair.cu (twice the warps: issue-bound) this brackets the one-env-per-thread mapping
// from both sides; it stays the default.  Kept as a validated variant.
//
// Scope: the FULL case only - float, step / rollout mode, every reference output requested, whole
// 64-env tiles, unit mass, one destination; everything else (reset launches, ragged batches, double,
// gather) takes particle.cu.  The per-tile chaining words are those of the 32-env tiling (two tickets
// per block), so launches of the two kernels can alternate on one state.
//
// What it computes (reference file:line): see particle.cu - _set_action (environment.py:177-225),
// World.step (core.py:117-196), Scenario.reward / observation / done / reset_world
// (multi-goal_spread.py:65-154), MultiAgentEnv.step tail (environment.py:95-123).
#include "particle_math.cuh"

namespace cm3 {

constexpr int kDuoEnvs = 2 * kWarp;  // envs per warp

template <int N>
struct DuoGeom {
    using G = PtGeom<N, float>;
    static constexpr int kRowBytes = 2 * G::kRowBytes, kOthBytes = 2 * G::kOthBytes;  // two sub-tiles side by side
    static constexpr int kOthOff = round_up(kRowBytes, 1024);
    static constexpr int kSetBytes = round_up(kOthOff + kOthBytes, 1024);
    static constexpr int kStages = 2;
    static constexpr int kActOff = kStages * kSetBytes;
    static constexpr int kSmemBytes = kActOff + ActionStream<N>::kSmemBytes + 1024 /* alignment slack */;
    static_assert(2 * G::kRowRows <= 256 && 2 * G::kOthRows <= 256, "TMA box dimension");
    static_assert(kDuoEnvs * N <= ActionStream<N>::kSlotBytes && kDuoEnvs * N / 4 <= kWarp, "action rows of a tile");
};

template <int N>
__global__ void __launch_bounds__(kWarp, 1) particle_duo_kernel(const __grid_constant__ PtParams p) {
    using Real = float;
    using Op = RealOps<float>;
    using G = PtGeom<N, float>;
    using Gm = DuoGeom<N>;
    constexpr int E = 2;
    constexpr int NO = G::NO, LO = G::LO;
    constexpr uint32_t RS = 4;

    const int lane = threadIdx.x;
    const bool leader = elect_one();
    const int tile = p.tile0 + (int)blockIdx.x;   // 64-env tile = 32-env tiles 2 * tile and 2 * tile + 1
    TileTicket ticket[E];
#pragma unroll
    for (int q = 0; q < E; ++q) ticket[q].take(p.sync, 2 * tile + q, lane);
    if ((ticket[0].mine & ticket[1].mine) != 0xFFFFFFFFu) pdl_launch_dependents();

    extern __shared__ unsigned char smem_duo[];
    unsigned char *stage_base = smem_duo + ((1024u - (smem_u32(smem_duo) & 1023u)) & 1023u);

    const int env0 = tile * kDuoEnvs;
    const size_t B = (size_t)p.B;
    const PtConsts<float> &K = p.kf;
    const Real dt = K.dt, keep = K.keep, dist_min = K.dist_min, sens = K.sensitivity, neg_reach = K.neg_reach;
    const Real far2 = K.far2, near2 = K.near2;

    const PtOut &o0 = p.out[0];
    if (leader) {
        tma_prefetch_map(&p.tm.oo); tma_prefetch_map(&p.tm.gs); tma_prefetch_map(&p.tm.os);
    }
    const size_t OB = (size_t)p.out_B, oe0 = (size_t)p.out_env0;
    const bool has_cl = o0.collisions != nullptr, has_rc = o0.reached != nullptr;
    // per-thread output cursors of slot t = 0 for my first env; the second one is 32 envs further
    Real *rn_ptr = reinterpret_cast<Real *>(o0.reward_n) + (oe0 + env0 + lane) * N;
    Real *rw_ptr = reinterpret_cast<Real *>(o0.reward) + (oe0 + env0 + lane);
    uint8_t *dn_ptr = o0.done + (oe0 + env0 + lane);
    int32_t *cl_ptr = o0.collisions + (oe0 + env0 + lane);
    uint8_t *rc_ptr = o0.reached + (oe0 + env0 + lane);
    const int tile_idx = (int)((oe0 + env0) / kDuoEnvs);
    const int tiles_per_slot = (int)(OB / kDuoEnvs);

    if (p.chained) {
#pragma unroll
        for (int q = 0; q < E; ++q) ticket[q].wait(lane);
    } else {
        pdl_wait();
    }

    // ---- state of my two envs
    Real vx[E][N], vy[E][N], px[E][N], py[E][N], lx[E][N], ly[E][N];
    int steps[E], collisions[E];
    uint32_t reached[E];
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int env = env0 + q * kWarp + lane;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ld4<Real>(reinterpret_cast<const Real *>(p.sv) + ((size_t)env * N + i) * 4, vx[q][i], vy[q][i], px[q][i], py[q][i]);
            ld2<Real>(reinterpret_cast<const Real *>(p.landmarks) + ((size_t)env * N + i) * 2, lx[q][i], ly[q][i]);
        }
        steps[q] = __ldcg(p.steps + env);
        collisions[q] = __ldcg(p.collisions + env);
        reached[q] = __ldcg(p.reached + env);
    }

    bool pending = false;
    ActionStream<N> acts;
    acts.init(stage_base + Gm::kActOff, p.actions, p.B, env0, kDuoEnvs, true, p.T, lane);
    uint32_t act_word[E] = {0u, 0u};
    if (acts.on) {
        act_word[0] = acts.begin(lane);
        act_word[1] = acts.read(0, kWarp + lane);
    }

    // observations of the current state of both envs -> outputs of slot t: one store per field
    auto emit = [&](int t) {
        const uint32_t act_loaded = acts.on ? acts.load(t + 1) : 0u;
        if (acts.on) acts.prefetch(t + 3);
        unsigned char *stage_row = stage_base + (t & 1) * Gm::kSetBytes;
        unsigned char *stage_oo = stage_row + Gm::kOthOff;
        if (pending && leader) bulk_wait_read_but_one();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < E; ++q) {
            unsigned char *row = stage_row + q * G::kRowBytes, *oth = stage_oo + q * G::kOthBytes;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                stage4<Real>(row, (uint32_t)(lane * N + i) * 4u * RS, G::kRowMask, vx[q][i], vy[q][i], px[q][i], py[q][i]);
                const uint32_t oo = (uint32_t)(lane * N + i) * (uint32_t)LO * RS;
                if (N == 1) {
                    stage4<Real>(oth, oo, G::kOthMask, Op::sub(vx[q][0], vx[q][0]), Op::sub(vy[q][0], vy[q][0]),
                                 Op::sub(px[q][0], px[q][0]), Op::sub(py[q][0], py[q][0]));
                } else {
#pragma unroll
                    for (int k = 0; k < NO; ++k) {
                        const int j = k + (k >= i ? 1 : 0);
                        stage4<Real>(oth, oo + 4u * k * RS, G::kOthMask, Op::sub(vx[q][j], vx[q][i]), Op::sub(vy[q][j], vy[q][i]),
                                     Op::sub(px[q][j], px[q][i]), Op::sub(py[q][j], py[q][i]));
                    }
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (leader) {
            const int tl = tile_idx + t * tiles_per_slot;
            const uint64_t pol = l2_policy_evict_first();
            tma_store_2d_hint(&p.tm.oo, stage_oo, 0, tl * (2 * G::kOthRows), pol);
            tma_store_2d_hint(&p.tm.gs, stage_row, 0, tl * (2 * G::kRowRows), pol);
            tma_store_2d_hint(&p.tm.os, stage_row, 0, tl * (2 * G::kRowRows), pol);
            bulk_commit();
        }
        pending = true;
        if (acts.on) {
            act_word[0] = acts.hand_over(t, act_loaded, lane);
            act_word[1] = acts.read(t + 1, kWarp + lane);
        }
    };

    for (int t = 0; t < p.T; ++t) {
        // Phase by phase over BOTH envs, not env by env: within a phase the two envs are independent
        // straight-line code, which is what lets ptxas interleave their dependent chains (an in-order
        // warp gets no instruction-level parallelism across a branch, and the per-env contact branch
        // of the first version of this kernel serialised the envs: 0.48 against particle.cu's 0.64).
        // ---- actions -> control forces (environment.py:194-214, core.py:134-140)
        Real fx[E][N], fy[E][N];
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int env = env0 + q * kWarp + lane;
            int act[N];
            if (p.actions != nullptr) {
                uint32_t w = act_word[q];
                if (!acts.on) w = load_actions_packed<N>(p.actions + ((size_t)t * B + env) * N);
#pragma unroll
                for (int i = 0; i < N; ++i) act[i] = unpack_action(w, i);
            } else {
                const Philox4 w = philox_action_words(p.seed, (uint64_t)(p.env_id_offset + env), (uint64_t)(p.t0 + t));
#pragma unroll
                for (int i = 0; i < N; ++i) act[i] = action_from_word(philox_word(w, i), 5);
            }
            if (p.actions_out != nullptr) {
#pragma unroll
                for (int i = 0; i < N; ++i) p.actions_out[((size_t)t * B + env) * N + i] = (int8_t)act[i];
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const Real ux = (act[i] == 1) ? (Real)-1 : (act[i] == 2) ? (Real)1 : (Real)0;
                const Real uy = (act[i] == 3) ? (Real)-1 : (act[i] == 4) ? (Real)1 : (Real)0;
                fx[q][i] = Op::mul(ux, sens); fy[q][i] = Op::mul(uy, sens);
            }
        }
        // ---- the one agent pair of each env (core.py:143-155, 180-196).  One branch for the warp; under
        // it both envs are evaluated and the result is SELECTED per env, so a far pair keeps its
        // p_force bit for bit (see particle.cu for the exact-zero skip) and there is no per-env branch.
        if (N == 2) {
            bool near[E];
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const Real ex = Op::sub(px[q][0], px[q][N - 1]), ey = Op::sub(py[q][0], py[q][N - 1]);
                near[q] = !(Op::add(Op::mul(ex, ex), Op::mul(ey, ey)) > far2);  // near, coincident or NaN
            }
            if (near[0] || near[1]) {
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const Force2<Real> F = contact_force_inl<Real>(px[q][0], py[q][0], px[q][N - 1], py[q][N - 1], p);
                    fx[q][0] = near[q] ? Op::add(F.x, fx[q][0]) : fx[q][0];
                    fy[q][0] = near[q] ? Op::add(F.y, fy[q][0]) : fy[q][0];
                    fx[q][N - 1] = near[q] ? Op::add(-F.x, fx[q][N - 1]) : fx[q][N - 1];
                    fy[q][N - 1] = near[q] ? Op::add(-F.y, fy[q][N - 1]) : fy[q][N - 1];
                }
            }
        }
        // ---- integrate (core.py:158-169; mass == 1)
#pragma unroll
        for (int q = 0; q < E; ++q) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                vx[q][i] = Op::mul(vx[q][i], keep); vy[q][i] = Op::mul(vy[q][i], keep);
                vx[q][i] = Op::add(vx[q][i], Op::mul(fx[q][i], dt)); vy[q][i] = Op::add(vy[q][i], Op::mul(fy[q][i], dt));
                px[q][i] = Op::add(px[q][i], Op::mul(vx[q][i], dt)); py[q][i] = Op::add(py[q][i], Op::mul(vy[q][i], dt));
            }
            steps[q] += 1;
        }
        // ---- rewards, reached, collisions (multi-goal_spread.py:114-138), done (environment.py:118)
        bool done[E];
#pragma unroll
        for (int q = 0; q < E; ++q) {
            Real rew[N];
            uint32_t reach_bits = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const Real tx = Op::sub(px[q][i], lx[q][i]), ty = Op::sub(py[q][i], ly[q][i]);
                rew[i] = Op::sub((Real)0, Op::sqrt(Op::add(Op::mul(tx, tx), Op::mul(ty, ty))));
                reach_bits |= (rew[i] >= neg_reach ? 1u : 0u) << i;
            }
            int hits = 0;
            if (N == 2) {
                const Real cx = Op::sub(px[q][0], px[q][N - 1]), cy = Op::sub(py[q][0], py[q][N - 1]);
                const Real d2r = Op::add(Op::mul(cx, cx), Op::mul(cy, cy));
                // same decision as particle.cu's (d2r <= near2 && sqrt(d2r) < dist_min): near2 > dist_min^2,
                // so the first test is implied whenever the second holds; evaluated without a branch
                const bool hit = !(d2r > near2) && (Op::sqrt(d2r) < dist_min);
                rew[0] = hit ? Op::sub(rew[0], (Real)1) : rew[0];
                rew[N - 1] = hit ? Op::sub(rew[N - 1], (Real)1) : rew[N - 1];
                hits = hit ? 2 : 0;
            }
            Real total = rew[0];
#pragma unroll
            for (int i = 1; i < N; ++i) total = Op::add(total, rew[i]);
            collisions[q] += hits;
            reached[q] = reach_bits;
            done[q] = (steps[q] == p.max_steps) || (reach_bits == (1u << N) - 1u);
            Real *rn = rn_ptr + q * kWarp * N;
#pragma unroll
            for (int i = 0; i < N; ++i) rn[i] = rew[i];
            rw_ptr[q * kWarp] = total;
            dn_ptr[q * kWarp] = done[q] ? 1 : 0;
            if (has_cl) cl_ptr[q * kWarp] = collisions[q];
            if (has_rc) rc_ptr[q * kWarp] = (uint8_t)reach_bits;
        }
        // ---- in-kernel episode reset (rare): multi-goal_spread.py:65-93 on Philox
#pragma unroll
        for (int q = 0; q < E; ++q) {
            if (p.auto_reset && done[q]) {
                const int env = env0 + q * kWarp + lane;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const ResetDraw<Real> d = draw_reset<Real>(p, (unsigned long long)(p.env_id_offset + env),
                                                               (unsigned long long)(p.t0 + t + 1), kTagAutoReset, i);
                    px[q][i] = d.px; py[q][i] = d.py; lx[q][i] = d.lx; ly[q][i] = d.ly;
                    vx[q][i] = 0; vy[q][i] = 0;
                }
                steps[q] = 0; collisions[q] = 0; reached[q] = 0;
            }
        }
        rn_ptr += OB * N; rw_ptr += OB; dn_ptr += OB; cl_ptr += OB; rc_ptr += OB;
        emit(t);
    }

#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int env = env0 + q * kWarp + lane;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            st4<Real>(reinterpret_cast<Real *>(p.sv) + ((size_t)env * N + i) * 4, vx[q][i], vy[q][i], px[q][i], py[q][i]);
            if (p.auto_reset) {
                Real *lm = reinterpret_cast<Real *>(p.landmarks) + ((size_t)env * N + i) * 2;
                lm[0] = lx[q][i]; lm[1] = ly[q][i];
            }
        }
        p.steps[env] = steps[q];
        p.collisions[env] = collisions[q];
        p.reached[env] = (uint8_t)reached[q];
    }
#pragma unroll
    for (int q = 0; q < E; ++q) ticket[q].publish(lane);
    if (pending && leader) bulk_wait_read();
}

// ------------------------------------------------------------------------ host side

template <int N>
static int launch_duo(PtParams p, cudaStream_t stream) {
    using G = PtGeom<N, float>;
    auto kern = particle_duo_kernel<N>;
    constexpr int kSmem = DuoGeom<N>::kSmemBytes;
    const size_t envs = (size_t)p.T * (size_t)p.out_B;
    const PtOut &o = p.out[0];
    // one box = one 64-env tile: twice the rows of particle.cu's 32-env box, same swizzle
    if (!encode_tile_map(&p.tm.gs, o.global_state, envs * N * 4 * sizeof(float), G::kRowSw, G::kRowW, 2 * G::kRowRows) ||
        !encode_tile_map(&p.tm.os, o.obs_self, envs * N * 4 * sizeof(float), G::kRowSw, G::kRowW, 2 * G::kRowRows) ||
        !encode_tile_map(&p.tm.oo, o.obs_others, envs * N * G::LO * sizeof(float), G::kOthSw, G::kOthW, 2 * G::kOthRows))
        return kDuoNotMine;  // not encodable: the caller falls back to particle.cu
    static std::atomic<uint64_t> attr_done{0};
    CM3_CUDA(ensure_smem_attr(kern, attr_done));
    const int nblocks = p.B / kDuoEnvs;
    const int smem_launch = p.T > 1 ? balance_waves((const void *)kern, kWarp, kSmem, nblocks) : kSmem;
    p.tma = 1;
    CM3_CUDA(launch_kernel(kern, nblocks, kWarp, smem_launch, stream, pdl_enabled(), p));
    return CM3_OK;
}

// returns kDuoNotMine when the launch is not this kernel's case (the caller then takes particle.cu)
int particle_duo_launch(int N, int real, const PtParams &p, cudaStream_t stream) {
    const PtOut &o = p.out[0];
    const bool mine = (N == 1 || N == 2) && real == CM3_REAL_F32 && p.n_dst == 1 && p.mode == kPtStep && p.mass == 1.0 &&
                      p.B % kDuoEnvs == 0 && p.out_B % kDuoEnvs == 0 && p.out_env0 % kDuoEnvs == 0 && o.global_state &&
                      o.obs_self && o.obs_others && o.reward && o.reward_n && o.done && tma_enabled() && full_enabled() &&
                      chain_parts(p.B / kDuoEnvs) == 1;
    if (!mine) return kDuoNotMine;
    return N == 1 ? launch_duo<1>(p, stream) : launch_duo<2>(p, stream);
}

}  // namespace cm3
