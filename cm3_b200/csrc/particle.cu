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
//  * one lane per (env, agent), agent-minor: lane = e*NP + a.  An env's agents sit in NP adjacent
//    lanes, so the all-pairs contact force, the collision penalty and the env-level reductions
//    (sum of rewards, all-reached, collision count) are width-NP warp shuffles - no shared memory;
//  * state rows are (vel, pos) = the reference's global_state rows: one 16-byte load and store per
//    lane, and consecutive lanes touch consecutive records: every global access of the kernel is
//    a fully coalesced 16-byte (or 4-byte) per-lane access;
//  * arithmetic follows the reference operation by operation with round-to-nearest intrinsics
//    (no FMA contraction, IEEE sqrt/div, no fast-math), in float (throughput mode) or double
//    (free-running parity mode);
//  * T steps can be fused in one launch with state in registers, Philox actions and in-kernel
//    episode reset, like the Checkers kernel.
#include "common.cuh"
#include "params.cuh"

namespace cm3 {

constexpr int kPtThreads = 128;

template <typename Real> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename Real> __device__ __forceinline__ void ld4(const Real *p, Real &a, Real &b, Real &c, Real &d);
template <> __device__ __forceinline__ void ld4<float>(const float *p, float &a, float &b, float &c, float &d) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
}
template <> __device__ __forceinline__ void ld4<double>(const double *p, double &a, double &b, double &c, double &d) {
    const double2 u = reinterpret_cast<const double2 *>(p)[0], v = reinterpret_cast<const double2 *>(p)[1];
    a = u.x; b = u.y; c = v.x; d = v.y;
}
template <typename Real> __device__ __forceinline__ void ld2(const Real *p, Real &a, Real &b);
template <> __device__ __forceinline__ void ld2<float>(const float *p, float &a, float &b) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void ld2<double>(const double *p, double &a, double &b) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    a = v.x; b = v.y;
}
template <typename Real> __device__ __forceinline__ const PtConsts<Real> &pt_consts(const PtParams &p);
template <> __device__ __forceinline__ const PtConsts<float> &pt_consts<float>(const PtParams &p) { return p.kf; }
template <> __device__ __forceinline__ const PtConsts<double> &pt_consts<double>(const PtParams &p) { return p.kd; }
template <typename Real> __device__ __forceinline__ void st4(Real *p, Real a, Real b, Real c, Real d);
template <> __device__ __forceinline__ void st4<float>(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void st4<double>(double *p, double a, double b, double c, double d) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2 *>(p)[1] = make_double2(c, d);
}

// np.logaddexp(0, x) - NumPy's npy_logaddexp with x1 = 0 (core.py:192)
template <typename Real> __device__ __forceinline__ Real logaddexp0(Real x) {
    using Op = RealOps<Real>;
    if (x == (Real)0) return (Real)0.693147180559945309417232121458176568;
    const Real tmp = Op::sub((Real)0, x);
    if (tmp > (Real)0) return Op::log1p(Op::exp(x));          // 0 + log1p(exp(-tmp))
    else if (tmp <= (Real)0) return Op::add(x, Op::log1p(Op::exp(tmp)));
    return tmp;  // NaN
}

// Contact geometry of one agent pair (core.py:186-192): delta, dist and the softplus argument
// x = -(dist - dist_min)/k.  k = 1e-3 makes x ill-conditioned - an ulp of dist is 1000 ulps of x -
// so the float kernel evaluates exactly this sub-expression in double from the (exact) float
// positions; everything else stays in Real.
template <typename Real> struct Contact;
template <> struct Contact<float> {
    static __device__ __forceinline__ void eval(float px, float py, float qx, float qy, double dist_min, double k,
                                                float &dx, float &dy, float &dist, float &x) {
        const double ddx = __dsub_rn((double)px, (double)qx), ddy = __dsub_rn((double)py, (double)qy);
        const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
        dx = (float)ddx; dy = (float)ddy; dist = (float)d;
        x = (float)(-__ddiv_rn(__dsub_rn(d, dist_min), k));
    }
};
template <> struct Contact<double> {
    static __device__ __forceinline__ void eval(double px, double py, double qx, double qy, double dist_min, double k,
                                                double &dx, double &dy, double &dist, double &x) {
        dx = __dsub_rn(px, qx); dy = __dsub_rn(py, qy);
        dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        x = -__ddiv_rn(__dsub_rn(dist, dist_min), k);
    }
};

// u in [0,1) with 24 random bits: exactly representable in float and double alike, so both
// precisions draw the same initial states
__device__ __forceinline__ double u01_24(uint32_t w) { return (double)(w >> 8) * (1.0 / 16777216.0); }

// reset_world() of one agent on Philox4x32-10 keyed by (seed; global env id, reset counter)
// (multi-goal_spread.py:75-91).  Deliberately NOT inlined: resets are rare (once per episode) and
// an inlined copy lets the compiler hoist ~300 instructions of Philox / Box-Muller arithmetic
// above the step loop, where every thread of every launch pays for them.
template <typename Real> struct ResetDraw { Real px, py, lx, ly; };

template <typename Real>
__device__ __noinline__ ResetDraw<Real> draw_reset(const PtParams &p, unsigned long long genv,
                                                   unsigned long long counter, int a) {
    using Op = RealOps<Real>;
    ResetDraw<Real> d;
    const uint32_t c0 = (uint32_t)genv, c1 = (uint32_t)(genv >> 32);
    const uint32_t c2 = (uint32_t)counter, c3 = kTagReset | ((uint32_t)(counter >> 32) & 0xFFFFu);
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    bool randomise = false;
    if (p.prob_random > 0.0) {  // rand_num, :75
        const Philox4 wb = philox4x32_10(c0, c1 ^ 0xF0000000u, c2, c3, k0, k1);
        randomise = u01_24(wb.x) < p.prob_random;
    }
    if (randomise) {  // :77-78, :88-89
        const Philox4 wa = philox4x32_10(c0, c1 ^ ((uint32_t)(a + 1) << 24), c2, c3, k0, k1);
        d.px = (Real)(-1.0 + 2.0 * u01_24(wa.x)); d.py = (Real)(-1.0 + 2.0 * u01_24(wa.y));
        d.lx = (Real)(-1.0 + 2.0 * u01_24(wa.z)); d.ly = (Real)(-1.0 + 2.0 * u01_24(wa.w));
    } else {  // :80-83, :91
        Real nx = 0, ny = 0;
        if (p.initial_std != 0.0) {  // Box-Muller on (0,1] x [0,1)
            const Philox4 wa = philox4x32_10(c0, c1 ^ ((uint32_t)(a + 1) << 24), c2, c3, k0, k1);
            const Real r0 = Op::sqrt(Op::mul((Real)-2, Op::log((Real)(u01_24(wa.x) + 1.0 / 16777216.0))));
            const Real r1 = Op::sqrt(Op::mul((Real)-2, Op::log((Real)(u01_24(wa.z) + 1.0 / 16777216.0))));
            Real s0, c0f, s1, c1f;
            Op::sincospi((Real)(2.0 * u01_24(wa.y)), &s0, &c0f);
            Op::sincospi((Real)(2.0 * u01_24(wa.w)), &s1, &c1f);
            nx = Op::mul(Op::mul(r0, c0f), (Real)p.initial_std);
            ny = Op::mul(Op::mul(r1, c1f), (Real)p.initial_std);
        }
        d.px = Op::add((Real)p.agents_x[a], nx); d.py = Op::add((Real)p.agents_y[a], ny);
        d.lx = (Real)p.landmarks_x[a]; d.ly = (Real)p.landmarks_y[a];
    }
    return d;
}

template <int N, typename Real>
__global__ void __launch_bounds__(kPtThreads) particle_kernel(const __grid_constant__ PtParams p) {
    using Op = RealOps<Real>;
    constexpr int NP = (N == 3) ? 4 : N;
    constexpr int EW = kWarp / NP;
    constexpr int NO = (N > 1) ? N - 1 : 1;  // "other" agents per agent
    constexpr int LO = 4 * NO;
    constexpr unsigned kFullMask = 0xFFFFFFFFu;

    pdl_launch_dependents();  // the next step's grid may become resident while this one drains

    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * kPtThreads + threadIdx.x) >> 5;
    const int env0 = gwarp * EW;
    if (env0 >= p.B) return;  // warp-uniform
    const int e = lane / NP, a = lane % NP;
    const int gb = lane - a;  // first lane of my env's group
    const int env = env0 + e;
    const bool valid = (a < N) && (env < p.B);
    const size_t B = (size_t)p.B;

    // world constants, rounded to Real once on the host (no per-thread F2F conversions)
    const PtConsts<Real> &K = pt_consts<Real>(p);
    const Real dt = K.dt, keep = K.keep, cf = K.contact_force, km = K.contact_margin, dist_min = K.dist_min;
    const Real mass = K.mass, sens = K.sensitivity, neg_reach = K.neg_reach;
    // squared distances beyond which a pair provably contributes exactly nothing (see below)
    const Real far2 = K.far2, near2 = K.near2;
    const bool unit_mass = (mass == (Real)1);  // x / 1 == x exactly: skip the IEEE division

    // lane of the k-th "other" agent of my env, in index order (multi-goal_spread.py:149-152)
    int src[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) src[k] = (N > 1) ? gb + k + (k >= a ? 1 : 0) : lane;

    pdl_wait();  // state written by the previous launch is visible from here on

    // ---- state
    Real vx = 0, vy = 0, px = (Real)(2 * lane), py = 0, lx = 0, ly = 0;  // idle lanes stay apart
    int steps = 0, collisions = 0;
    uint32_t reached = 0;
    if (valid) {
        ld4<Real>(reinterpret_cast<const Real *>(p.sv) + ((size_t)env * N + a) * 4, vx, vy, px, py);
        ld2<Real>(reinterpret_cast<const Real *>(p.landmarks) + ((size_t)env * N + a) * 2, lx, ly);
        steps = p.steps[env];
        collisions = p.collisions[env];
        reached = p.reached[env];
    }

    // multi-goal_spread.py:65-93 on Philox (or the injected host draws)
    auto reset_state = [&](unsigned long long counter) {
        if (p.init_pos != nullptr && p.mode == kPtReset) {
            const Real *ip = reinterpret_cast<const Real *>(p.init_pos) + ((size_t)env * N + a) * 2;
            const Real *il = reinterpret_cast<const Real *>(p.init_landmarks) + ((size_t)env * N + a) * 2;
            px = ip[0]; py = ip[1]; lx = il[0]; ly = il[1];
        } else {
            const ResetDraw<Real> d = draw_reset<Real>(p, (unsigned long long)(p.env_id_offset + env), counter,
                                                       a < N ? a : 0);
            px = d.px; py = d.py; lx = d.lx; ly = d.ly;
        }
        vx = 0; vy = 0; steps = 0; collisions = 0; reached = 0;  // :84-86, :93; environment.py:148
    };

    // (vel, pos) of the other agents of my env, straight from their lanes
    Real ovx[NO], ovy[NO], opx[NO], opy[NO];
    auto gather_others = [&]() {
#pragma unroll
        for (int k = 0; k < NO; ++k) {
            ovx[k] = __shfl_sync(kFullMask, vx, src[k]); ovy[k] = __shfl_sync(kFullMask, vy, src[k]);
            opx[k] = __shfl_sync(kFullMask, px, src[k]); opy[k] = __shfl_sync(kFullMask, py, src[k]);
        }
    };

    // observations of the current state -> outputs of slot t (multi-goal_spread.py:145-154,
    // environment.py:113-116); needs gather_others() of the current state
    const size_t OB = (size_t)p.out_B, oe0 = (size_t)p.out_env0;
    auto emit = [&](int t) {
        if (!valid) return;
        const size_t rec = ((size_t)t * OB + oe0 + env) * N + a;
        Real dvx[NO], dvy[NO], dpx[NO], dpy[NO];
#pragma unroll
        for (int k = 0; k < NO; ++k) {  // N == 1: "others" is the agent itself, :148-153
            dvx[k] = Op::sub(ovx[k], vx); dvy[k] = Op::sub(ovy[k], vy);
            dpx[k] = Op::sub(opx[k], px); dpy[k] = Op::sub(opy[k], py);
        }
        // n_dst > 1 (rollout_gather): the same records go to the rollout buffers of every GPU of
        // the node - peer memory over NVLink
        for (int d = 0; d < p.n_dst; ++d) {
            const PtOut &o = p.out[d];
            if (o.global_state != nullptr) st4<Real>(reinterpret_cast<Real *>(o.global_state) + rec * 4, vx, vy, px, py);
            if (o.obs_self != nullptr) st4<Real>(reinterpret_cast<Real *>(o.obs_self) + rec * 4, vx, vy, px, py);
            if (o.obs_others != nullptr) {
                Real *oo = reinterpret_cast<Real *>(o.obs_others) + rec * LO;
#pragma unroll
                for (int k = 0; k < NO; ++k) st4<Real>(oo + 4 * k, dvx[k], dvy[k], dpx[k], dpy[k]);
            }
        }
    };

    const int T_eff = (p.mode == kPtReset) ? 1 : p.T;
    for (int t = 0; t < T_eff; ++t) {
        bool sel = false;
        if (p.mode == kPtReset) {
            sel = valid && (p.env_mask == nullptr || p.env_mask[env] != 0);
            if (sel) reset_state((unsigned long long)p.reset_counter);
            gather_others();
        } else {
            // ---- action -> control force (environment.py:194-214, core.py:134-140)
            int act = 0;
            if (p.actions != nullptr) {
                act = valid ? (int)p.actions[((size_t)t * B + env) * N + a] : 0;
            } else {
                const Philox4 w = philox_action_words(p.seed, (uint64_t)(p.env_id_offset + env), (uint64_t)(p.t0 + t));
                act = action_from_word(philox_word(w, a & 3), 5);
            }
            if (p.actions_out != nullptr && valid) p.actions_out[((size_t)t * B + env) * N + a] = (int8_t)act;
            Real fx = (act == 1) ? (Real)-1 : (act == 2) ? (Real)1 : (Real)0;
            Real fy = (act == 3) ? (Real)-1 : (act == 4) ? (Real)1 : (Real)0;
            fx = Op::mul(fx, sens); fy = Op::mul(fy, sens);

            // ---- contact forces, other agents in index order (core.py:143-155, 180-196).
            // The reference evaluates the softplus penetration for EVERY pair at every distance;
            // beyond dist_min + ~110 k (float) / ~760 k (double) exp() underflows to exactly 0, so
            // pen == 0, the force is +-0 and "F + p_force" returns p_force bit for bit (p_force is
            // never -0: it starts as +0 or +-sensitivity).  Those pairs are skipped: same bits,
            // none of the double-precision sqrt / div / exp / log1p instructions.
            if (N > 1) {
#pragma unroll
                for (int k = 0; k < NO; ++k) {
                    const Real qx = __shfl_sync(kFullMask, px, src[k]), qy = __shfl_sync(kFullMask, py, src[k]);
                    const Real ex = Op::sub(px, qx), ey = Op::sub(py, qy);
                    const Real d2 = Op::add(Op::mul(ex, ex), Op::mul(ey, ey));
                    if (!(d2 > far2)) {  // near, coincident or NaN: the literal evaluation
                        Real dx, dy, dist, x;
                        Contact<Real>::eval(px, py, qx, qy, p.dist_min, p.contact_margin, dx, dy, dist, x);
                        const Real pen = Op::mul(logaddexp0<Real>(x), km);
                        fx = Op::add(Op::mul(Op::div(Op::mul(cf, dx), dist), pen), fx);
                        fy = Op::add(Op::mul(Op::div(Op::mul(cf, dy), dist), pen), fy);
                    }
                }
            }
            // ---- integrate (core.py:158-169)
            vx = Op::mul(vx, keep); vy = Op::mul(vy, keep);
            if (!unit_mass) { fx = Op::div(fx, mass); fy = Op::div(fy, mass); }
            vx = Op::add(vx, Op::mul(fx, dt)); vy = Op::add(vy, Op::mul(fy, dt));
            px = Op::add(px, Op::mul(vx, dt)); py = Op::add(py, Op::mul(vy, dt));
            steps += 1;  // environment.py:93
            gather_others();

            // ---- reward (multi-goal_spread.py:121-138)
            const Real tx = Op::sub(px, lx), ty = Op::sub(py, ly);
            Real rew = Op::sub((Real)0, Op::sqrt(Op::add(Op::mul(tx, tx), Op::mul(ty, ty))));
            const bool my_reached = rew >= neg_reach;
            int hits = 0;
            if (N > 1) {
#pragma unroll
                for (int k = 0; k < NO; ++k) {
                    const Real dx = Op::sub(opx[k], px), dy = Op::sub(opy[k], py);
                    const Real d2 = Op::add(Op::mul(dx, dx), Op::mul(dy, dy));
                    // sqrt is monotonic: d2 > near2 = (1.01 dist_min)^2 cannot round below dist_min
                    if (!(d2 > near2)) {
                        if (Op::sqrt(d2) < dist_min) { rew = Op::sub(rew, (Real)1); hits += 1; }
                    }
                }
            }
            // ---- env-level reductions over the NP lanes of my env
            Real total = 0;
            int all_hits = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const Real rj = __shfl_sync(kFullMask, rew, gb + j);
                total = (j == 0) ? rj : Op::add(total, rj);  // np.sum, environment.py:107
                if (N > 1) all_hits += __shfl_sync(kFullMask, hits, gb + j);
            }
            const uint32_t reach_bits = (__ballot_sync(kFullMask, my_reached) >> gb) & ((1u << N) - 1u);
            collisions += all_hits;
            reached = reach_bits;
            const bool done = (steps == p.max_steps) || (reach_bits == (1u << N) - 1u);  // environment.py:118
            if (valid) {
                const size_t orow = (size_t)t * OB + oe0 + env;
                for (int d = 0; d < p.n_dst; ++d) {
                    const PtOut &o = p.out[d];
                    if (o.reward_n != nullptr) reinterpret_cast<Real *>(o.reward_n)[orow * N + a] = rew;
                    if (a == 0) {
                        if (o.reward != nullptr) reinterpret_cast<Real *>(o.reward)[orow] = total;
                        if (o.done != nullptr) o.done[orow] = done ? 1 : 0;
                    }
                }
            }
            if (p.auto_reset) {
                if (done) reset_state((unsigned long long)(p.t0 + t + 1));
                if (__any_sync(kFullMask, done)) gather_others();
            }
        }
        emit(t);
        if (sel && a == 0 && p.out[0].done != nullptr) p.out[0].done[oe0 + env] = 0;  // np.any(done_n), environment.py:149
    }

    if (valid) {
        st4<Real>(reinterpret_cast<Real *>(p.sv) + ((size_t)env * N + a) * 4, vx, vy, px, py);
        if (p.mode == kPtReset || p.auto_reset) {  // landmarks only change on a reset
            Real *lm = reinterpret_cast<Real *>(p.landmarks) + ((size_t)env * N + a) * 2;
            lm[0] = lx; lm[1] = ly;
        }
        if (a == 0) {
            p.steps[env] = steps;
            p.collisions[env] = collisions;
            p.reached[env] = (uint8_t)reached;
        }
    }
}

template <int N, typename Real>
static int launch_pt(const PtParams &p, cudaStream_t stream) {
    constexpr int NP = (N == 3) ? 4 : N;
    constexpr int EW = kWarp / NP;
    const int nwarps = (p.B + EW - 1) / EW;
    const int nblocks = (nwarps + kPtThreads / kWarp - 1) / (kPtThreads / kWarp);
    CM3_CUDA(launch_kernel(particle_kernel<N, Real>, nblocks, kPtThreads, 0, stream, pdl_enabled(), p));
    return CM3_OK;
}

int particle_launch(int N, int real, const PtParams &p, cudaStream_t stream) {
#define CASE(n)                                                        \
    case n:                                                            \
        return real == CM3_REAL_F64 ? launch_pt<n, double>(p, stream) : launch_pt<n, float>(p, stream);
    switch (N) {
        CASE(1) CASE(2) CASE(3) CASE(4)
        default: break;
    }
#undef CASE
    set_error("no compiled particle kernel for n_agents=%d (1..%d supported)", N, CM3_MAX_AGENTS);
    return CM3_ERR_UNSUPPORTED;
}

}  // namespace cm3
