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
template <typename Real> __device__ __forceinline__ void st4(Real *p, Real a, Real b, Real c, Real d);
template <> __device__ __forceinline__ void st4<float>(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void st4<double>(double *p, double a, double b, double c, double d) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2 *>(p)[1] = make_double2(c, d);
}

template <typename Real, int N> __device__ __forceinline__ Real pickr(const Real (&v)[N], int idx) {
    Real r = v[0];
#pragma unroll
    for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
    return r;
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

template <int N, typename Real>
__global__ void __launch_bounds__(kPtThreads) particle_kernel(const __grid_constant__ PtParams p) {
    using Op = RealOps<Real>;
    constexpr int NP = (N == 3) ? 4 : N;
    constexpr int EW = kWarp / NP;
    constexpr int LO = 4 * (N > 1 ? N - 1 : 1);
    constexpr unsigned kFullMask = 0xFFFFFFFFu;

    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * kPtThreads + threadIdx.x) >> 5;
    const int env0 = gwarp * EW;
    if (env0 >= p.B) return;  // warp-uniform
    const int e = lane / NP, a = lane % NP;
    const int gb = lane - a;  // first lane of my env's group
    const int env = env0 + e;
    const bool valid = (a < N) && (env < p.B);
    const size_t B = (size_t)p.B;

    const Real dt = (Real)p.dt, keep = (Real)(1.0 - p.damping), cf = (Real)p.contact_force;
    const Real km = (Real)p.contact_margin, dist_min = (Real)p.dist_min, mass = (Real)p.mass;
    const Real sens = (Real)p.sensitivity, neg_reach = (Real)(-p.reach_thresh);

    // ---- state
    Real vx = 0, vy = 0, px = (Real)(2 * lane), py = 0, lx = 0, ly = 0;  // idle lanes stay apart
    int steps = 0, collisions = 0;
    uint32_t reached = 0;
    if (valid) {
        ld4<Real>(reinterpret_cast<const Real *>(p.sv) + ((size_t)env * N + a) * 4, vx, vy, px, py);
        const Real *lm = reinterpret_cast<const Real *>(p.landmarks) + ((size_t)env * N + a) * 2;
        lx = lm[0]; ly = lm[1];
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
            const unsigned long long genv = (unsigned long long)(p.env_id_offset + env);
            const uint32_t c0 = (uint32_t)genv, c1 = (uint32_t)(genv >> 32);
            const uint32_t c2 = (uint32_t)counter, c3 = kTagReset | ((uint32_t)(counter >> 32) & 0xFFFFu);
            const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
            const Philox4 wb = philox4x32_10(c0, c1 ^ 0xF0000000u, c2, c3, k0, k1);  // rand_num, :75
            const Philox4 wa = philox4x32_10(c0, c1 ^ ((uint32_t)(a + 1) << 24), c2, c3, k0, k1);
            if (u01_24(wb.x) < p.prob_random) {  // :77-78, :88-89
                px = (Real)(-1.0 + 2.0 * u01_24(wa.x)); py = (Real)(-1.0 + 2.0 * u01_24(wa.y));
                lx = (Real)(-1.0 + 2.0 * u01_24(wa.z)); ly = (Real)(-1.0 + 2.0 * u01_24(wa.w));
            } else {  // :80-83, :91
                Real nx = 0, ny = 0;
                if (p.initial_std != 0.0) {  // Box-Muller on (0,1] x [0,1)
                    const Real r0 = Op::sqrt(Op::mul((Real)-2, Op::log((Real)(u01_24(wa.x) + 1.0 / 16777216.0))));
                    const Real r1 = Op::sqrt(Op::mul((Real)-2, Op::log((Real)(u01_24(wa.z) + 1.0 / 16777216.0))));
                    Real s0, c0f, s1, c1f;
                    Op::sincospi((Real)(2.0 * u01_24(wa.y)), &s0, &c0f);
                    Op::sincospi((Real)(2.0 * u01_24(wa.w)), &s1, &c1f);
                    nx = Op::mul(Op::mul(r0, c0f), (Real)p.initial_std);
                    ny = Op::mul(Op::mul(r1, c1f), (Real)p.initial_std);
                }
                const int ai = a < N ? a : 0;
                px = Op::add((Real)p.agents_x[ai], nx); py = Op::add((Real)p.agents_y[ai], ny);
                lx = (Real)p.landmarks_x[ai]; ly = (Real)p.landmarks_y[ai];
            }
        }
        vx = 0; vy = 0; steps = 0; collisions = 0; reached = 0;  // :84-86, :93; environment.py:148
    };

    // observations of the current state -> outputs of slot t (multi-goal_spread.py:145-154,
    // environment.py:113-116)
    auto emit = [&](int t) {
        Real avx[N], avy[N], apx[N], apy[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            avx[j] = __shfl_sync(kFullMask, vx, gb + j); avy[j] = __shfl_sync(kFullMask, vy, gb + j);
            apx[j] = __shfl_sync(kFullMask, px, gb + j); apy[j] = __shfl_sync(kFullMask, py, gb + j);
        }
        if (!valid) return;
        const size_t rec = ((size_t)t * B + env) * N + a;
        if (p.global_state != nullptr) st4<Real>(reinterpret_cast<Real *>(p.global_state) + rec * 4, vx, vy, px, py);
        if (p.obs_self != nullptr) st4<Real>(reinterpret_cast<Real *>(p.obs_self) + rec * 4, vx, vy, px, py);
        if (p.obs_others != nullptr) {
            Real *oo = reinterpret_cast<Real *>(p.obs_others) + rec * LO;
            if (N == 1) {
                st4<Real>(oo, Op::sub(vx, vx), Op::sub(vy, vy), Op::sub(px, px), Op::sub(py, py));
            } else {
#pragma unroll
                for (int k = 0; k < N - 1; ++k) {
                    const int j = k + (k >= a ? 1 : 0);
                    st4<Real>(oo + 4 * k, Op::sub(pickr<Real, N>(avx, j), vx), Op::sub(pickr<Real, N>(avy, j), vy),
                              Op::sub(pickr<Real, N>(apx, j), px), Op::sub(pickr<Real, N>(apy, j), py));
                }
            }
        }
    };

    const int T_eff = (p.mode == kPtReset) ? 1 : p.T;
    for (int t = 0; t < T_eff; ++t) {
        bool sel = false;
        if (p.mode == kPtReset) {
            sel = valid && (p.env_mask == nullptr || p.env_mask[env] != 0);
            if (sel) reset_state((unsigned long long)p.reset_counter);
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

            // ---- contact forces, other agents in index order (core.py:143-155, 180-196)
            if (N > 1) {
                Real opx[N], opy[N];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    opx[j] = __shfl_sync(kFullMask, px, gb + j);
                    opy[j] = __shfl_sync(kFullMask, py, gb + j);
                }
#pragma unroll
                for (int k = 0; k < N - 1; ++k) {
                    const int j = k + (k >= a ? 1 : 0);
                    Real dx, dy, dist, x;
                    Contact<Real>::eval(px, py, pickr<Real, N>(opx, j), pickr<Real, N>(opy, j), p.dist_min,
                                        p.contact_margin, dx, dy, dist, x);
                    const Real pen = Op::mul(logaddexp0<Real>(x), km);
                    fx = Op::add(Op::mul(Op::div(Op::mul(cf, dx), dist), pen), fx);
                    fy = Op::add(Op::mul(Op::div(Op::mul(cf, dy), dist), pen), fy);
                }
            }
            // ---- integrate (core.py:158-169)
            vx = Op::mul(vx, keep); vy = Op::mul(vy, keep);
            vx = Op::add(vx, Op::mul(Op::div(fx, mass), dt)); vy = Op::add(vy, Op::mul(Op::div(fy, mass), dt));
            px = Op::add(px, Op::mul(vx, dt)); py = Op::add(py, Op::mul(vy, dt));
            steps += 1;  // environment.py:93

            // ---- reward (multi-goal_spread.py:121-138)
            const Real tx = Op::sub(px, lx), ty = Op::sub(py, ly);
            Real rew = Op::sub((Real)0, Op::sqrt(Op::add(Op::mul(tx, tx), Op::mul(ty, ty))));
            const bool my_reached = rew >= neg_reach;
            int hits = 0;
            if (N > 1) {
                Real npx[N], npy[N];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    npx[j] = __shfl_sync(kFullMask, px, gb + j);
                    npy[j] = __shfl_sync(kFullMask, py, gb + j);
                }
#pragma unroll
                for (int k = 0; k < N - 1; ++k) {
                    const int j = k + (k >= a ? 1 : 0);
                    const Real dx = Op::sub(pickr<Real, N>(npx, j), px), dy = Op::sub(pickr<Real, N>(npy, j), py);
                    const Real dist = Op::sqrt(Op::add(Op::mul(dx, dx), Op::mul(dy, dy)));
                    if (dist < dist_min) { rew = Op::sub(rew, (Real)1); hits += 1; }
                }
            }
            // ---- env-level reductions over the NP lanes of my env
            Real total = 0;
            int all_hits = 0;
            uint32_t reach_bits = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const Real rj = __shfl_sync(kFullMask, rew, gb + j);
                total = (j == 0) ? rj : Op::add(total, rj);  // np.sum, environment.py:107
                all_hits += __shfl_sync(kFullMask, hits, gb + j);
                reach_bits |= (__shfl_sync(kFullMask, (int)my_reached, gb + j) ? 1u : 0u) << j;
            }
            collisions += all_hits;
            reached = reach_bits;
            const bool done = (steps == p.max_steps) || (reach_bits == (1u << N) - 1u);  // environment.py:118
            if (valid) {
                if (p.reward_n != nullptr) reinterpret_cast<Real *>(p.reward_n)[((size_t)t * B + env) * N + a] = rew;
                if (a == 0) {
                    if (p.reward != nullptr) reinterpret_cast<Real *>(p.reward)[(size_t)t * B + env] = total;
                    if (p.done != nullptr) p.done[(size_t)t * B + env] = done ? 1 : 0;
                }
            }
            if (p.auto_reset && done) reset_state((unsigned long long)(p.t0 + t + 1));
        }
        emit(t);
        if (sel && a == 0 && p.done != nullptr) p.done[env] = 0;  // np.any(done_n), environment.py:149
    }

    if (valid) {
        st4<Real>(reinterpret_cast<Real *>(p.sv) + ((size_t)env * N + a) * 4, vx, vy, px, py);
        Real *lm = reinterpret_cast<Real *>(p.landmarks) + ((size_t)env * N + a) * 2;
        lm[0] = lx; lm[1] = ly;
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
    particle_kernel<N, Real><<<nblocks, kPtThreads, 0, stream>>>(p);
    CM3_CUDA(cudaGetLastError());
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
