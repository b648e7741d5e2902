// particle_math.cuh - device helpers shared by the particle kernels (particle.cu: one thread per env;
// particle_pair.cu: one lane per agent for two-agent envs): vector loads / stores of the state rows,
// NumPy's logaddexp, the contact geometry and force of one agent pair (multiagent/core.py:180-196),
// the reset draws (multi-goal_spread.py:65-93 on Philox4x32-10) and the staging-tile store.
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace cm3 {

template <typename Real> __device__ __forceinline__ void ld4(const Real *p, Real &a, Real &b, Real &c, Real &d);
template <> __device__ __forceinline__ void ld4<float>(const float *p, float &a, float &b, float &c, float &d) {
    const float4 v = __ldcg(reinterpret_cast<const float4 *>(p));  // state: L2 (launch chaining reads it under an acquire)
    a = v.x; b = v.y; c = v.z; d = v.w;
}
template <> __device__ __forceinline__ void ld4<double>(const double *p, double &a, double &b, double &c, double &d) {
    const double2 u = __ldcg(reinterpret_cast<const double2 *>(p)), v = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    a = u.x; b = u.y; c = v.x; d = v.y;
}
template <typename Real> __device__ __forceinline__ void ld2(const Real *p, Real &a, Real &b);
template <> __device__ __forceinline__ void ld2<float>(const float *p, float &a, float &b) {
    const float2 v = __ldcg(reinterpret_cast<const float2 *>(p));
    a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void ld2<double>(const double *p, double &a, double &b) {
    const double2 v = __ldcg(reinterpret_cast<const double2 *>(p));
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

// np.logaddexp(0, x) - NumPy's npy_logaddexp with x1 = 0 (core.py:192):  x == 0 -> 0 + ln 2; tmp = 0 - x; tmp > 0 -> 0 + log1p(exp(-tmp)) = log1p(exp(x));
// tmp <= 0 -> x + log1p(exp(tmp)); NaN otherwise.  Both branches take log1p(exp(.)) of the same number
// -|x| (x itself when negative, the exact negation 0 - x when positive), so they are evaluated ONCE and
// the branch becomes a select: the same bits, but the lanes of a warp whose pairs sit on both sides of
// contact (x > 0 inside dist_min, x < 0 outside) no longer run two copies of exp + log1p back to back.
template <typename Real> __device__ __forceinline__ Real logaddexp0(Real x) {
    using Op = RealOps<Real>;
#ifdef CM3_LAE_BRANCHY  // round 1 / early round 2: the literal three-way branch (A/B builds only)
    if (x == (Real)0) return (Real)0.693147180559945309417232121458176568;
    const Real tmp = Op::sub((Real)0, x);
    if (tmp > (Real)0) return Op::log1p(Op::exp(x));
    else if (tmp <= (Real)0) return Op::add(x, Op::log1p(Op::exp(tmp)));
    return tmp;
#else
    const bool pos = x > (Real)0;
    const Real e = pos ? Op::sub((Real)0, x) : x;   // -|x| (NaN stays NaN)
    const Real l = Op::log1p(Op::exp(e));
    const Real r = pos ? Op::add(x, l) : l;
    return (x == (Real)0) ? (Real)0.693147180559945309417232121458176568 : r;
#endif
}

// Contact geometry of one agent pair (core.py:186-192): delta, dist and the softplus argument
// x = -(dist - dist_min)/k.  k = 1e-3 makes x ill-conditioned - an ulp of dist is 1000 ulps of x -
// so the float kernel cannot evaluate it in float.  Round 1 evaluated the literal expression in
// double (DSQRT + DDIV: two ~100-cycle dependent sequences on every in-contact pair, which is
// every step of the merge scenario).  The cancellation is all in (dist - dist_min); written as
//     dist - dist_min = (dist^2 - dist_min^2) / (dist + dist_min)
// it sits in the NUMERATOR, which is exact-ish in double from the exact float positions with five
// short double operations (sub, sub, mul, fma, sub), while the denominator (dist + dist_min) * k has
// no cancellation and is evaluated in float: x agrees with the double evaluation to ~2 float ulps
// (tests/test_gpu_particle.py compares against the float64 oracle at rtol 1e-5).
template <typename Real> struct Contact;
template <> struct Contact<float> {
    static __device__ __forceinline__ void eval(float px, float py, float qx, float qy, const PtParams &p,
                                                float &dx, float &dy, float &dist, float &x) {
        const double ddx = __dsub_rn((double)px, (double)qx), ddy = __dsub_rn((double)py, (double)qy);
        const double q = __fma_rn(ddx, ddx, __dmul_rn(ddy, ddy));   // dist^2, exact to a double ulp
        const float num = (float)__dsub_rn(q, p.dist_min2);
        dx = (float)ddx; dy = (float)ddy;
        dist = __fsqrt_rn((float)q);
        const float den = __fmul_rn(__fadd_rn(dist, p.kf.dist_min), p.kf.contact_margin);
        x = -__fdiv_rn(num, den);
    }
};
template <> struct Contact<double> {
    static __device__ __forceinline__ void eval(double px, double py, double qx, double qy, const PtParams &p,
                                                double &dx, double &dy, double &dist, double &x) {
        dx = __dsub_rn(px, qx); dy = __dsub_rn(py, qy);
        dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        x = -__ddiv_rn(__dsub_rn(dist, p.dist_min), p.contact_margin);
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
                                                   unsigned long long counter, uint32_t tag, int a) {
    using Op = RealOps<Real>;
    ResetDraw<Real> d;
    const uint32_t c0 = (uint32_t)genv, c1 = (uint32_t)(genv >> 32);
    const uint32_t c2 = (uint32_t)counter, c3 = tag | ((uint32_t)(counter >> 32) & 0xFFFFu);
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

// get_collision_force for one pair in (or near) contact, core.py:180-196.  Out of line where
// contacts are the rare path (N >= 3: for the same reason as draw_reset), inline for N <= 2, where
// the merge scenario takes it on nine steps out of ten.
template <typename Real> struct Force2 { Real x, y; };

template <typename Real>
__device__ __forceinline__ Force2<Real> contact_force_inl(Real ax, Real ay, Real bx, Real by, const PtParams &p) {
    using Op = RealOps<Real>;
    const PtConsts<Real> &K = pt_consts<Real>(p);
    Real dx, dy, dist, x;
    Contact<Real>::eval(ax, ay, bx, by, p, dx, dy, dist, x);
    const Real pen = Op::mul(logaddexp0<Real>(x), K.contact_margin);
    return Force2<Real>{Op::mul(Op::div(Op::mul(K.contact_force, dx), dist), pen),
                        Op::mul(Op::div(Op::mul(K.contact_force, dy), dist), pen)};
}
template <typename Real>
__device__ __noinline__ Force2<Real> contact_force_ool(Real ax, Real ay, Real bx, Real by, const PtParams &p) {
    return contact_force_inl<Real>(ax, ay, bx, by, p);
}

// Stores 4 Reals at byte offset `off` of a staging tile laid out in the TMA swizzle pattern
// `mask` (0 = linear).  The tile base is 1024-byte aligned.
template <typename Real> __device__ __forceinline__ void stage4(unsigned char *tile, uint32_t off, uint32_t mask,
                                                                Real a, Real b, Real c, Real d);
template <> __device__ __forceinline__ void stage4<float>(unsigned char *tile, uint32_t off, uint32_t mask,
                                                          float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(tile + (off ^ ((off >> 3) & mask))) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void stage4<double>(unsigned char *tile, uint32_t off, uint32_t mask,
                                                           double a, double b, double c, double d) {
    const uint32_t o1 = off + 16;
    *reinterpret_cast<double2 *>(tile + (off ^ ((off >> 3) & mask))) = make_double2(a, b);
    *reinterpret_cast<double2 *>(tile + (o1 ^ ((o1 >> 3) & mask))) = make_double2(c, d);
}

// 16-byte chunks per env record S = odd * 2^k: lanes one record apart collide on the 8 bank groups
// unless the 16-byte chunk index is XOR-ed with higher address bits; k = 0 needs no swizzle,
// k = 1 / 2 / >= 3 the TMA 32- / 64- / 128-byte swizzle (conflict-free for every S, checked
// exhaustively in tests/test_particle_layout.py).
__host__ __device__ constexpr int sw_bits_for(int chunks) {
    // (128-byte rows everywhere - fewer, wider TMA rows at the price of 2- to 3-way conflicts for
    // S = 6, 12 - measured within +-1.5 % of this choice: profiles/r01l_ab.txt, r01w)
    return (chunks % 2) ? 0 : (chunks % 4) ? 1 : (chunks % 8) ? 2 : 3;
}
__host__ __device__ constexpr int sw_row_bytes(int bits) { return bits ? (16 << bits) : 128; }

#ifndef CM3_PT_STAGE_MAXN
#define CM3_PT_STAGE_MAXN 2   // largest agent count with a double-buffered staging set; 3 fits as well and was measured
                              // again in round 2 (profiles/r02r_ab.txt): fused within noise, chained per-step 2 % slower
#endif
template <int N, typename Real>
struct PtGeom {
    static constexpr int NO = (N > 1) ? N - 1 : 1;  // "other" agents per agent
    static constexpr int LO = 4 * NO;
    static constexpr int kRowBytes = kWarp * N * 4 * (int)sizeof(Real);   // global_state / obs_self tile
    static constexpr int kOthBytes = kWarp * N * LO * (int)sizeof(Real);  // obs_others tile
    static constexpr int kRowSw = sw_bits_for(N * 4 * (int)sizeof(Real) / 16);
    static constexpr int kOthSw = sw_bits_for(N * LO * (int)sizeof(Real) / 16);
    static constexpr uint32_t kRowMask = ((1u << kRowSw) - 1u) << 4;  // bits [4, 4+sw) ^= bits [7, 7+sw)
    static constexpr uint32_t kOthMask = ((1u << kOthSw) - 1u) << 4;
    static constexpr int kRowW = sw_row_bytes(kRowSw), kOthW = sw_row_bytes(kOthSw);  // tensor-map row widths
    static constexpr int kRowRows = kRowBytes / kRowW, kOthRows = kOthBytes / kOthW;  // box rows per tile
    static constexpr int kOthOff = round_up(kRowBytes, 1024);  // swizzle patterns repeat every 1024 bytes
    // One staging set = row tile + others tile.  Two sets (double-buffered staging: the stores of
    // step t drain while step t + 1 is staged into the other set) for N <= 2, where a step is short
    // compared with the time a store takes to drain: measured +12 % for PM2, -3 % for PA3
    // (profiles/r01r_ab.txt); PA4's 8 KB sets would not fit twice in the 14 blocks per SM of the
    // one-wave 65 536-env batch anyway.
    static constexpr int kSetBytes = round_up(kOthOff + kOthBytes, 1024);
    static constexpr int kOverhead = ActionStream<N>::kSmemBytes + 1024 /* alignment slack */;
    static constexpr int kStages = (N <= CM3_PT_STAGE_MAXN && 14 * (2 * kSetBytes + kOverhead + 1024 /* driver-reserved */) <= 227 * 1024) ? 2 : 1;
    static constexpr int kActOff = kStages * kSetBytes;             // ActionStream slots
    static constexpr int kSmemBytes = kActOff + kOverhead;
    // a TMA box has at most 256 rows: the wide records of N >= 6 (and N = 5..8 in double) do not fit
    // one box per tile and take the linear tile + plain bulk stores instead
    static constexpr bool kTmaOk = kRowRows <= 256 && kOthRows <= 256;
};


// A [T][out_B] output field seen as a 2-D byte tensor of `width`-byte rows; one box = one tile of
// consecutive envs (box_rows rows).  Defined in particle.cu.
bool encode_tile_map(CUtensorMap *tm, void *base, size_t total_bytes, int sw_bits, int width, int box_rows);

}  // namespace cm3
