// checkers.cu - fused Checkers reset/step/rollout kernel for sm_100a.
//
// What it computes: env/checkers.py of the reference - agent_act (:157-187), get_reward
// (:190-225), step (:228-262), reset + populate_world (:265-291, :38-63), get_global_state
// (:79-94) and get_local_observation (:128-154) - for B env instances at once.
//
// Design:
//  * state is compact: a bitboard of the cells that still hold a reward, packed agent words
//    and a step/goal word - 8 + 4N + 4 bytes per env; the dense [rows][cols][3] float world of
//    the reference is never materialised;
//  * one lane per (env, agent): a warp owns EW = 32/N consecutive envs.  Every lane replays the
//    (sequential, order-dependent) N-agent move/collect update of its env redundantly in
//    registers - it is ~40 instructions per agent - so no lane waits for another;
//  * the two bulky outputs (the per-agent (2O+1)^2 x 3 window and the R x (C+1) x 2 grid, 90 % of
//    the bytes) are expanded from bit patterns into a per-warp shared-memory tile whose layout is
//    exactly the layout of the output arrays for those EW envs, and leave the SM as ONE TMA bulk
//    store per field (cp.async.bulk.global.shared::cta) - full-line HBM writes with no store
//    instructions spent on them.  The lane->(env,agent) mapping is agent-major so that for N = 2
//    the two half-warps hit odd/even banks: the staging stores are conflict free;
//  * the small per-agent vectors are written straight from registers: consecutive lanes write
//    consecutive 16-byte records;
//  * T steps can be fused in one launch (rollout): state stays in registers, actions come from
//    an int8 stream or from Philox4x32-10, and finished episodes can be reset in-kernel.
#include "common.cuh"
#include "params.cuh"

#ifndef CM3_CK_REAL
#error "compile with -DCM3_CK_REAL=0 (float), =1 (double), =2 (float + int8 tiles) and =3 (float + 2-bit packed tiles)"
#endif

namespace cm3 {

// One warp per block: with 13 KB of staging per warp 16 blocks fit an SM (ncu: shared-memory limited),
// i.e. 2368 of the 4096 CK2 tiles of the 65 536-env headline batch are resident at once, 1.73 waves
// (measured: +7 % over 2 warps per block, profiles/r01b_ab.txt).
#ifndef CM3_CK_WPB
#define CM3_CK_WPB 1
#endif
constexpr int kCkWarpsPerBlock = CM3_CK_WPB;

// Cells hold a value in {0, +1, -1}, described by two bit masks over the cells of a row: nz (cell
// is non-zero) and neg (cell is -1; neg implies nz).  tri_pack() puts the two masks side by side,
// w = nz | neg << 8 (at most 8 cells at a time), and tri_cell<T>(w, j) expands cell j:
//   float : (w & (0x101 << j)) * (0x3F800000 >> j) - the nz bit lands on 0x3F800000 (1.0f), the neg
//           bit on 0x7F << 31 = 0x80000000 (mod 2^32, 0x7F is odd), their sum is 0xBF800000 (-1.0f):
//           one LOP3 and one IMAD per output word, no conversions;
//   int8  : the same numbers as signed bytes (compact tiles: lossless, 4x fewer bytes);
//   double: a conversion (parity mode only).
__device__ __forceinline__ uint32_t tri_pack(uint32_t nz, uint32_t neg) { return (nz & 0xFFu) | ((neg & 0xFFu) << 8); }
template <typename Tile> __device__ __forceinline__ Tile tri_cell(uint32_t w, int j);
template <> __device__ __forceinline__ float tri_cell<float>(uint32_t w, int j) {
    return __uint_as_float((w & (0x101u << j)) * (0x3F800000u >> j));
}
template <> __device__ __forceinline__ double tri_cell<double>(uint32_t w, int j) {
    return (double)((int)((w >> j) & 1u) - 2 * (int)((w >> (j + 8)) & 1u));
}
template <> __device__ __forceinline__ int8_t tri_cell<int8_t>(uint32_t w, int j) {
    return (int8_t)(((w >> j) & 1u) | (((w >> (j + 8)) & 1u) * 0xFEu));
}

// CM3_TILE_U2: the two bulky outputs as 2 bits per cell (0 -> 0, 1 -> +1, 3 -> -1: two's complement),
// least significant cell first, packed into 32-bit words:
//   obs_self_t [B][N][W][RW]   one window ROW (W cells x 3 channels, cell (dc, ch) at bits 2 (3 dc + ch))
//                              per RW = ceil(6 W / 32) words: one word for n_obs <= 2, two for n_obs = 3
//   grid       [B][R][GW]      one grid ROW (C + 1 cells x 2 channels, cell (j, ch) at bits 2 (2 j + ch)),
//                              8 cells per word, GW = ceil((C + 1) / 8) words
// The same numbers in 1/16 of the float bytes (CK2: 816 -> 64 bytes of tiles per env-step).
struct TileU2 { uint32_t w; };
template <typename Tile> struct is_u2 { static constexpr bool value = false; };
template <> struct is_u2<TileU2> { static constexpr bool value = true; };
__host__ __device__ constexpr int u2_row_words(int W) { return (6 * W + 31) / 32; }
__host__ __device__ constexpr int u2_grid_words(int C) { return (C + 1 + 7) / 8; }
// bit i of x (i < 5) -> bit 6 i: the 25 partial products of the multiplication land on distinct bit
// positions (i + 5 k = i' + 5 k' with |i - i'| <= 4 forces equality), so there are no carries
__device__ __forceinline__ uint32_t u2_spread6(uint32_t x5) { return (x5 * 0x00108421u) & 0x01041041u; }
// bit i of x (i < 8) -> bit 4 i
__device__ __forceinline__ uint32_t u2_spread4(uint32_t x) {
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}

template <typename Real> __device__ __forceinline__ void store4(Real *p, Real a, Real b, Real c, Real d);
template <> __device__ __forceinline__ void store4<float>(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<double>(double *p, double a, double b, double c, double d) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2 *>(p)[1] = make_double2(c, d);
}
template <typename Real> __device__ __forceinline__ void store2(Real *p, Real a, Real b);
template <> __device__ __forceinline__ void store2<float>(float *p, float a, float b) {
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void store2<double>(double *p, double a, double b) {
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
}

template <int N> __device__ __forceinline__ int pick(const int (&v)[N], int idx) {
    int r = v[0];
#pragma unroll
    for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
    return r;
}

// ---------------------------------------------------------------- board geometry
// The kernel is written against a geometry provider.  StaticGeo compiles (n_rows, n_columns, n_obs)
// in - every loop unrolls, every mask is an immediate - and exists for the boards of the reference's
// configs; DynGeo reads them from the parameter block, so that ANY board the bitboards can address
// runs (Checkers.__init__ takes the geometry as data, env/checkers.py:5-35).
template <int R_, int C_, int O_>
struct StaticGeo {
    static constexpr bool kStatic = true;
    __host__ __device__ StaticGeo() {}
    __device__ __forceinline__ explicit StaticGeo(const CkParams &) {}
    __host__ __device__ static constexpr int R() { return R_; }
    __host__ __device__ static constexpr int C() { return C_; }
    __host__ __device__ static constexpr int O() { return O_; }
    // cells (i,j) with (i+j) even are green, odd orange (checkers.py:54-63)
    __host__ __device__ static constexpr uint64_t color_const(int color) {
        uint64_t m = 0;
        for (int i = 0; i < R_; ++i)
            for (int j = 0; j < C_; ++j)
                if (((i + j) & 1) == color) m |= 1ull << (i * C_ + j);
        return m;
    }
    __device__ __forceinline__ uint64_t color_all(const CkParams &, int color) const {
        return color ? color_const(1) : color_const(0);
    }
    static_assert(R_ % 2 == 1 && C_ % 2 == 0, "checkers.py:16-17");
    static_assert(R_ * C_ <= 64 && C_ + 2 * O_ + 1 <= 32 && R_ + 2 * O_ <= kCkMaxTR && 2 * O_ + 1 <= 8, "geometry");
};
struct DynGeo {
    static constexpr bool kStatic = false;
    int r, c, o;
    __device__ __forceinline__ explicit DynGeo(const CkParams &p) : r(p.R), c(p.C), o(p.O) {}
    __host__ DynGeo(int r_, int c_, int o_) : r(r_), c(c_), o(o_) {}
    __host__ __device__ __forceinline__ int R() const { return r; }
    __host__ __device__ __forceinline__ int C() const { return c; }
    __host__ __device__ __forceinline__ int O() const { return o; }
    __device__ __forceinline__ uint64_t color_all(const CkParams &p, int color) const { return p.color_mask[color]; }
};

// lanes reserved per env: one per agent, rounded up to a power of two - and TWO for a single agent
// (the second lane expands the global grid while the first expands the window: twice the warps for
// the stage-1 board, whose one-thread-per-env launch was latency-bound in round 1)
__host__ __device__ constexpr int ck_lanes_per_env(int N) { return N <= 2 ? 2 : N <= 4 ? 4 : 8; }

// Real: type of the small vectors and rewards; Tile: element type of the two bulky outputs (the
// per-agent window and the global grid) - Real, or int8_t for the compact encoding.
template <int N, typename Real, typename Tile>
struct CkLayout {
    static constexpr int L = 2 * (N > 1 ? N - 1 : 1);
    static constexpr int NP = ck_lanes_per_env(N);
    static constexpr int EW = kWarp / NP;  // envs per warp
    // elements of Tile per agent window / per env grid
    template <class Geo> __host__ __device__ static int win_elems(const Geo &g) {
        const int W = 2 * g.O() + 1;
        return is_u2<Tile>::value ? W * u2_row_words(W) : W * W * 3;
    }
    template <class Geo> __host__ __device__ static int grid_elems(const Geo &g) {
        return is_u2<Tile>::value ? g.R() * u2_grid_words(g.C()) : g.R() * (g.C() + 1) * 2;
    }
    template <class Geo> __host__ __device__ static int win_bytes(const Geo &g) {
        return round_up(EW * N * win_elems(g) * (int)sizeof(Tile), 16);
    }
    template <class Geo> __host__ __device__ static int grid_bytes(const Geo &g) {
        return round_up(EW * grid_elems(g) * (int)sizeof(Tile), 16);
    }
    template <class Geo> __host__ __device__ static int lut_bytes(const Geo &g) {
        const int TR = g.R() + 2 * g.O(), TC = g.C() + 2 * g.O() + 1, CNT = g.R() * g.C() / 2 + 1;
        return round_up((TR + TC + CNT) * (int)sizeof(Real), 128);
    }
    template <class Geo> __host__ __device__ static int warp_stage_bytes(const Geo &g) {
        return win_bytes(g) + grid_bytes(g) + ActionStream<N>::kSmemBytes;
    }
    template <class Geo> __host__ __device__ static int smem_bytes(const Geo &g) {
        return lut_bytes(g) + kCkWarpsPerBlock * warp_stage_bytes(g);
    }
    static_assert(N >= 1 && N <= CM3_MAX_AGENTS, "agents");
};

constexpr uint32_t kTagGoal = 0x60A10000u;  // Philox stream tag of the stage-1 goal redraw

template <class Geo, int N, typename Real, typename Tile>
__global__ void __launch_bounds__(kCkWarpsPerBlock *kWarp)
checkers_kernel(const __grid_constant__ CkParams p) {
    using Ly = CkLayout<N, Real, Tile>;
    const Geo geo(p);
    const int R = geo.R(), C = geo.C(), O = geo.O();
    const int TR = R + 2 * O, TC = C + 2 * O + 1, W = 2 * O + 1;
    const int WW3 = Ly::win_elems(geo), G = Ly::grid_elems(geo);  // Tile elements per agent window / per env grid
    const int CNT = R * C / 2 + 1;
    constexpr int L = Ly::L, EW = Ly::EW;
    const uint32_t CM = (C >= 32) ? 0xFFFFFFFFu : ((1u << C) - 1u);
    const uint32_t kEven = 0x55555555u & CM, kOdd = 0xAAAAAAAAu & CM;  // columns j with j even / odd
    const uint64_t kFull = (R * C >= 64) ? ~0ull : ((1ull << (R * C)) - 1ull);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool leader = elect_one();  // issues, commits and waits for this warp's bulk stores (common.cuh)
    // A block normally owns one tile per warp.  A chained single-step launch with more tiles than resident slots
    // gives every block p.tpb (2..4) tiles instead, one grid stride apart, which it steps one after the other: the
    // launch is then a single resident wave (or as close to one as four tiles per block get it), the early release
    // applies, and fewer blocks queue for a slot (launch_ck).
    const int tile_first = p.tile0 + blockIdx.x * kCkWarpsPerBlock + warp;
    const int tile_stride = (int)gridDim.x * kCkWarpsPerBlock;
    const int n_my_tiles = p.tpb > 1 ? p.tpb : 1;
    // launch chaining: the tickets of ALL my tiles are taken BEFORE the next grid may be scheduled (common.cuh)
    // (0xFFFFFFFF = not taken: neutral in the AND below, which must depend on every atomic that WAS issued)
    uint32_t mine0 = 0xFFFFFFFFu, mine1 = 0xFFFFFFFFu, mine2 = 0xFFFFFFFFu, mine3 = 0xFFFFFFFFu;
    if (p.sync != nullptr) {
        // all atomics in flight together (lane 0), then one broadcast each
        const int t1 = tile_first + tile_stride, t2 = t1 + tile_stride, t3 = t2 + tile_stride;
        const bool h0 = tile_first * EW < p.B, h1 = n_my_tiles > 1 && t1 * EW < p.B, h2 = n_my_tiles > 2 && t2 * EW < p.B,
                   h3 = n_my_tiles > 3 && t3 * EW < p.B;
        if (lane == 0) {
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
    // the trigger is issued under a branch on the returned tickets: every atomic has been performed at L2 before
    // any block of the next grid can take its own (a ticket never reaches 0xFFFFFFFF: the chain traps long before)
    if ((mine0 & mine1 & mine2 & mine3) != 0xFFFFFFFFu) pdl_launch_dependents();  // the next step's grid may become resident while this one drains

    extern __shared__ __align__(128) unsigned char smem_raw[];
    Real *lut_row = reinterpret_cast<Real *>(smem_raw);
    Real *lut_col = lut_row + TR;
    Real *lut_cnt = lut_col + TC;
    for (int i = threadIdx.x; i < TR; i += blockDim.x) lut_row[i] = (Real)p.norm_row[i];
    for (int i = threadIdx.x; i < TC; i += blockDim.x) lut_col[i] = (Real)p.norm_col[i];
    for (int i = threadIdx.x; i < CNT; i += blockDim.x) lut_cnt[i] = (Real)p.norm_cnt[i];
    __syncthreads();

    const int win_bytes = Ly::win_bytes(geo), grid_bytes = Ly::grid_bytes(geo);
    Tile *stage_win = reinterpret_cast<Tile *>(smem_raw + Ly::lut_bytes(geo) + warp * Ly::warp_stage_bytes(geo));
    Tile *stage_grid = reinterpret_cast<Tile *>(reinterpret_cast<unsigned char *>(stage_win) + win_bytes);

#pragma unroll 1
    for (int my = 0; my < n_my_tiles; ++my) {
    const int tile = tile_first + my * tile_stride;
    const int env0 = tile * EW;
    if (env0 >= p.B) return;  // warp-uniform; no block-level sync below this point (later tiles of mine lie further out)
    TileTicket ticket;
    ticket.on = p.sync != nullptr;
    ticket.w = p.sync + 2 * (size_t)tile;
    ticket.mine = my == 0 ? mine0 : my == 1 ? mine1 : my == 2 ? mine2 : mine3;
    const int a = lane / EW, e = lane % EW;  // agent-major: lanes [a*EW, (a+1)*EW) hold agent a
    const int env = env0 + e;
    const int nenv = min(EW, p.B - env0);
    const bool live = e < nenv;            // this lane works for an env of the batch ...
    const bool owner = live && (a < N);    // ... and owns agent a of it
    const size_t B = (size_t)p.B;

    // the state this tile's previous launch wrote is visible from here on
    if (p.chained) ticket.wait(lane); else pdl_wait();

    // ---- load compact state
    uint64_t rem = 0;
    int ar[N], ac[N], ng[N], no[N];
    uint32_t meta = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) { ar[i] = O; ac[i] = O + i; ng[i] = 0; no[i] = 0; }
    if (live) {
        rem = __ldcg(reinterpret_cast<const unsigned long long *>(p.remaining) + env);
        meta = __ldcg(p.meta + env);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t w = __ldcg(p.agents + (size_t)env * N + i);
            ar[i] = w & 0xFF; ac[i] = (w >> 8) & 0xFF; ng[i] = (w >> 16) & 0xFF; no[i] = w >> 24;
        }
    }
    int steps = meta & 0xFFFFFF;
    uint32_t goals = meta >> 24;

    auto reset_state = [&]() {
        rem = kFull;
        steps = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ar[i] = p.start_r[i]; ac[i] = p.start_c[i]; ng[i] = 0; no[i] = 0;
        }
        if (N == 1) ar[0] = ((goals & 1) ? 2 : 0) + O;  // checkers.py:271-276
    };

    bool pending = false;  // bulk stores of this warp whose smem source may still be in flight

    // action rows: streamed through shared memory (multi-step launches on whole tiles, see
    // ActionStream), else loaded directly at the top of each step
    ActionStream<N> acts;
    acts.init(reinterpret_cast<unsigned char *>(stage_grid) + grid_bytes, p.mode == kCkReset ? nullptr : p.actions, p.B,
              env0, EW, nenv == EW, p.T, lane);
    uint32_t act_word = acts.on ? acts.begin(e) : 0u;

    // Expands the observations of the current state into the outputs of time slot t.
    const CkOut &o0 = p.out[0];
    const size_t OB = (size_t)p.out_B, oe0 = (size_t)p.out_env0;
    auto emit = [&](int t) {
        const size_t slot = (size_t)t * OB + oe0;
        const int my_r = pick<N>(ar, a), my_c = pick<N>(ac, a);
        // next step's action word: in flight from here to the end of this phase (ActionStream)
        const uint32_t act_loaded = acts.on ? acts.load(t + 1) : 0u;
        if (acts.on) acts.prefetch(t + 3);
        if (pending) {
            if (leader) bulk_wait_read();
        }
        __syncwarp();
        // ---------------- window of agent a (get_obs, checkers.py:97-109)
        if (o0.obs_self_t != nullptr && owner) {
            Tile *win = stage_win + (e * N + a) * WW3;
            const int sh = my_c - O;  // leftmost window column, >= 0
#pragma unroll
            for (int dr = 0; dr < W; ++dr) {
                const int pr = my_r - O + dr;  // expanded row
                const int i = pr - O;          // valid-grid row
                const bool inr = (unsigned)i < (unsigned)R;
                const uint32_t rowrem = inr ? ((uint32_t)(rem >> (inr ? i * C : 0)) & CM) : 0u;
                const uint32_t gmask = (i & 1) ? kOdd : kEven;
                const uint32_t omask = CM & ~gmask;
                uint32_t occ = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) occ |= (ar[k] == pr) ? (1u << ac[k]) : 0u;
                if (dr == O) occ &= ~(1u << my_c);  // own cell reads as free, :107
                // patterns over expanded columns, then cut to the W window columns
                const uint32_t border = inr ? (((1u << O) - 1u) | (~0u << (O + C + 1))) : ~0u;
                const uint32_t nz0 = (inr ? (gmask << O) : 0u) >> sh;
                const uint32_t ng0 = ((rowrem & gmask) << O) >> sh;
                const uint32_t nz1 = (inr ? (omask << O) : 0u) >> sh;
                const uint32_t ng1 = ((rowrem & omask) << O) >> sh;
                const uint32_t nz2 = (border | occ) >> sh;
                const uint32_t ng2 = occ >> sh;
                if constexpr (is_u2<Tile>::value) {
                    // one packed word (two for n_obs = 3) per window row
                    const uint32_t wm = (1u << W) - 1u;
                    uint32_t *row = reinterpret_cast<uint32_t *>(win) + dr * u2_row_words(W);
                    if (W <= 5) {
                        row[0] = u2_spread6(nz0 & wm) | (u2_spread6(ng0 & wm) << 1) | (u2_spread6(nz1 & wm) << 2) |
                                 (u2_spread6(ng1 & wm) << 3) | (u2_spread6(nz2 & wm) << 4) | (u2_spread6(ng2 & wm) << 5);
                    } else {
                        unsigned long long acc = 0ull;
                        for (int dc = 0; dc < W; ++dc) {
                            const unsigned long long code = ((nz0 >> dc) & 1u) | (((ng0 >> dc) & 1u) << 1) | (((nz1 >> dc) & 1u) << 2) |
                                                            (((ng1 >> dc) & 1u) << 3) | (((nz2 >> dc) & 1u) << 4) | (((ng2 >> dc) & 1u) << 5);
                            acc |= code << (6 * dc);
                        }
                        row[0] = (uint32_t)acc; row[1] = (uint32_t)(acc >> 32);
                    }
                } else {
                    const uint32_t w0 = tri_pack(nz0, ng0), w1 = tri_pack(nz1, ng1), w2 = tri_pack(nz2, ng2);
#pragma unroll
                    for (int dc = 0; dc < W; ++dc) {
                        Tile *cell = win + (dr * W + dc) * 3;
                        cell[0] = tri_cell<Tile>(w0, dc);
                        cell[1] = tri_cell<Tile>(w1, dc);
                        cell[2] = tri_cell<Tile>(w2, dc);
                    }
                }
            }
        }
        // the window tile (3/4 of the bytes) leaves first and drains while the rest is expanded
        pending = false;
        if (o0.obs_self_t != nullptr) {
            fence_proxy_async();
            __syncwarp();
            const uint32_t bytes = (uint32_t)(nenv * N * WW3 * sizeof(Tile));
            for (int d = 0; d < p.n_dst; ++d) {
                Tile *g = reinterpret_cast<Tile *>(p.out[d].obs_self_t) + (slot + env0) * (size_t)(N * WW3);
                if (nenv == EW && ((reinterpret_cast<uintptr_t>(g) | bytes) & 15u) == 0) {
#ifdef CM3_L2_HINT_BULK
                    if (leader) bulk_store_hint(g, stage_win, bytes, l2_policy_evict_first());
#else
                    if (leader) bulk_store(g, stage_win, bytes);
#endif
                    pending = true;
                } else {
                    for (int idx = lane; idx < nenv * N * WW3; idx += kWarp) g[idx] = stage_win[idx];
                }
            }
            if (leader) bulk_commit();
        }
        // ---------------- global grid (get_valid_grid, :66-76).  N > 1: lane a < 2 writes channel a;
        // N == 1: the env's second lane writes both channels
        if constexpr (is_u2<Tile>::value) {
            // packed grid rows: N > 1 - lane a < 2 takes the rows i with (i & 1) == a; N == 1 - the second lane takes all
            if (o0.grid != nullptr && live && (N == 1 ? a == 1 : a < 2)) {
                uint32_t *gr = reinterpret_cast<uint32_t *>(stage_grid) + e * G;
                const int GW = u2_grid_words(C);
                for (int i = (N == 1 ? 0 : a); i < R; i += (N == 1 ? 1 : 2)) {
                    const uint32_t rowrem = (uint32_t)(rem >> (i * C)) & CM;
                    const uint32_t c0 = (i & 1) ? kOdd : kEven, c1 = (i & 1) ? kEven : kOdd;  // channel 0 green, 1 orange
                    for (int gw = 0; gw < GW; ++gw) {
                        const int j0 = 8 * gw;
                        gr[i * GW + gw] = u2_spread4((c0 >> j0) & 0xFFu) | (u2_spread4(((rowrem & c0) >> j0) & 0xFFu) << 1) |
                                          (u2_spread4((c1 >> j0) & 0xFFu) << 2) | (u2_spread4(((rowrem & c1) >> j0) & 0xFFu) << 3);
                    }
                }
            }
        } else
        if (o0.grid != nullptr && live && (N == 1 ? a == 1 : a < 2)) {
            Tile *gr = stage_grid + e * G;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const uint32_t rowrem = (uint32_t)(rem >> (i * C)) & CM;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    if (N > 1 && ch == 1) break;           // N > 1: one channel per lane
                    const int mych = (N > 1) ? a : ch;
                    const uint32_t cmask = (((i & 1) ^ mych) ? kOdd : kEven);
                    const uint32_t neg = rowrem & cmask;
                    // columns in groups of 8 (tri_pack); column C (the start column) holds no reward
#pragma unroll
                    for (int j0 = 0; j0 <= C; j0 += 8) {
                        const uint32_t w = tri_pack(cmask >> j0, neg >> j0);
#pragma unroll
                        for (int j = j0; j < j0 + 8 && j <= C; ++j)
                            gr[(i * (C + 1) + j) * 2 + mych] = tri_cell<Tile>(w, j - j0);
                    }
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (o0.grid != nullptr) {
            const uint32_t bytes = (uint32_t)(nenv * G * sizeof(Tile));
            for (int d = 0; d < p.n_dst; ++d) {
                Tile *g = reinterpret_cast<Tile *>(p.out[d].grid) + (slot + env0) * (size_t)G;
                if (nenv == EW && ((reinterpret_cast<uintptr_t>(g) | bytes) & 15u) == 0) {
#ifdef CM3_L2_HINT_BULK
                    if (leader) bulk_store_hint(g, stage_grid, bytes, l2_policy_evict_first());
#else
                    if (leader) bulk_store(g, stage_grid, bytes);
#endif
                    pending = true;
                } else {
                    for (int idx = lane; idx < nenv * G; idx += kWarp) g[idx] = stage_grid[idx];
                }
            }
            if (leader) bulk_commit();
        }
        if (acts.on) act_word = acts.hand_over(t, act_loaded, e);
        // ---------------- small per-agent vectors, straight from registers
        if (owner) {
            const size_t rec = (slot + env) * N + a;
            const int my_g = pick<N>(ng, a), my_o = pick<N>(no, a);
            for (int d = 0; d < p.n_dst; ++d) {
                const CkOut &o = p.out[d];
                if (o.vec != nullptr)  // get_global_state, :89-93
                    store4<Real>(reinterpret_cast<Real *>(o.vec) + rec * 4, (Real)my_r, (Real)my_c, (Real)my_g, (Real)my_o);
                if (o.obs_self_v != nullptr)  // :137-139
                    store4<Real>(reinterpret_cast<Real *>(o.obs_self_v) + rec * 4, lut_row[my_r], lut_col[my_c],
                                 lut_cnt[my_g], lut_cnt[my_o]);
                if (o.obs_others != nullptr) {  // :143-151
                    Real *oo = reinterpret_cast<Real *>(o.obs_others) + rec * L;
                    if (N == 1) {
                        store2<Real>(oo, lut_row[my_r], lut_col[my_c]);
                    } else {
#pragma unroll
                        for (int k = 0; k < N - 1; ++k) {
                            const int j = k + (k >= a ? 1 : 0);
                            store2<Real>(oo + 2 * k, lut_row[pick<N>(ar, j)], lut_col[pick<N>(ac, j)]);
                        }
                    }
                }
                if (o.goal_idx != nullptr) o.goal_idx[rec] = (uint8_t)((goals >> a) & 1u);  // :235
            }
        }
    };

    // ---- store compact state
    auto store_state = [&]() {
        if (live && a == 0) {
            p.remaining[env] = rem;
            p.meta[env] = (uint32_t)steps | (goals << 24);
#pragma unroll
            for (int i = 0; i < N; ++i)
                p.agents[(size_t)env * N + i] =
                    (uint32_t)ar[i] | ((uint32_t)ac[i] << 8) | ((uint32_t)ng[i] << 16) | ((uint32_t)no[i] << 24);
        }
    };

    const int T_eff = (p.mode == kCkReset) ? 1 : p.T;
    for (int t = 0; t < T_eff; ++t) {
        bool sel = false;
        if (p.mode == kCkReset) {
            sel = live && (p.env_mask == nullptr || p.env_mask[env] != 0);
            if (sel) {
                if (p.goal_idx != nullptr) {  // NULL: the env keeps the goals it has
                    goals = 0;
#pragma unroll
                    for (int i = 0; i < N; ++i) goals |= (uint32_t)(p.goal_idx[(size_t)env * N + i] & 1u) << i;
                }
                reset_state();
            }
        } else {
            // ---- actions of all agents of my env
            int act[N];
            if (p.actions != nullptr) {
                if constexpr (N <= 4) {
                    uint32_t w = act_word;  // picked up from the stream during the previous emit
                    if (!acts.on && live) w = load_actions_packed<N>(p.actions + ((size_t)t * B + env) * N);
#pragma unroll
                    for (int i = 0; i < N; ++i) act[i] = unpack_action(w, i);
                } else {  // more than four agents do not fit the packed word: plain byte loads
#pragma unroll
                    for (int i = 0; i < N; ++i) act[i] = live ? (int)p.actions[((size_t)t * B + env) * N + i] : 0;
                }
            } else {
                // one Philox block of 4 words per 4 agents
#pragma unroll
                for (int i0 = 0; i0 < N; i0 += 4) {
                    const Philox4 w = philox_action_words(p.seed, (uint64_t)(p.env_id_offset + env),
                                                          (uint64_t)(p.t0 + t), i0 / 4);
#pragma unroll
                    for (int i = i0; i < N && i < i0 + 4; ++i) act[i] = action_from_word(philox_word(w, i - i0), 5);
                }
            }
            if (p.actions_out != nullptr && owner)
                p.actions_out[((size_t)t * B + env) * N + a] = (int8_t)pick<N>(act, a);

            // ---- agents act and collect strictly in index order (checkers.py:233-237)
            double rew[N];
            double total = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int ai = act[i];
                const int tr = ar[i] + (ai == 2) - (ai == 1);
                const int tc = ac[i] + (ai == 4) - (ai == 3);
                const bool wants = (unsigned)(ai - 1) < 4u;
                // world[..,2] == 0 <=> inside the valid grid and no agent there (:43-51)
                bool free_cell = (unsigned)(tr - O) < (unsigned)R && (unsigned)(tc - O) <= (unsigned)C;
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (k != i) free_cell = free_cell && !(ar[k] == tr && ac[k] == tc);
                const bool moved = wants && free_cell;
                const bool penal = (ai != 0) && !moved;  // :184-186, any other action included
                if (moved) { ar[i] = tr; ac[i] = tc; }
                double col = 0.0;  // get_reward, :190-225
                const int ii = ar[i] - O, jj = ac[i] - O;
                if (jj < C) {
                    const int bit = ii * C + jj;
                    if ((rem >> bit) & 1ull) {
                        rem &= ~(1ull << bit);
                        const int color = (ii + jj) & 1;
                        if (color == 0) ng[i] += 1; else no[i] += 1;
                        col = (color == (int)((goals >> i) & 1u)) ? 1.0 : -0.5;
                    }
                }
                rew[i] = (penal ? -0.1 : 0.0) + col;  // penalty + get_reward, :236
                total += rew[i];                      // np.sum, :243
            }
            steps = (steps + 1) & 0xFFFFFF;  // :244
            bool done;                       // :246-260
            if (N == 1) {
                const uint64_t mine = geo.color_all(p, (int)(goals & 1));
                done = (steps == p.max_steps) || ((rem & mine) == 0ull);
            } else {
                done = (steps == p.max_steps) || (rem == 0ull);
            }
            if (owner) {
                double mine = rew[0];
#pragma unroll
                for (int i = 1; i < N; ++i) mine = (a == i) ? rew[i] : mine;
                const size_t orow = (size_t)t * OB + oe0 + env;
                for (int d = 0; d < p.n_dst; ++d) {
                    const CkOut &o = p.out[d];
                    if (o.local_rewards != nullptr) reinterpret_cast<Real *>(o.local_rewards)[orow * N + a] = (Real)mine;
                    if (a == 0) {
                        if (o.reward != nullptr) reinterpret_cast<Real *>(o.reward)[orow] = (Real)total;
                        if (o.done != nullptr) o.done[orow] = done ? 1 : 0;
                    }
                }
            }
            if (p.auto_reset && done) {
                if (N == 1 && p.random_goal) {  // a fresh goal per episode, train_offpolicy.py:291-296
                    const unsigned long long genv = (unsigned long long)(p.env_id_offset + env), ctr = (unsigned long long)(p.t0 + t + 1);
                    const Philox4 w = philox4x32_10((uint32_t)genv, (uint32_t)(genv >> 32), (uint32_t)ctr,
                                                    kTagGoal | ((uint32_t)(ctr >> 32) & 0xFFFFu), (uint32_t)p.seed,
                                                    (uint32_t)(p.seed >> 32));
                    goals = w.x >> 31;
                }
                reset_state();
            }
        }
        // The state is final here: the next launch of this tile needs IT, not the observations.  p.early (chained
        // launches, params.cuh: chain_early_mode) writes it back - and, at 2, releases the tile - before the
        // tiles of the last step are expanded and stored.
        if (p.early != 0 && t == T_eff - 1) {
            // Every lane of an env loaded the env's state for itself, and lane a == 0 is about to overwrite it: all
            // those loads must have been performed first.  Nothing else orders them - the lanes of a warp need not
            // run in step behind the spin-wait (a __syncwarp is a barrier, not a promise of lockstep), and with the
            // store this early the one lane that polled was seen to finish its step and store before the others
            // had loaded (outputs of the tile's first env one action ahead, profiles/r02u_ab.txt (5)).
            __syncwarp();
            store_state();
            if (p.early == 2) ticket.publish(lane);
        }
        emit(t);
        if (sel && a == 0 && o0.done != nullptr) o0.done[oe0 + env] = 0;  // checkers.py:291
    }

    if (p.early == 0 || T_eff < 1) store_state();
    if (p.early != 2 || T_eff < 1) ticket.publish(lane);  // this tile's next launch may go ahead
    // smem must outlive the async reads (and be free for my next tile); the global writes themselves complete with the grid
    if (pending && leader) bulk_wait_read();
    __syncwarp();
    }  // my tiles
}

// ------------------------------------------------------------------------ host side

constexpr int kCkDynSmemMax = 160 * 1024;  // ceiling requested once for the geometry-as-data kernels

template <class Geo, int N, typename Real, typename Tile>
static int launch_ck(const Geo &geo, const CkParams &p, cudaStream_t stream) {
    using Ly = CkLayout<N, Real, Tile>;
    auto kern = checkers_kernel<Geo, N, Real, Tile>;
    static std::atomic<uint64_t> attr_done{0};
    const int smem = Ly::smem_bytes(geo);
    if (smem > kCkDynSmemMax) {
        set_error("board needs %d bytes of staging per block (limit %d)", smem, kCkDynSmemMax);
        return CM3_ERR_UNSUPPORTED;
    }
    CM3_CUDA(ensure_smem_attr(kern, attr_done));
    const int ntiles = (p.B + Ly::EW - 1) / Ly::EW;
    const int nblocks = (ntiles + kCkWarpsPerBlock - 1) / kCkWarpsPerBlock;
    // multi-step launches: equal waves (common.cuh: balance_waves)
    const int smem_launch = (p.mode == kCkStep && p.T > 1) ? balance_waves((const void *)kern, kCkWarpsPerBlock * kWarp, smem, nblocks) : smem;
    const int parts = (p.chained && kCkWarpsPerBlock == 1) ? chain_parts(ntiles) : 1;
    // chained single-step launches: as many tiles per block as it takes to make the launch one resident wave
    // (at most 4; params.cuh: chain_tiles_per_block), then the early release applies
    const int tpb = (p.chained && p.T == 1 && parts == 1 && kCkWarpsPerBlock == 1)
                        ? chain_tiles_per_block((const void *)kern, kWarp, smem_launch, nblocks) : 1;
    const int grid_blocks = (nblocks + tpb - 1) / tpb;
    const int early = (p.chained && p.early < 0) ? chain_early_mode((const void *)kern, kCkWarpsPerBlock * kWarp, smem_launch, grid_blocks)
                                                 : (p.early < 0 ? 0 : p.early);
    for (int i = 0; i < parts; ++i) {  // disjoint tile ranges; one grid unless chained (params.cuh: chain_parts)
        CkParams q = p;
        q.early = early;
        q.tpb = tpb;
        q.tile0 = (int)((long long)grid_blocks * i / parts);
        const int n = (int)((long long)grid_blocks * (i + 1) / parts) - q.tile0;
        if (n > 0) CM3_CUDA(launch_kernel(kern, n, kCkWarpsPerBlock * kWarp, smem_launch, stream, pdl_enabled(), q));
    }
    return CM3_OK;
}

// boards with the geometry compiled in: the reference's configs (alg/config_checkers_stage{1,2}.json:
// 3 x 8, n_obs 2) and the constructor's default board (env/checkers.py:5: 3 x 16, n_obs 2)
#define CM3_CK_STATIC_GEOMS(X) \
    X(3, 8, 2)                 \
    X(3, 16, 2)

template <typename Real, typename Tile>
static int dispatch_ck(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream) {
    if (N <= 4 && !dyn_geometry_forced()) {
#define X(r, c, o)                                                                             \
    if (R == r && C == c && O == o) {                                                          \
        using Geo = StaticGeo<r, c, o>;                                                        \
        const Geo g{};                                                                         \
        switch (N) {                                                                           \
            case 1: return launch_ck<Geo, 1, Real, Tile>(g, p, stream);                        \
            case 2: return launch_ck<Geo, 2, Real, Tile>(g, p, stream);                        \
            case 3: return launch_ck<Geo, 3, Real, Tile>(g, p, stream);                        \
            case 4: return launch_ck<Geo, 4, Real, Tile>(g, p, stream);                        \
            default: break;                                                                    \
        }                                                                                      \
    }
        CM3_CK_STATIC_GEOMS(X)
#undef X
    }
    const DynGeo g(R, C, O);
    switch (N) {
        case 1: return launch_ck<DynGeo, 1, Real, Tile>(g, p, stream);
        case 2: return launch_ck<DynGeo, 2, Real, Tile>(g, p, stream);
        case 3: return launch_ck<DynGeo, 3, Real, Tile>(g, p, stream);
        case 4: return launch_ck<DynGeo, 4, Real, Tile>(g, p, stream);
        case 5: return launch_ck<DynGeo, 5, Real, Tile>(g, p, stream);
        case 6: return launch_ck<DynGeo, 6, Real, Tile>(g, p, stream);
        case 7: return launch_ck<DynGeo, 7, Real, Tile>(g, p, stream);
        case 8: return launch_ck<DynGeo, 8, Real, Tile>(g, p, stream);
        default: break;
    }
    set_error("n_agents=%d outside 1..%d", N, CM3_MAX_AGENTS);
    return CM3_ERR_UNSUPPORTED;
}

// This file is compiled three times (cm3_b200/build.py): -DCM3_CK_REAL=0 instantiates the float
// kernels, =1 the double ones, =2 the float kernels with int8 tiles; the objects build in parallel.
#if CM3_CK_REAL == 0
// what the bitboards and the packed words can address (the header states the same limits)
bool checkers_geometry_supported(int R, int C, int O, int N) {
    return N >= 1 && N <= CM3_MAX_AGENTS && R >= 1 && C >= 2 && O >= 1 && O <= 3 && R * C <= 64 &&
           C + 2 * O + 1 <= 32 && R + 2 * O <= kCkMaxTR && C + 2 * O + 1 <= kCkMaxTC && R * C / 2 + 1 <= kCkMaxCnt;
}

int checkers_tile_envs(int N) { return kWarp / ck_lanes_per_env(N); }

int checkers_launch_f32(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream) {
    return dispatch_ck<float, float>(R, C, O, N, p, stream);
}

int checkers_launch(int R, int C, int O, int N, int real, int tile, const CkParams &p, cudaStream_t stream) {
    if (tile == CM3_TILE_I8 || tile == CM3_TILE_U2) {
        if (real != CM3_REAL_F32) {
            set_error("compact tiles are compiled for float outputs only");
            return CM3_ERR_UNSUPPORTED;
        }
        return tile == CM3_TILE_I8 ? checkers_launch_f32_i8(R, C, O, N, p, stream) : checkers_launch_f32_u2(R, C, O, N, p, stream);
    }
    if (real == CM3_REAL_F64) return checkers_launch_f64(R, C, O, N, p, stream);
    return checkers_launch_f32(R, C, O, N, p, stream);
}
#elif CM3_CK_REAL == 1
int checkers_launch_f64(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream) {
    return dispatch_ck<double, double>(R, C, O, N, p, stream);
}
#elif CM3_CK_REAL == 2
int checkers_launch_f32_i8(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream) {
    return dispatch_ck<float, int8_t>(R, C, O, N, p, stream);
}
#else
int checkers_launch_f32_u2(int R, int C, int O, int N, const CkParams &p, cudaStream_t stream) {
    return dispatch_ck<float, TileU2>(R, C, O, N, p, stream);
}
#endif

}  // namespace cm3
