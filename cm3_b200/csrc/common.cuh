// common.cuh - shared device helpers for the sm_100a env-step kernels.
#pragma once
#include <cuda.h>            // CUtensorMap (types only: the driver entry point is looked up at run time)
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/cm3env.h"

namespace cm3 {

constexpr int kWarp = 32;

// ---------------------------------------------------------------- host-side errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
#define CM3_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cm3::cuda_fail(e__, #call); \
    } while (0)

// ---------------------------------------------------------------- Real traits
template <typename Real> struct RealOps;
template <> struct RealOps<float> {
    // _rn intrinsics are never contracted into FMAs: the reference rounds after every
    // NumPy op, and so must we.
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float exp(float a) { return expf(a); }
    static __device__ __forceinline__ float log1p(float a) { return log1pf(a); }
    static __device__ __forceinline__ float log(float a) { return logf(a); }
    static __device__ __forceinline__ void sincospi(float a, float *s, float *c) { sincospif(a, s, c); }
};
template <> struct RealOps<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double exp(double a) { return ::exp(a); }
    static __device__ __forceinline__ double log1p(double a) { return ::log1p(a); }
    static __device__ __forceinline__ double log(double a) { return ::log(a); }
    static __device__ __forceinline__ void sincospi(double a, double *s, double *c) { ::sincospi(a, s, c); }
};

// ---------------------------------------------------------------- Philox4x32-10
// Salmon et al., SC'11.  The CPU twin is oracle_philox4x32_10 (oracle/cm3_oracle.c).
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

constexpr uint32_t kTagAction = 0xAC710000u;  // stream tags live in the top half of ctr[3]
constexpr uint32_t kTagReset = 0x5E5E0000u;      // explicit resets: counter = the caller's reset_counter
constexpr uint32_t kTagAutoReset = 0xA57E0000u;  // in-kernel resets: counter = global step index + 1

// Uniform action stream: counter = (env id, step index), agent i takes word i.
// Agents 4 blk .. 4 blk + 3 take the words of block blk (blk is folded into the top bits of the
// env-id half of the counter; block 0 is the round-1 stream).
__device__ __forceinline__ Philox4 philox_action_words(uint64_t seed, uint64_t env, uint64_t step, uint32_t blk = 0) {
    return philox4x32_10((uint32_t)env, (uint32_t)(env >> 32) ^ (blk << 24), (uint32_t)step,
                         kTagAction | (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ int action_from_word(uint32_t w, int n_actions) {
    return (int)__umulhi(w, (uint32_t)n_actions);
}
__device__ __forceinline__ uint32_t philox_word(const Philox4 &p, int i) {
    return i == 0 ? p.x : i == 1 ? p.y : i == 2 ? p.z : p.w;
}

// ---------------------------------------------------------------- L2 cache policies
// The kernels write a long, write-once output stream and re-read a small action stream and the
// compact state.  The particle kernel's tensor-map stores carry an evict-first policy so that the
// outputs do not push the actions out of the 126 MB L2 (measured, profiles/r01o_ab.txt: +5-7 % at
// 65 536 envs with actions from HBM).  The same hint on the Checkers kernel's plain bulk stores
// measured 2 % slower and is off (CM3_L2_HINT_BULK builds it); an evict-last hint on the cp.async
// action loads (cp.async...L2::cache_hint) raised an illegal-instruction fault on sm_100a, so the
// action rows get an evict-last L2 prefetch instead (ActionStream::prefetch).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// The thread of a warp that issues, commits and waits for its TMA stores.  Default: lane 0.
// Measured and dropped (round 2, profiles/r02g_ab.txt): electing it with elect.sync, which lets ptxas
// see that a single thread is active - under `if (lane == 0)` it wraps every TMA instruction in a
// loop that moves the uniform-register operands over one distinct value at a time, ~15 instructions
// per store.  The elected variant executes ~45 fewer instructions per particle step and is SLOWER on
// the same box, every time: PA4 0.869 vs 0.905 of the roofline, PA3 0.824 vs 0.856, CK1 0.904 vs
// 0.913 (CK2, PM2 within noise).  The back-to-back store issue apparently costs more than the
// operand loops did.  CM3_ELECT=1 rebuilds it.
#ifndef CM3_ELECT
#define CM3_ELECT 0
#endif
__device__ __forceinline__ bool elect_one() {
#if !CM3_ELECT
    return (threadIdx.x & 31) == 0;
#endif
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- action stream
// The N int8 actions of one env ([B][N] rows) as one packed word: a single 1/2/4-byte load for
// N = 1/2/4 (rows are naturally aligned), three byte loads for N = 3.
template <int N> __device__ __forceinline__ uint32_t load_actions_packed(const int8_t *row) {
    if (N == 4) return *reinterpret_cast<const uint32_t *>(row);
    if (N == 2) return *reinterpret_cast<const uint16_t *>(row);
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) w |= (uint32_t)(uint8_t)row[i] << (8 * i);
    return w;
}
__device__ __forceinline__ int unpack_action(uint32_t w, int i) { return (int)(int8_t)(w >> (8 * i)); }

// Multi-step launches stream the action rows of a warp's env tile through shared memory, one step
// ahead: at the top of the emit phase of step t every lane loads one 32-bit word of the rows of
// step t + 1 (they were pulled into L2, evict-last, two steps earlier), the load is in flight while
// the observation tiles are staged and their stores issued, and at the end of the phase the word is
// parked in shared memory, from where each thread picks up the packed actions of its own env.
// What round 1 measured on the way here:
//  * a register prefetch across the loop back-edge does not work: ptxas makes the first use of the
//    CURRENT word wait on the scoreboard of the load just issued for the NEXT one (27 % of all warp
//    stall samples on that instruction, profiles/r01g_pa4_fused_stalls.txt);
//  * cp.async (LDGSTS) works but shares its completion scoreboard with the bulk stores
//    (cp.async.wait_group and cp.async.bulk.wait_group.read are the same DEPBAR.LE SB0), so a
//    wait for the rows is also a wait for the warp's output tiles - which rules out keeping one
//    tile in flight while the next is staged (double-buffered staging, particle.cu);
//  * without the L2 prefetch the output stream evicts the action stream and every row comes from
//    HBM in between the writes (profiles/r01o_ab.txt: 2-4 % of the roofline).
// The tile's rows of one step are tile_envs * N contiguous bytes of the [T][B][N] int8 array.
template <int N>
struct ActionStream {
    static constexpr int kSlotBytes = 128;            // >= 32 envs * 4 agents
    static constexpr int kSmemBytes = 2 * kSlotBytes;  // double buffered
    unsigned char *slots;
    const int8_t *tile0;  // row of the tile's first env at step 0
    size_t step_bytes;    // B * N
    int nwords, lane, T;
    bool on;

    // on: whole tile, multi-step launch, every step's rows 4-byte aligned; otherwise the caller
    // falls back to direct loads
    __device__ __forceinline__ void init(unsigned char *smem, const int8_t *actions, int B, int env0, int tile_envs,
                                         bool whole_tile, int T_, int lane_) {
        slots = smem; lane = lane_; T = T_;
        step_bytes = (size_t)B * N;
        tile0 = actions + (size_t)env0 * N;
        nwords = tile_envs * N / 4;
        on = actions != nullptr && T > 1 && whole_tile && (tile_envs * N) % 4 == 0 && nwords <= kWarp &&
             ((reinterpret_cast<uintptr_t>(tile0) | step_bytes) & 3u) == 0;
    }
    // pulls the rows of step t into L2 with evict-last priority (one request per 32-byte sector)
    __device__ __forceinline__ void prefetch(int t) const {
        if (t < T && lane < nwords && (lane & 7) == 0)
            asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(tile0 + (size_t)t * step_bytes + 4 * lane));
    }
    // this lane's word of the rows of step t; volatile so that it stays where it is written (early)
    __device__ __forceinline__ uint32_t load(int t) const {
        uint32_t w = 0;
        if (t < T && lane < nwords)
            asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(w) : "l"(tile0 + (size_t)t * step_bytes + 4 * lane) : "memory");
        return w;
    }
    // parks the word loaded for step t in slot t & 1 (volatile: stays behind the store issue)
    __device__ __forceinline__ void stash(int t, uint32_t w) const {
        if (lane < nwords)
            asm volatile("st.shared.b32 [%0], %1;" :: "r"(smem_u32(slots + (t & 1) * kSlotBytes + 4 * lane)), "r"(w) : "memory");
    }
    // packed action word (byte i = agent i) of local env e of the tile at step t
    __device__ __forceinline__ uint32_t read(int t, int e) const {
        const unsigned char *row = slots + (t & 1) * kSlotBytes + e * N;
        if (N == 4) return *reinterpret_cast<const volatile uint32_t *>(row);
        if (N == 2) return *reinterpret_cast<const volatile uint16_t *>(row);
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) w |= (uint32_t)reinterpret_cast<const volatile unsigned char *>(row)[i] << (8 * i);
        return w;
    }
    // launch prologue: the word of step 0 (exposed latency, once per launch)
    __device__ __forceinline__ uint32_t begin(int e) const {
        prefetch(1);
        prefetch(2);
        stash(0, load(0));
        __syncwarp();
        return read(0, e);
    }
    // end of emit(t): the word loaded at its top becomes the packed actions of step t + 1
    __device__ __forceinline__ uint32_t hand_over(int t, uint32_t loaded, int e) const {
        stash(t + 1, loaded);
        __syncwarp();
        return read(t + 1, e);
    }
};

// ---------------------------------------------------------------- TMA bulk store
// Shared -> global bulk copy (SASS: UBLKCP).  Source and destination 16-byte aligned, size a
// multiple of 16.  Issued by ONE thread after every writer executed fence_proxy_async() and
// the writers were synchronised with the issuer.
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the smem source of every committed group has been read (safe to overwrite the staging tile)
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// ... of every committed group but the NEWEST (double-buffered staging: the tile staged one step ago
// may still be in flight while the other buffer is refilled)
__device__ __forceinline__ void bulk_wait_read_but_one() {
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
// every committed group has fully completed
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// TMA tensor store of one box (SASS: UTMASTG): shared tile -> the box at coordinates (c0, c1) of
// the 2-D tensor described by `tm` (a __grid_constant__ kernel parameter), un-doing the map's
// swizzle pattern.  Same bulk-group completion mechanism as bulk_store.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, const void *ssrc, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(ssrc)), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap *tm, const void *ssrc, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 :: "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(ssrc)), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}

// brings a tensor map into the TMA descriptor cache ahead of its first use
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------- programmatic dependent launch
// Consecutive step launches on a stream depend on each other only through the compact state.
// With the launch attribute below the next grid becomes resident (block scheduling, parameter
// and LUT setup) while the previous grid drains its last stores, and blocks at pdl_wait() until
// the previous grid has completed and its writes are visible.  Both instructions are no-ops for
// a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
#ifdef CM3_EXP_NO_PDL_WAIT  // timing experiment only (racy results): the per-step bound without the
                            // inter-launch dependency, profiles/r01g_nowait.txt
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---------------------------------------------------------------- per-tile launch chaining
// Consecutive launches on one compact state depend on each other TILE BY TILE: tile i of launch
// k + 1 needs the state tile i of launch k wrote, nothing else.  griddepcontrol.wait expresses a
// much coarser dependency (the WHOLE previous grid has completed and flushed), which serialises
// single-step launches into compute phase / store phase / drain (round 1: 12.7 us per CK2 step where
// the stores alone take 9.6).  A ticket lock per tile expresses the real one:
//   sync[2 * tile + 0]  tickets handed out (atomicAdd at kernel entry, BEFORE launch_dependents),
//   sync[2 * tile + 1]  launches that have finished this tile (st.release after the state stores).
// Every launch takes a ticket and publishes; a CHAINED launch (cm3_*_step_chained) waits for
// "finished == my ticket" with ld.acquire instead of executing griddepcontrol.wait.  The words live
// with the state (cm3_*_state.sync, zero-initialised), so the protocol needs no host-side sequence
// number and survives CUDA-graph replay.
// No deadlock: a dependent grid is only scheduled after EVERY block of its predecessor has executed
// launch_dependents - i.e. is resident and holds its ticket - so the block a waiter spins on is
// always running or finished.  The trigger is issued under a branch on the ticket value: the atomic
// has RETURNED (was performed at L2) before any block of the next grid can take its own ticket.
// (Measured and dropped, profiles/r02q_ab.txt: issuing the atomic first and consuming its result only
// after the block's set-up - to hide its round trip - changes nothing; a block that starts late is
// waiting for a slot, not for its ticket.)
struct TileTicket {
    uint32_t *w;
    uint32_t mine;
    bool on;
    __device__ __forceinline__ void take(uint32_t *sync, int tile, int lane) {
        on = sync != nullptr;
        mine = 0;
        if (on) {
            w = sync + 2 * (size_t)tile;
            if (lane == 0) mine = atomicAdd(w, 1u);
            mine = __shfl_sync(0xFFFFFFFFu, mine, 0);
        }
    }
    // blocks until every earlier launch has finished this tile; bounded (a lost predecessor traps
    // instead of hanging the device)
    __device__ __forceinline__ void wait(int lane) const {
        if (!on) return;
        if (lane == 0) {
            uint32_t v;
            unsigned spins = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(w + 1) : "memory");
                if (v == mine) break;
                __nanosleep(40);
                if (++spins > (1u << 25)) __trap();
            }
        }
        __syncwarp();
    }
    // after the state stores of this tile
    __device__ __forceinline__ void publish(int lane) const {
        if (!on) return;
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(w + 1), "r"(mine + 1u) : "memory");
    }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): raised to the device's
// opt-in maximum once per pair, safely from any number of host threads (one bit per device ordinal;
// ordinals >= 64 set it every time)
template <typename K>
inline cudaError_t ensure_smem_attr(K kern, std::atomic<uint64_t> &done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = dev < 64 ? (1ull << dev) : 0ull;
    if (bit && (done_mask.load(std::memory_order_acquire) & bit)) return cudaSuccess;
    int optin = 0;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e == cudaSuccess && bit) done_mask.fetch_or(bit, std::memory_order_release);
    return e;
}

// Wave balancing for multi-step launches (api.cu).  A T-step rollout keeps a block on its tile for
// the whole launch, so a grid that needs a little more than w full waves of resident blocks ends
// with a nearly empty wave running at a fraction of the machine (stage-1 Checkers at 65 536 envs:
// 4096 tiles over 3404 slots - the last 692 blocks ran alone for a quarter of the launch).  Where the
// last wave would be less than half full, the launch asks for enough EXTRA dynamic shared memory to
// lower the blocks per SM to ceil(blocks / (waves * SMs)): the same number of waves, all of them
// full.  Returns the dynamic shared-memory size to launch with (>= smem).
int balance_waves(const void *kern, int threads, int smem, int nblocks);

template <typename P>
inline cudaError_t launch_kernel(void (*kern)(P), int nblocks, int nthreads, size_t smem, cudaStream_t stream,
                                 bool pdl, const P &params) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nblocks, 1, 1);
    cfg.blockDim = dim3((unsigned)nthreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, params);
}

}  // namespace cm3
