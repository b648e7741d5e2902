// common.cuh - shared device helpers for the sm_100a env-step kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
constexpr uint32_t kTagReset = 0x5E5E0000u;

// Uniform action stream: counter = (env id, step index), agent i takes word i.
__device__ __forceinline__ Philox4 philox_action_words(uint64_t seed, uint64_t env, uint64_t step) {
    return philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)step,
                         kTagAction | (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ int action_from_word(uint32_t w, int n_actions) {
    return (int)__umulhi(w, (uint32_t)n_actions);
}
__device__ __forceinline__ uint32_t philox_word(const Philox4 &p, int i) {
    return i == 0 ? p.x : i == 1 ? p.y : i == 2 ? p.z : p.w;
}

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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the smem source of every committed group has been read (safe to overwrite the staging tile)
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// every committed group has fully completed
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__host__ __device__ constexpr int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------- programmatic dependent launch
// Consecutive step launches on a stream depend on each other only through the compact state.
// With the launch attribute below the next grid becomes resident (block scheduling, parameter
// and LUT setup) while the previous grid drains its last stores, and blocks at pdl_wait() until
// the previous grid has completed and its writes are visible.  Both instructions are no-ops for
// a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
