// api.cu - the extern "C" boundary declared in include/cm3env.h.
// Handles hold configuration only; every buffer is the caller's (see the header).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "params.cuh"

namespace cm3 {

static thread_local char g_last_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
        set_error("%s: %s - this library has no CPU fallback", what, cudaGetErrorString(e));
        return CM3_ERR_NO_DEVICE;
    }
    set_error("%s: %s", what, cudaGetErrorString(e));
    return CM3_ERR_CUDA;
}

static int require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device visible (%s) - this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return CM3_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (0..%d)", device, n - 1);
        return CM3_ERR_BAD_ARG;
    }
    return CM3_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

bool pdl_enabled() {
    static const bool on = [] {
        const char *v = getenv("CM3_PDL");
        return !(v && v[0] == '0');
    }();
    return on;
}

bool full_enabled() {
    static const bool on = [] {
        const char *v = getenv("CM3_PT_FULL");
        return !(v && v[0] == '0');
    }();
    return on;
}

int chain_parts(int ntiles) {
    static const int forced = [] {
        const char *v = getenv("CM3_CHAIN_PARTS");
        return v ? atoi(v) : 0;
    }();
    const int parts = forced > 0 ? forced : 1;  // measured: 2, 3, 4 parts are all slower (profiles/r02c_ab.txt)
    return ntiles >= 296 * parts ? parts : 1;   // keep at least two blocks per SM in every partial grid
}

// resident blocks of `kern` on the current device (occupancy query, cached per (kernel, device, smem)); 0 on failure
static long resident_slots(const void *kern, int threads, int smem) {
    struct Entry { const void *kern; int dev, smem; long resident; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (const Entry &e : cache)
            if (e.kern == kern && e.dev == dev && e.smem == smem) return e.resident;
    }
    int sms = 0, per_sm = 0;
    long resident = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) == cudaSuccess)
        resident = (long)per_sm * sms;
    (void)cudaGetLastError();
    std::lock_guard<std::mutex> lk(mu);
    cache.push_back(Entry{kern, dev, smem, resident});
    return resident;
}

int chain_early_mode(const void *kern, int threads, int smem, int nblocks) {
    static const int forced = [] {
        const char *v = getenv("CM3_CHAIN_EARLY");
        return v ? atoi(v) : -1;
    }();
    if (forced >= 0) return forced > 2 ? 2 : forced;
    return nblocks <= resident_slots(kern, threads, smem) ? 2 : 0;  // one resident wave -> 2, else 0 (params.cuh)
}

int chain_tiles_per_block(const void *kern, int threads, int smem, int nblocks) {
    static const int forced = [] {
        const char *v = getenv("CM3_CHAIN_TPB");
        return v ? atoi(v) : 0;
    }();
    if (forced > 0) return forced > 4 ? 4 : forced;
    // as many tiles per block as it takes to make the launch one resident wave, at most 4 (params.cuh)
    const long resident = resident_slots(kern, threads, smem);
    if (resident <= 0 || nblocks <= resident) return 1;
    const long tpb = (nblocks + resident - 1) / resident;
    return tpb > 4 ? 4 : (int)tpb;
}

int particle_chain_tpb(const void *kern, int threads, int smem, int nblocks) {
    static const int forced = [] {
        const char *v = getenv("CM3_PT_TPB");
        return v ? atoi(v) : 0;
    }();
    if (forced > 0) {
        const int t = forced > 4 ? 4 : forced;
        return nblocks >= 2 * t ? t : 1;
    }
    // the one-wave rule of the Checkers launches (multi-wave batches: 131 072 envs and more)
    const long resident = resident_slots(kern, threads, smem);
    if (resident <= 0 || nblocks <= resident) return 1;
    const long tpb = (nblocks + resident - 1) / resident;
    return tpb > 4 ? 4 : (int)tpb;
}

// Off by default: measured slower than one env per thread (profiles/r02k_ab.txt, r02l_ab.txt)
bool duo_enabled() {
    static const bool on = [] {
        const char *v = getenv("CM3_PT_DUO");
        return v && v[0] == '1';
    }();
    return on;
}

// Off by default: measured slower than one thread per env on the fused rollout (profiles/r02h_ab.txt)
bool pair_enabled() {
    static const bool on = [] {
        const char *v = getenv("CM3_PT_PAIR");
        return v && v[0] == '1';
    }();
    return on;
}

// two-agent envs run one lane per agent (16 envs per warp); the sync words are sized for that tiling
// whichever kernel ends up stepping them
int particle_tile_envs(int N) { return N == 2 ? kWarp / 2 : kWarp; }

bool dyn_geometry_forced() {
    static const bool on = [] {
        const char *v = getenv("CM3_CK_DYNAMIC");
        return v && v[0] == '1';
    }();
    return on;
}

bool tma_enabled() {
    static const bool on = [] {
        const char *v = getenv("CM3_TMA");
        return !(v && v[0] == '0');
    }();
    return on;
}

bool balance_enabled() {
    static const bool on = [] {
        const char *v = getenv("CM3_BALANCE");
        return !(v && v[0] == '0');
    }();
    return on;
}

int balance_waves(const void *kern, int threads, int smem, int nblocks) {
    if (!balance_enabled()) return smem;
    struct Entry { const void *kern; int dev, smem, nblocks, result; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return smem;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (const Entry &e : cache)
            if (e.kern == kern && e.dev == dev && e.smem == smem && e.nblocks == nblocks) return e.result;
    }
    int result = smem;
    int sms = 0, smem_sm = 0, optin = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) == cudaSuccess && per_sm > 0 && sms > 0) {
        const long resident = (long)per_sm * sms;
        if (nblocks > resident) {
            const long waves = (nblocks + resident - 1) / resident;
            const long last = nblocks - (waves - 1) * resident;
            const int target = (int)((nblocks + waves * sms - 1) / (waves * sms));  // blocks per SM of equal waves
            // rebalance when the last wave would be less than half full (CM3_BALANCE_FRAC overrides the 0.5).
            // A fuller last wave is better left alone: CK2's is 73 % full, and forcing it into two equal waves of
            // 14 blocks per SM (CM3_BALANCE_FRAC=0.9) costs more in resident blocks than the tail returns -
            // 0.988 -> 0.875 of the roofline at K = 3300, 0.93 -> 0.84 on a single 20-step launch (run r02n).
            static const double frac = [] { const char *v = getenv("CM3_BALANCE_FRAC"); return v ? atof(v) : 0.5; }();
            if ((double)last < frac * (double)resident && target < per_sm) {
                // grow the request until the occupancy calculator agrees with the target
                int lo = smem, hi = std::min(optin, smem_sm / target);
                for (int s = hi; s >= lo; s -= 1024) {
                    int got = 0;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&got, kern, threads, s) != cudaSuccess) break;
                    if (got >= target) { if (got == target) result = s; break; }
                }
            }
        }
    }
    (void)cudaGetLastError();
    std::lock_guard<std::mutex> lk(mu);
    cache.push_back(Entry{kern, dev, smem, nblocks, result});
    return result;
}

static size_t real_size(int real) { return real == CM3_REAL_F64 ? 8 : 4; }

// Device-to-host delivery of the output fields of *_step_host: one copy per requested field.
struct FieldCopy { void *dst; const void *src; size_t bytes; };

static int copy_fields_to_host(const FieldCopy *cp, int n, cudaStream_t s) {
    for (int i = 0; i < n; ++i) {
        if (!cp[i].dst) continue;
        if (!cp[i].src) { set_error("host output requested for a field with no device buffer"); return CM3_ERR_BAD_ARG; }
        CM3_CUDA(cudaMemcpyAsync(cp[i].dst, cp[i].src, cp[i].bytes, cudaMemcpyDeviceToHost, s));
    }
    return CM3_OK;
}

// *_step_host_packed: every device field must lie inside the caller's block, which then travels
// as ONE copy (the facades carve the single-step outputs and their pinned mirror out of one
// allocation each, cm3_b200/_buffers.py).
static int check_fields_in_block(const FieldCopy *cp, int n, const void *dev_block, size_t block_bytes) {
    const char *lo = (const char *)dev_block, *hi = lo + block_bytes;
    for (int i = 0; i < n; ++i) {
        const char *f = (const char *)cp[i].src;
        if (f && (f < lo || f + cp[i].bytes > hi)) {
            set_error("output field %d lies outside [dev_block, dev_block + block_bytes)", i);
            return CM3_ERR_BAD_ARG;
        }
    }
    return CM3_OK;
}

}  // namespace cm3

using namespace cm3;

// copy stream + events of *_rollout_host (created on first use, on the handle's device)
struct HostPipe {
    cudaStream_t copy = nullptr;
    cudaEvent_t stepped[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    bool ready = false;
    int init() {
        if (ready) return CM3_OK;
        CM3_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CM3_CUDA(cudaEventCreateWithFlags(&stepped[i], cudaEventDisableTiming));
            CM3_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        }
        ready = true;
        return CM3_OK;
    }
    void destroy() {
        if (copy) cudaStreamDestroy(copy);
        for (int i = 0; i < 2; ++i) {
            if (stepped[i]) cudaEventDestroy(stepped[i]);
            if (copied[i]) cudaEventDestroy(copied[i]);
        }
        copy = nullptr; ready = false;
    }
};

// The double-buffered host rollout shared by both games: step(t, slot) enqueues the kernel of step t
// writing output set `slot` on the compute stream.
template <typename StepFn>
static int rollout_host_loop(HostPipe &hp, const int8_t *actions_host, int8_t *actions_dev, size_t act_bytes, int T,
                             const void *const *dev_blocks, void *host_blocks, size_t block_bytes, size_t host_stride,
                             cudaStream_t s, StepFn step) {
    int rc = hp.init();
    if (rc != CM3_OK) return rc;
    for (int t = 0; t < T; ++t) {
        const int slot = t & 1;
        if (t >= 2) CM3_CUDA(cudaStreamWaitEvent(s, hp.copied[slot], 0));  // set `slot` has left the device
        CM3_CUDA(cudaMemcpyAsync(actions_dev + (size_t)slot * act_bytes, actions_host + (size_t)t * act_bytes, act_bytes,
                                 cudaMemcpyHostToDevice, s));
        if ((rc = step(t, slot)) != CM3_OK) return rc;
        CM3_CUDA(cudaEventRecord(hp.stepped[slot], s));
        CM3_CUDA(cudaStreamWaitEvent(hp.copy, hp.stepped[slot], 0));
        CM3_CUDA(cudaMemcpyAsync((char *)host_blocks + (size_t)t * host_stride, dev_blocks[slot], block_bytes,
                                 cudaMemcpyDeviceToHost, hp.copy));
        CM3_CUDA(cudaEventRecord(hp.copied[slot], hp.copy));
    }
    CM3_CUDA(cudaStreamSynchronize(hp.copy));
    CM3_CUDA(cudaStreamSynchronize(s));
    return CM3_OK;
}

struct cm3_checkers_s {
    cm3_checkers_config cfg;
    CkParams base;  // geometry-derived constants, pointers zero
    HostPipe pipe;
};
struct cm3_particle_s {
    cm3_particle_config cfg;
    PtParams base;
    HostPipe pipe;
};

// the kernels read an env's N action bytes as one word (common.cuh: load_actions_packed)
static int check_actions_alignment(const int8_t *actions, int N) {
    const uintptr_t need = (N == 4) ? 4 : (N == 2) ? 2 : 1;
    if (actions && (reinterpret_cast<uintptr_t>(actions) % need) != 0) {
        set_error("actions must be %d-byte aligned for n_agents=%d", (int)need, N);
        return CM3_ERR_BAD_ARG;
    }
    return CM3_OK;
}

static CkOut ck_out(const cm3_checkers_outputs &o) {
    return CkOut{(char *)o.grid, (char *)o.vec, (char *)o.obs_others, (char *)o.obs_self_t, (char *)o.obs_self_v,
                 (char *)o.reward, (char *)o.local_rewards, o.done, o.goal_idx};
}

// rollout_gather: n_dst destination sets with identical NULL patterns, this shard at rows
// [dst_env0, dst_env0 + B) of [T][dst_B][...]
template <typename Out, typename POut, typename Conv>
static int fill_destinations(int32_t n_dst, const Out *dsts, int64_t dst_B, int64_t dst_env0, int B,
                             POut *out, int &n_out, long long &out_B, long long &out_env0, Conv conv) {
    if (n_dst < 1 || n_dst > CM3_MAX_DST || !dsts) {
        set_error("n_dst must be 1..%d and dsts non-NULL", CM3_MAX_DST);
        return CM3_ERR_BAD_ARG;
    }
    if (dst_env0 < 0 || dst_B < dst_env0 + B) {
        set_error("destination rows [%lld, %lld) do not fit dst_B=%lld", (long long)dst_env0,
                  (long long)(dst_env0 + B), (long long)dst_B);
        return CM3_ERR_BAD_ARG;
    }
    const void *const *first = reinterpret_cast<const void *const *>(&dsts[0]);
    for (int d = 0; d < n_dst; ++d) {
        const void *const *f = reinterpret_cast<const void *const *>(&dsts[d]);
        for (size_t i = 0; i < sizeof(Out) / sizeof(void *); ++i)
            if ((f[i] == nullptr) != (first[i] == nullptr)) {
                set_error("destination %d differs from destination 0 in which fields are NULL", d);
                return CM3_ERR_BAD_ARG;
            }
        out[d] = conv(dsts[d]);
    }
    n_out = n_dst; out_B = dst_B; out_env0 = dst_env0;
    return CM3_OK;
}

static PtOut pt_out(const cm3_particle_outputs &o) {
    return PtOut{(char *)o.global_state, (char *)o.obs_others, (char *)o.obs_self, (char *)o.reward,
                 (char *)o.reward_n, o.done, o.collisions, o.reached};
}

extern "C" {

int cm3_abi_version(void) { return CM3_ABI_VERSION; }
const char *cm3_last_error(void) { return g_last_error; }

int cm3_device_count(int *count) {
    if (!count) { set_error("count is NULL"); return CM3_ERR_BAD_ARG; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    *count = n;
    return CM3_OK;
}

int cm3_stream_synchronize(void *stream) {
    CM3_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return CM3_OK;
}

/* ------------------------------------------------------------------ Checkers */

int cm3_checkers_create(const cm3_checkers_config *cfg, cm3_checkers_t *out) {
    if (!cfg || !out) { set_error("cfg/out is NULL"); return CM3_ERR_BAD_ARG; }
    *out = nullptr;
    const int R = cfg->n_rows, C = cfg->n_columns, O = cfg->n_obs, N = cfg->n_agents;
    // the reference's own asserts, env/checkers.py:16-17
    if (R < 1 || C < 2 || R % 2 != 1 || C % 2 != 0) {
        set_error("n_rows must be odd and n_columns even (got %d, %d)", R, C);
        return CM3_ERR_BAD_SHAPE;
    }
    if (O < 1 || N < 1 || cfg->num_envs < 1 || cfg->max_steps < 0 || cfg->max_steps > 0xFFFFFF ||
        (cfg->real != CM3_REAL_F32 && cfg->real != CM3_REAL_F64) ||
        (cfg->tile != CM3_TILE_REAL && cfg->tile != CM3_TILE_I8 && cfg->tile != CM3_TILE_U2)) {
        set_error("bad n_obs/n_agents/num_envs/max_steps/real/tile");
        return CM3_ERR_BAD_ARG;
    }
    if (cfg->tile != CM3_TILE_REAL && cfg->real != CM3_REAL_F32) {
        set_error("compact tiles are compiled for float outputs only");
        return CM3_ERR_UNSUPPORTED;
    }
    if (N > CM3_MAX_AGENTS || !checkers_geometry_supported(R, C, O, N)) {
        set_error("n_rows=%d n_columns=%d n_obs=%d n_agents=%d is outside what the bitboards address "
                  "(n_rows*n_columns <= 64, n_columns + 2 n_obs + 1 <= 32, n_obs <= 3, n_agents <= %d)", R, C, O, N, CM3_MAX_AGENTS);
        return CM3_ERR_UNSUPPORTED;
    }
    if (cfg->random_goal != 0 && cfg->random_goal != 1) { set_error("random_goal must be 0 or 1"); return CM3_ERR_BAD_ARG; }
    if (cfg->random_goal && N != 1) {
        set_error("random_goal is the stage-1 (n_agents == 1) episode protocol (train_offpolicy.py:291-296)");
        return CM3_ERR_BAD_ARG;
    }
    if (N == 1 && R < 3) {  // checkers.py:276 starts the agent on row 2
        set_error("n_agents == 1 needs n_rows >= 3 (checkers.py:276)");
        return CM3_ERR_BAD_SHAPE;
    }
    for (int i = 0; i < N; ++i) {
        // bitboard occupancy == the reference's world[..,2] only while agents stand on distinct
        // cells of the valid grid; the reference configs all satisfy this
        if (cfg->agents_r[i] < 0 || cfg->agents_r[i] >= R || cfg->agents_c[i] < 0 || cfg->agents_c[i] > C) {
            set_error("agent %d starts outside the valid grid", i);
            return CM3_ERR_BAD_ARG;
        }
        for (int j = 0; j < i; ++j)
            if (N > 1 && cfg->agents_r[i] == cfg->agents_r[j] && cfg->agents_c[i] == cfg->agents_c[j]) {
                set_error("agents %d and %d start on the same cell", j, i);
                return CM3_ERR_BAD_ARG;
            }
    }
    int st = require_device(cfg->device);
    if (st != CM3_OK) return st;

    cm3_checkers_s *h = new (std::nothrow) cm3_checkers_s();
    if (!h) { set_error("out of host memory"); return CM3_ERR_BAD_ARG; }
    h->cfg = *cfg;
    CkParams &p = h->base;
    memset(&p, 0, sizeof(p));
    p.B = cfg->num_envs;
    p.max_steps = cfg->max_steps;
    p.env_id_offset = cfg->env_id_offset;
    p.R = R; p.C = C; p.O = O;
    p.random_goal = cfg->random_goal;
    for (int i = 0; i < R; ++i)  // cells (i,j) with (i+j) even are green, odd orange (checkers.py:54-63)
        for (int j = 0; j < C; ++j) p.color_mask[(i + j) & 1] |= 1ull << (i * C + j);
    const int TR = R + 2 * O, TC = C + 2 * O + 1;  // checkers.py:24-25
    for (int i = 0; i < N; ++i) {
        p.start_r[i] = cfg->agents_r[i] + O;  // checkers.py:34-35
        p.start_c[i] = cfg->agents_c[i] + O;
    }
    // normalize(), checkers.py:120-121, and collected/(max_collectible/2.0), :139 - same
    // float64 expressions as the reference, tabulated over every reachable integer
    for (int r = 0; r < TR; ++r) p.norm_row[r] = ((double)r - TR / 2.0) / TR;
    for (int c = 0; c < TC; ++c) p.norm_col[c] = ((double)c - TC / 2.0) / TC;
    const int max_collectible = R * C;
    for (int k = 0; k <= max_collectible / 2; ++k) p.norm_cnt[k] = (double)k / (max_collectible / 2.0);
    *out = h;
    return CM3_OK;
}

int cm3_checkers_destroy(cm3_checkers_t h) {
    if (!h) { set_error("handle is NULL"); return CM3_ERR_BAD_ARG; }
    if (h->pipe.ready) { DeviceGuard g(h->cfg.device); h->pipe.destroy(); }
    delete h;
    return CM3_OK;
}

int cm3_checkers_tiles(cm3_checkers_t h, int32_t *tiles) {
    if (!h || !tiles) { set_error("handle/tiles is NULL"); return CM3_ERR_BAD_ARG; }
    const int ew = checkers_tile_envs(h->cfg.n_agents);
    *tiles = (h->cfg.num_envs + ew - 1) / ew;
    return CM3_OK;
}

static int ck_fill(cm3_checkers_t h, const cm3_checkers_state *st, const cm3_checkers_outputs *outs,
                   CkParams &p) {
    if (!h || !st || !st->remaining || !st->agents || !st->meta) {
        set_error("handle/state pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    p = h->base;
    p.remaining = st->remaining; p.agents = st->agents; p.meta = st->meta; p.sync = st->sync;
    p.n_dst = 1; p.out_B = p.B; p.out_env0 = 0;
    if (outs) p.out[0] = ck_out(*outs);
    return CM3_OK;
}

static int ck_launch(cm3_checkers_t h, const CkParams &p, void *stream) {
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    return checkers_launch(h->cfg.n_rows, h->cfg.n_columns, h->cfg.n_obs, h->cfg.n_agents, h->cfg.real, h->cfg.tile, p,
                           (cudaStream_t)stream);
}

int cm3_checkers_reset(cm3_checkers_t h, const cm3_checkers_state *st, const uint8_t *goal_idx,
                       const uint8_t *env_mask, const cm3_checkers_outputs *outs, void *stream) {
    CkParams p;
    int rc = ck_fill(h, st, outs, p);
    if (rc != CM3_OK) return rc;
    p.mode = 1; p.T = 1;
    p.goal_idx = goal_idx; p.env_mask = env_mask;
    p.out[0].reward = nullptr; p.out[0].local_rewards = nullptr;
    return ck_launch(h, p, stream);
}

int cm3_checkers_rollout(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                         uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset, int8_t *actions_out,
                         const cm3_checkers_outputs *outs, void *stream) {
    CkParams p;
    int rc = ck_fill(h, st, outs, p);
    if (rc != CM3_OK) return rc;
    if (T < 1) { set_error("T must be >= 1"); return CM3_ERR_BAD_ARG; }
    if ((rc = check_actions_alignment(actions, h->cfg.n_agents)) != CM3_OK) return rc;
    p.mode = 0; p.T = T; p.auto_reset = auto_reset ? 1 : 0;
    p.actions = actions; p.actions_out = actions_out;
    p.seed = seed; p.t0 = t0;
    return ck_launch(h, p, stream);
}

int cm3_checkers_rollout_gather(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                                uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset, int8_t *actions_out,
                                int32_t n_dst, const cm3_checkers_outputs *dsts, int64_t dst_B,
                                int64_t dst_env0, void *stream) {
    CkParams p;
    int rc = ck_fill(h, st, nullptr, p);
    if (rc != CM3_OK) return rc;
    if (T < 1) { set_error("T must be >= 1"); return CM3_ERR_BAD_ARG; }
    rc = fill_destinations(n_dst, dsts, dst_B, dst_env0, p.B, p.out, p.n_dst, p.out_B, p.out_env0, ck_out);
    if (rc != CM3_OK) return rc;
    if ((rc = check_actions_alignment(actions, h->cfg.n_agents)) != CM3_OK) return rc;
    p.mode = 0; p.T = T; p.auto_reset = auto_reset ? 1 : 0;
    p.actions = actions; p.actions_out = actions_out;
    p.seed = seed; p.t0 = t0;
    return ck_launch(h, p, stream);
}

int cm3_checkers_step(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                      const cm3_checkers_outputs *outs, void *stream) {
    if (!actions) { set_error("actions is NULL"); return CM3_ERR_BAD_ARG; }
    return cm3_checkers_rollout(h, st, actions, 0, 0, 1, 0, nullptr, outs, stream);
}

int cm3_checkers_step_chained(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions,
                              uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_checkers_outputs *outs, void *stream) {
    if (!actions) { set_error("actions is NULL"); return CM3_ERR_BAD_ARG; }
    CkParams p;
    int rc = ck_fill(h, st, outs, p);
    if (rc != CM3_OK) return rc;
    if (!st->sync) { set_error("step_chained needs state.sync (the per-tile chaining words)"); return CM3_ERR_BAD_ARG; }
    if ((rc = check_actions_alignment(actions, h->cfg.n_agents)) != CM3_OK) return rc;
    p.mode = 0; p.T = 1; p.auto_reset = auto_reset ? 1 : 0; p.chained = 1;
    p.early = -1;  // resolved per launch shape (params.cuh: chain_early_mode)
    p.actions = actions; p.seed = seed; p.t0 = t0;
    return ck_launch(h, p, stream);
}

static int ck_step_host(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions_host,
                        int8_t *actions_dev, const cm3_checkers_outputs *od, const cm3_checkers_outputs *oh,
                        const void *dev_block, void *host_block, size_t block_bytes, void *stream) {
    if (!h || !actions_host || !actions_dev || !od || (!oh && !host_block) || (host_block && !dev_block)) {
        set_error("handle/actions/outputs pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = h->cfg.num_envs, N = h->cfg.n_agents, rs = real_size(h->cfg.real);
    const size_t W = 2 * h->cfg.n_obs + 1, L = 2 * (N > 1 ? N - 1 : 1);
    const size_t ts = h->cfg.tile == CM3_TILE_I8 ? 1 : rs;
    const bool u2 = h->cfg.tile == CM3_TILE_U2;
    // bytes per env of the two tile fields (CM3_TILE_U2: packed rows, see the header)
    const size_t grid_b = u2 ? (size_t)h->cfg.n_rows * ((h->cfg.n_columns + 1 + 7) / 8) * 4 : (size_t)h->cfg.n_rows * (h->cfg.n_columns + 1) * 2 * ts;
    const size_t win_b = u2 ? W * ((6 * W + 31) / 32) * 4 : W * W * 3 * ts;
    static const cm3_checkers_outputs none = {};
    if (!oh) oh = &none;
    const FieldCopy cp[] = {
        {oh->grid, od->grid, B * grid_b},
        {oh->vec, od->vec, B * N * 4 * rs},
        {oh->obs_others, od->obs_others, B * N * L * rs},
        {oh->obs_self_t, od->obs_self_t, B * N * win_b},
        {oh->obs_self_v, od->obs_self_v, B * N * 4 * rs},
        {oh->reward, od->reward, B * rs},
        {oh->local_rewards, od->local_rewards, B * N * rs},
        {oh->done, od->done, B},
        {oh->goal_idx, od->goal_idx, B * N},
    };
    const int n = (int)(sizeof(cp) / sizeof(cp[0]));
    int rc;
    if (host_block && (rc = check_fields_in_block(cp, n, dev_block, block_bytes)) != CM3_OK) return rc;
    CM3_CUDA(cudaMemcpyAsync(actions_dev, actions_host, B * N, cudaMemcpyHostToDevice, s));
    if ((rc = cm3_checkers_step(h, st, actions_dev, od, stream)) != CM3_OK) return rc;
    if (host_block) CM3_CUDA(cudaMemcpyAsync(host_block, dev_block, block_bytes, cudaMemcpyDeviceToHost, s));
    else if ((rc = copy_fields_to_host(cp, n, s)) != CM3_OK) return rc;
    CM3_CUDA(cudaStreamSynchronize(s));
    return CM3_OK;
}

int cm3_checkers_step_host(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions_host,
                           int8_t *actions_dev, const cm3_checkers_outputs *od,
                           const cm3_checkers_outputs *oh, void *stream) {
    if (!oh) { set_error("handle/actions/outputs pointer is NULL"); return CM3_ERR_BAD_ARG; }
    return ck_step_host(h, st, actions_host, actions_dev, od, oh, nullptr, nullptr, 0, stream);
}

int cm3_checkers_step_host_packed(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions_host,
                                  int8_t *actions_dev, const cm3_checkers_outputs *od, const void *dev_block,
                                  void *host_block, size_t block_bytes, void *stream) {
    if (!dev_block || !host_block) { set_error("dev_block/host_block is NULL"); return CM3_ERR_BAD_ARG; }
    return ck_step_host(h, st, actions_host, actions_dev, od, nullptr, dev_block, host_block, block_bytes, stream);
}

int cm3_checkers_rollout_host(cm3_checkers_t h, const cm3_checkers_state *st, const int8_t *actions_host,
                              int8_t *actions_dev, int32_t T, uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_checkers_outputs *outs_dev, const void *const *dev_blocks,
                              void *host_blocks, size_t block_bytes, size_t host_stride, void *stream) {
    if (!h || !actions_host || !actions_dev || !outs_dev || !dev_blocks || !dev_blocks[0] || !dev_blocks[1] || !host_blocks) {
        set_error("handle/actions/outputs/blocks pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    if (T < 1 || host_stride < block_bytes) { set_error("T must be >= 1 and host_stride >= block_bytes"); return CM3_ERR_BAD_ARG; }
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    const size_t act_bytes = (size_t)h->cfg.num_envs * h->cfg.n_agents;
    return rollout_host_loop(h->pipe, actions_host, actions_dev, act_bytes, T, dev_blocks, host_blocks, block_bytes,
                             host_stride, (cudaStream_t)stream, [&](int t, int slot) {
                                 return cm3_checkers_rollout(h, st, actions_dev + (size_t)slot * act_bytes, seed, t0 + t, 1,
                                                             auto_reset, nullptr, &outs_dev[slot], stream);
                             });
}

static int ck_state_copy(cm3_checkers_t h, const cm3_checkers_state *dev, const cm3_checkers_state *host,
                         bool to_host, void *stream) {
    if (!h || !dev || !host || !dev->remaining || !dev->agents || !dev->meta) {
        set_error("handle/state pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = h->cfg.num_envs, N = h->cfg.n_agents;
    struct { void *d; void *hst; size_t bytes; } f[] = {
        {dev->remaining, host->remaining, B * 8}, {dev->agents, host->agents, B * N * 4}, {dev->meta, host->meta, B * 4}};
    for (auto &c : f) {
        if (!c.hst) continue;
        if (to_host) CM3_CUDA(cudaMemcpyAsync(c.hst, c.d, c.bytes, cudaMemcpyDeviceToHost, s));
        else CM3_CUDA(cudaMemcpyAsync(c.d, c.hst, c.bytes, cudaMemcpyHostToDevice, s));
    }
    if (to_host) CM3_CUDA(cudaStreamSynchronize(s));
    return CM3_OK;
}

int cm3_checkers_get_state(cm3_checkers_t h, const cm3_checkers_state *dev, const cm3_checkers_state *host, void *stream) {
    return ck_state_copy(h, dev, host, true, stream);
}
int cm3_checkers_set_state(cm3_checkers_t h, const cm3_checkers_state *dev, const cm3_checkers_state *host, void *stream) {
    return ck_state_copy(h, dev, host, false, stream);
}

/* ------------------------------------------------------------------ Particle */

void cm3_particle_default_config(cm3_particle_config *cfg, int32_t n_agents, int32_t max_steps) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->n_agents = n_agents;
    cfg->max_steps = max_steps;
    cfg->num_envs = 1;
    cfg->real = CM3_REAL_F32;
    cfg->dt = 0.1;              /* core.py:94 */
    cfg->damping = 0.25;        /* core.py:96 */
    cfg->contact_force = 1e+2;  /* core.py:98 */
    cfg->contact_margin = 1e-3; /* core.py:99 */
    cfg->agent_size = 0.15;     /* multi-goal_spread.py:47 */
    cfg->mass = 1.0;            /* core.py:47-51 */
    cfg->sensitivity = 5.0;     /* environment.py:211 */
    cfg->reach_thresh = 0.05;   /* multi-goal_spread.py:126 */
    cfg->contact_cutoff = 0.0;  /* exact-zero criterion; see the header */
}

int cm3_particle_create(const cm3_particle_config *cfg, cm3_particle_t *out) {
    if (!cfg || !out) { set_error("cfg/out is NULL"); return CM3_ERR_BAD_ARG; }
    *out = nullptr;
    if (cfg->n_agents < 1 || cfg->num_envs < 1 || cfg->max_steps < 0 ||
        (cfg->real != CM3_REAL_F32 && cfg->real != CM3_REAL_F64)) {
        set_error("bad n_agents/num_envs/max_steps/real");
        return CM3_ERR_BAD_ARG;
    }
    if (!(cfg->contact_cutoff >= 0.0)) { set_error("contact_cutoff must be >= 0"); return CM3_ERR_BAD_ARG; }
    if (cfg->n_agents > CM3_MAX_AGENTS) {
        set_error("n_agents=%d: the particle kernels keep 1..%d agents per env in registers", cfg->n_agents, CM3_MAX_AGENTS);
        return CM3_ERR_UNSUPPORTED;
    }
    int st = require_device(cfg->device);
    if (st != CM3_OK) return st;
    cm3_particle_s *h = new (std::nothrow) cm3_particle_s();
    if (!h) { set_error("out of host memory"); return CM3_ERR_BAD_ARG; }
    h->cfg = *cfg;
    PtParams &p = h->base;
    memset(&p, 0, sizeof(p));
    p.B = cfg->num_envs; p.max_steps = cfg->max_steps; p.env_id_offset = cfg->env_id_offset;
    p.dt = cfg->dt; p.damping = cfg->damping; p.contact_force = cfg->contact_force;
    p.contact_margin = cfg->contact_margin;
    p.dist_min = cfg->agent_size + cfg->agent_size;  // core.py:189 / multi-goal_spread.py:117
    p.dist_min2 = p.dist_min * p.dist_min;
    p.mass = cfg->mass; p.sensitivity = cfg->sensitivity; p.reach_thresh = cfg->reach_thresh;
    for (int i = 0; i < CM3_MAX_AGENTS; ++i) {
        p.agents_x[i] = cfg->agents_x[i]; p.agents_y[i] = cfg->agents_y[i];
        p.landmarks_x[i] = cfg->landmarks_x[i]; p.landmarks_y[i] = cfg->landmarks_y[i];
    }
    p.initial_std = cfg->initial_std; p.prob_random = cfg->prob_random;
    // Exact shortcuts of the step kernel (particle.cu): with x = -(dist - dist_min)/contact_margin
    // (core.py:191), expf(x) == 0 for x <= -110 and exp(x) == 0 for x <= -760, so the contact
    // force of a pair further apart than dist_min + 110 k (760 k) is exactly +-0.  2 % headroom
    // covers the rounding of the squared distance.  Degenerate constants disable the shortcut.
    const double k = cfg->contact_margin;
    const bool sane = k > 0.0 && std::isfinite(k) && p.dist_min >= 0.0 && std::isfinite(p.dist_min) &&
                      std::isfinite(cfg->contact_force);
    const double inf = HUGE_VAL;
    double far_f32 = (p.dist_min + 110.0 * k) * 1.02;
    // contact_cutoff: force <= contact_force * k * exp(x) < cutoff  <=>  dist > dist_min - k ln(cutoff / (contact_force k))
    if (sane && cfg->contact_cutoff > 0.0 && cfg->contact_force > 0.0) {
        const double lg = std::log(cfg->contact_cutoff / (cfg->contact_force * k));
        if (lg < 0.0) far_f32 = std::min(far_f32, (p.dist_min - k * lg) * 1.0001);
    }
    const double far2_f32 = sane ? far_f32 * far_f32 : inf;
    const double far2_f64 = sane ? std::pow((p.dist_min + 760.0 * k) * 1.02, 2) : inf;
    const double near2 = (p.dist_min >= 0.0 && std::isfinite(p.dist_min)) ? std::pow(p.dist_min * 1.01, 2) + 1e-30 : inf;
    p.kd = PtConsts<double>{p.dt, 1.0 - p.damping, p.contact_force, p.contact_margin, p.dist_min, p.mass,
                            p.sensitivity, -p.reach_thresh, far2_f64, near2};
    p.kf = PtConsts<float>{(float)p.dt, (float)(1.0 - p.damping), (float)p.contact_force, (float)p.contact_margin,
                           (float)p.dist_min, (float)p.mass, (float)p.sensitivity, (float)(-p.reach_thresh),
                           (float)far2_f32, (float)near2};
    *out = h;
    return CM3_OK;
}

int cm3_particle_destroy(cm3_particle_t h) {
    if (!h) { set_error("handle is NULL"); return CM3_ERR_BAD_ARG; }
    if (h->pipe.ready) { DeviceGuard g(h->cfg.device); h->pipe.destroy(); }
    delete h;
    return CM3_OK;
}

int cm3_particle_tiles(cm3_particle_t h, int32_t *tiles) {
    if (!h || !tiles) { set_error("handle/tiles is NULL"); return CM3_ERR_BAD_ARG; }
    const int ew = particle_tile_envs(h->cfg.n_agents);
    *tiles = (h->cfg.num_envs + ew - 1) / ew;
    return CM3_OK;
}

static int pt_fill(cm3_particle_t h, const cm3_particle_state *st, const cm3_particle_outputs *outs,
                   PtParams &p) {
    if (!h || !st || !st->sv || !st->landmarks || !st->steps || !st->collisions || !st->reached) {
        set_error("handle/state pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    p = h->base;
    p.sv = (char *)st->sv; p.landmarks = (char *)st->landmarks;
    p.steps = st->steps; p.collisions = st->collisions; p.reached = st->reached; p.sync = st->sync;
    p.n_dst = 1; p.out_B = p.B; p.out_env0 = 0;
    if (outs) p.out[0] = pt_out(*outs);
    return CM3_OK;
}

static int pt_launch(cm3_particle_t h, const PtParams &p, void *stream) {
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    return particle_launch(h->cfg.n_agents, h->cfg.real, p, (cudaStream_t)stream);
}

int cm3_particle_reset(cm3_particle_t h, const cm3_particle_state *st, const void *init_pos,
                       const void *init_landmarks, const uint8_t *env_mask, uint64_t seed,
                       int64_t reset_counter, const cm3_particle_outputs *outs, void *stream) {
    PtParams p;
    int rc = pt_fill(h, st, outs, p);
    if (rc != CM3_OK) return rc;
    if ((init_pos == nullptr) != (init_landmarks == nullptr)) {
        set_error("init_pos and init_landmarks must be given together");
        return CM3_ERR_BAD_ARG;
    }
    p.mode = 1; p.T = 1;
    p.init_pos = (const char *)init_pos; p.init_landmarks = (const char *)init_landmarks;
    p.env_mask = env_mask; p.seed = seed; p.reset_counter = reset_counter;
    p.out[0].reward = nullptr; p.out[0].reward_n = nullptr; p.out[0].collisions = nullptr; p.out[0].reached = nullptr;
    return pt_launch(h, p, stream);
}

int cm3_particle_rollout(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                         uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset, int8_t *actions_out,
                         const cm3_particle_outputs *outs, void *stream) {
    PtParams p;
    int rc = pt_fill(h, st, outs, p);
    if (rc != CM3_OK) return rc;
    if (T < 1) { set_error("T must be >= 1"); return CM3_ERR_BAD_ARG; }
    if ((rc = check_actions_alignment(actions, h->cfg.n_agents)) != CM3_OK) return rc;
    p.mode = 0; p.T = T; p.auto_reset = auto_reset ? 1 : 0;
    p.actions = actions; p.actions_out = actions_out; p.seed = seed; p.t0 = t0;
    return pt_launch(h, p, stream);
}

int cm3_particle_rollout_gather(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                                uint64_t seed, int64_t t0, int32_t T, int32_t auto_reset, int8_t *actions_out,
                                int32_t n_dst, const cm3_particle_outputs *dsts, int64_t dst_B,
                                int64_t dst_env0, void *stream) {
    PtParams p;
    int rc = pt_fill(h, st, nullptr, p);
    if (rc != CM3_OK) return rc;
    if (T < 1) { set_error("T must be >= 1"); return CM3_ERR_BAD_ARG; }
    rc = fill_destinations(n_dst, dsts, dst_B, dst_env0, p.B, p.out, p.n_dst, p.out_B, p.out_env0, pt_out);
    if (rc != CM3_OK) return rc;
    if ((rc = check_actions_alignment(actions, h->cfg.n_agents)) != CM3_OK) return rc;
    p.mode = 0; p.T = T; p.auto_reset = auto_reset ? 1 : 0;
    p.actions = actions; p.actions_out = actions_out; p.seed = seed; p.t0 = t0;
    return pt_launch(h, p, stream);
}

int cm3_particle_step(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                      const cm3_particle_outputs *outs, void *stream) {
    if (!actions) { set_error("actions is NULL"); return CM3_ERR_BAD_ARG; }
    return cm3_particle_rollout(h, st, actions, 0, 0, 1, 0, nullptr, outs, stream);
}

int cm3_particle_step_chained(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions,
                              uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_particle_outputs *outs, void *stream) {
    if (!actions) { set_error("actions is NULL"); return CM3_ERR_BAD_ARG; }
    PtParams p;
    int rc = pt_fill(h, st, outs, p);
    if (rc != CM3_OK) return rc;
    if (!st->sync) { set_error("step_chained needs state.sync (the per-tile chaining words)"); return CM3_ERR_BAD_ARG; }
    if ((rc = check_actions_alignment(actions, h->cfg.n_agents)) != CM3_OK) return rc;
    p.mode = 0; p.T = 1; p.auto_reset = auto_reset ? 1 : 0; p.chained = 1;
    p.early = -1;  // resolved per launch shape (params.cuh: chain_early_mode)
    p.actions = actions; p.seed = seed; p.t0 = t0;
    return pt_launch(h, p, stream);
}

static int pt_step_host(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions_host,
                        int8_t *actions_dev, const cm3_particle_outputs *od, const cm3_particle_outputs *oh,
                        const void *dev_block, void *host_block, size_t block_bytes, void *stream) {
    if (!h || !actions_host || !actions_dev || !od || (!oh && !host_block) || (host_block && !dev_block)) {
        set_error("handle/actions/outputs pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = h->cfg.num_envs, N = h->cfg.n_agents, rs = real_size(h->cfg.real);
    const size_t LO = 4 * (N > 1 ? N - 1 : 1);
    static const cm3_particle_outputs none = {};
    if (!oh) oh = &none;
    const FieldCopy cp[] = {
        {oh->global_state, od->global_state, B * N * 4 * rs},
        {oh->obs_others, od->obs_others, B * N * LO * rs},
        {oh->obs_self, od->obs_self, B * N * 4 * rs},
        {oh->reward, od->reward, B * rs},
        {oh->reward_n, od->reward_n, B * N * rs},
        {oh->done, od->done, B},
        {oh->collisions, od->collisions, B * 4},
        {oh->reached, od->reached, B},
    };
    const int n = (int)(sizeof(cp) / sizeof(cp[0]));
    int rc;
    if (host_block && (rc = check_fields_in_block(cp, n, dev_block, block_bytes)) != CM3_OK) return rc;
    CM3_CUDA(cudaMemcpyAsync(actions_dev, actions_host, B * N, cudaMemcpyHostToDevice, s));
    if ((rc = cm3_particle_step(h, st, actions_dev, od, stream)) != CM3_OK) return rc;
    if (host_block) CM3_CUDA(cudaMemcpyAsync(host_block, dev_block, block_bytes, cudaMemcpyDeviceToHost, s));
    else if ((rc = copy_fields_to_host(cp, n, s)) != CM3_OK) return rc;
    CM3_CUDA(cudaStreamSynchronize(s));
    return CM3_OK;
}

int cm3_particle_step_host(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions_host,
                           int8_t *actions_dev, const cm3_particle_outputs *od,
                           const cm3_particle_outputs *oh, void *stream) {
    if (!oh) { set_error("handle/actions/outputs pointer is NULL"); return CM3_ERR_BAD_ARG; }
    return pt_step_host(h, st, actions_host, actions_dev, od, oh, nullptr, nullptr, 0, stream);
}

int cm3_particle_step_host_packed(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions_host,
                                  int8_t *actions_dev, const cm3_particle_outputs *od, const void *dev_block,
                                  void *host_block, size_t block_bytes, void *stream) {
    if (!dev_block || !host_block) { set_error("dev_block/host_block is NULL"); return CM3_ERR_BAD_ARG; }
    return pt_step_host(h, st, actions_host, actions_dev, od, nullptr, dev_block, host_block, block_bytes, stream);
}

int cm3_particle_rollout_host(cm3_particle_t h, const cm3_particle_state *st, const int8_t *actions_host,
                              int8_t *actions_dev, int32_t T, uint64_t seed, int64_t t0, int32_t auto_reset,
                              const cm3_particle_outputs *outs_dev, const void *const *dev_blocks,
                              void *host_blocks, size_t block_bytes, size_t host_stride, void *stream) {
    if (!h || !actions_host || !actions_dev || !outs_dev || !dev_blocks || !dev_blocks[0] || !dev_blocks[1] || !host_blocks) {
        set_error("handle/actions/outputs/blocks pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    if (T < 1 || host_stride < block_bytes) { set_error("T must be >= 1 and host_stride >= block_bytes"); return CM3_ERR_BAD_ARG; }
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    const size_t act_bytes = (size_t)h->cfg.num_envs * h->cfg.n_agents;
    return rollout_host_loop(h->pipe, actions_host, actions_dev, act_bytes, T, dev_blocks, host_blocks, block_bytes,
                             host_stride, (cudaStream_t)stream, [&](int t, int slot) {
                                 return cm3_particle_rollout(h, st, actions_dev + (size_t)slot * act_bytes, seed, t0 + t, 1,
                                                             auto_reset, nullptr, &outs_dev[slot], stream);
                             });
}

static int pt_state_copy(cm3_particle_t h, const cm3_particle_state *dev, const cm3_particle_state *host,
                         bool to_host, void *stream) {
    if (!h || !dev || !host || !dev->sv || !dev->landmarks || !dev->steps || !dev->collisions || !dev->reached) {
        set_error("handle/state pointer is NULL");
        return CM3_ERR_BAD_ARG;
    }
    DeviceGuard g(h->cfg.device);
    if (!g.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t B = h->cfg.num_envs, N = h->cfg.n_agents, rs = real_size(h->cfg.real);
    struct { void *d; void *hst; size_t bytes; } f[] = {
        {dev->sv, host->sv, B * N * 4 * rs}, {dev->landmarks, host->landmarks, B * N * 2 * rs},
        {dev->steps, host->steps, B * 4}, {dev->collisions, host->collisions, B * 4}, {dev->reached, host->reached, B}};
    for (auto &c : f) {
        if (!c.hst) continue;
        if (to_host) CM3_CUDA(cudaMemcpyAsync(c.hst, c.d, c.bytes, cudaMemcpyDeviceToHost, s));
        else CM3_CUDA(cudaMemcpyAsync(c.d, c.hst, c.bytes, cudaMemcpyHostToDevice, s));
    }
    if (to_host) CM3_CUDA(cudaStreamSynchronize(s));
    return CM3_OK;
}

int cm3_particle_get_state(cm3_particle_t h, const cm3_particle_state *dev, const cm3_particle_state *host, void *stream) {
    return pt_state_copy(h, dev, host, true, stream);
}
int cm3_particle_set_state(cm3_particle_t h, const cm3_particle_state *dev, const cm3_particle_state *host, void *stream) {
    return pt_state_copy(h, dev, host, false, stream);
}

}  // extern "C"
