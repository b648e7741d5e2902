// comm.cu - cm3_comm_*: the rollout exchange for bindings WITHOUT torch.distributed (SURVEY.md §8b
// sketched cm3_comm_init / cm3_allgather).  The Python facades use torch.distributed for rendezvous
// and collectives (cm3_b200/sharding.py); a cgo / JNI / plain-C binding has neither, so the same
// NCCL all-gather is offered behind the C ABI: rank 0 makes a 128-byte unique id, the binding carries
// it to the other ranks by whatever means it has (file, socket, MPI), every rank calls cm3_comm_init.
// NCCL is resolved at run time (dlopen of libnccl.so.2 - the copy PyTorch already loaded when there
// is one, else CM3_NCCL_LIBRARY or the loader path), so libcm3env.so itself has no NCCL dependency.
// The reference has no counterpart: its only fan-out is one process per seed without communication
// (alg/train_multiprocess.py:31-43).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "common.cuh"

namespace {

struct NcclId { char internal[128]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void *NcclComm;
typedef int (*GetUniqueIdFn)(NcclId *);
typedef int (*CommInitRankFn)(NcclComm *, int, NcclId, int);
typedef int (*CommDestroyFn)(NcclComm);
typedef int (*AllGatherFn)(const void *, void *, size_t, int /* ncclDataType_t */, NcclComm, cudaStream_t);
typedef const char *(*GetErrorStringFn)(int);

struct Nccl {
    void *so = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllGatherFn all_gather = nullptr;
    GetErrorStringFn error_string = nullptr;
    bool ok = false;
};

Nccl &nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *cands[] = {getenv("CM3_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
        for (const char *c : cands) {
            if (!c || !c[0]) continue;
            n.so = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
            if (n.so) break;
        }
        if (!n.so) return;
        n.get_unique_id = (GetUniqueIdFn)dlsym(n.so, "ncclGetUniqueId");
        n.comm_init_rank = (CommInitRankFn)dlsym(n.so, "ncclCommInitRank");
        n.comm_destroy = (CommDestroyFn)dlsym(n.so, "ncclCommDestroy");
        n.all_gather = (AllGatherFn)dlsym(n.so, "ncclAllGather");
        n.error_string = (GetErrorStringFn)dlsym(n.so, "ncclGetErrorString");
        n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_gather && n.error_string;
    });
    return n;
}

int nccl_fail(int rc, const char *what) {
    cm3::set_error("%s: %s", what, nccl().error_string ? nccl().error_string(rc) : "NCCL error");
    return CM3_ERR_NCCL;
}

int need_nccl() {
    if (nccl().ok) return CM3_OK;
    cm3::set_error("NCCL not found (dlopen libnccl.so.2; set CM3_NCCL_LIBRARY to its path)");
    return CM3_ERR_NCCL;
}

}  // namespace

struct cm3_comm_s {
    NcclComm comm;
    int rank, world, device;
};

extern "C" {

int cm3_comm_unique_id(uint8_t *id) {
    if (!id) { cm3::set_error("id is NULL"); return CM3_ERR_BAD_ARG; }
    int rc = need_nccl();
    if (rc != CM3_OK) return rc;
    NcclId nid;
    if ((rc = nccl().get_unique_id(&nid)) != 0) return nccl_fail(rc, "ncclGetUniqueId");
    memcpy(id, nid.internal, sizeof(nid.internal));
    return CM3_OK;
}

int cm3_comm_init(const uint8_t *id, int32_t rank, int32_t world, int32_t device, cm3_comm_t *out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) {
        cm3::set_error("id/out is NULL or rank %d outside world %d", rank, world);
        return CM3_ERR_BAD_ARG;
    }
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        cm3::set_error("no CUDA device visible - this library has no CPU fallback");
        return CM3_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { cm3::set_error("device %d out of range (0..%d)", device, n - 1); return CM3_ERR_BAD_ARG; }
    int rc = need_nccl();
    if (rc != CM3_OK) return rc;
    CM3_CUDA(cudaSetDevice(device));
    cm3_comm_s *c = new (std::nothrow) cm3_comm_s();
    if (!c) { cm3::set_error("out of host memory"); return CM3_ERR_BAD_ARG; }
    NcclId nid;
    memcpy(nid.internal, id, sizeof(nid.internal));
    if ((rc = nccl().comm_init_rank(&c->comm, world, nid, rank)) != 0) {
        delete c;
        return nccl_fail(rc, "ncclCommInitRank");
    }
    c->rank = rank; c->world = world; c->device = device;
    *out = c;
    return CM3_OK;
}

int cm3_comm_allgather(cm3_comm_t c, const void *send, void *recv, size_t bytes_per_rank, void *stream) {
    if (!c || !send || !recv) { cm3::set_error("comm/send/recv is NULL"); return CM3_ERR_BAD_ARG; }
    const int rc = nccl().all_gather(send, recv, bytes_per_rank, 1 /* ncclUint8 */, c->comm, (cudaStream_t)stream);
    return rc == 0 ? CM3_OK : nccl_fail(rc, "ncclAllGather");
}

int cm3_comm_destroy(cm3_comm_t c) {
    if (!c) { cm3::set_error("comm is NULL"); return CM3_ERR_BAD_ARG; }
    const int rc = nccl().ok ? nccl().comm_destroy(c->comm) : 0;
    delete c;
    return rc == 0 ? CM3_OK : nccl_fail(rc, "ncclCommDestroy");
}

}  // extern "C"
