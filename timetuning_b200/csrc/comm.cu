// Multi-GPU plumbing: an ncclComm_t owned by this library, bootstrapped from a unique id that
// the host side broadcasts (timetuning_b200/dist.py).  The only data-path collective of the hot
// path is the K-float all-reduce of the Sinkhorn prototype marginals
// (/root/reference/my_utils.py:259-272), enqueued on the caller's stream between two passes.
// libnccl.so.2 is dlopen'ed lazily so that the library loads (and every single-GPU path works)
// without NCCL.  Minimal NCCL prototypes are declared here instead of including nccl.h.
#include <dlfcn.h>

#include "common.cuh"

namespace timet {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
        set_error("NCCL not available: %s", api.handle ? "missing symbols in libnccl" : dlerror());
        return nullptr;
    }
    return &api;
}

struct Comm {
    ncclComm_t nccl;
    int rank, world_size;
    // peer-memory exchange (NVLink P2P through CUDA IPC) used by the resident Sinkhorn kernel
    void *p2p_local = nullptr;                 // this rank's P2PBuf (cudaMalloc'ed here, written by the peers)
    void *p2p_peers_host[P2P_MAX_RANKS] = {};   // device pointers of every rank's P2PBuf (own at [rank])
    void **p2p_peers_dev = nullptr;            // the same table in device memory
    unsigned long long p2p_epoch[2] = {0, 0};  // exchanges performed so far per channel (identical on every rank)
    bool p2p_ready = false;
};

#define TIMET_NCCL(call)                                                                          \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != 0) {                                                                           \
            set_error("%s failed: %s", #call, api->GetErrorString ? api->GetErrorString(r__) : "?"); \
            return TIMET_ERR_NCCL;                                                                \
        }                                                                                         \
    } while (0)

int comm_allreduce_f32(timet_comm_t comm, float *buf, int64_t n, cudaStream_t st) {
    NcclApi *api = nccl();
    if (!api) return TIMET_ERR_NCCL;
    TIMET_CHECK_ARG(comm != nullptr, "allreduce: NULL communicator");
    Comm *c = (Comm *)comm;
    TIMET_NCCL(api->AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, c->nccl, st));
    return TIMET_OK;
}

bool comm_p2p_info(timet_comm_t comm, void ***peers_dev, int *rank, int *ws, unsigned long long **epoch) {
    Comm *c = (Comm *)comm;
    if (!c || !c->p2p_ready) return false;
    *peers_dev = c->p2p_peers_dev; *rank = c->rank; *ws = c->world_size; *epoch = c->p2p_epoch;
    return true;
}

}  // namespace timet

using namespace timet;

extern "C" {

int timet_comm_unique_id(void *id_out) {
    TIMET_CHECK_ARG(id_out != nullptr, "comm_unique_id: NULL output");
    NcclApi *api = nccl();
    if (!api) return TIMET_ERR_NCCL;
    ncclUniqueId id;
    TIMET_NCCL(api->GetUniqueId(&id));
    memcpy(id_out, &id, TIMET_UNIQUE_ID_BYTES);
    return TIMET_OK;
}

int timet_comm_init(const void *id, int rank, int world_size, timet_comm_t *comm_out) {
    TIMET_CHECK_ARG(id && comm_out, "comm_init: NULL pointer");
    TIMET_CHECK_ARG(world_size >= 1 && rank >= 0 && rank < world_size, "comm_init: bad rank %d / %d", rank, world_size);
    NcclApi *api = nccl();
    if (!api) return TIMET_ERR_NCCL;
    ncclUniqueId uid;
    memcpy(&uid, id, TIMET_UNIQUE_ID_BYTES);
    Comm *c = new Comm{nullptr, rank, world_size};
    ncclResult_t r = api->CommInitRank(&c->nccl, world_size, uid, rank);
    if (r != 0) {
        set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
        delete c;
        return TIMET_ERR_NCCL;
    }
    *comm_out = c;
    return TIMET_OK;
}

int timet_comm_p2p_handle(timet_comm_t comm, void *handle_out) {
    TIMET_CHECK_ARG(comm && handle_out, "comm_p2p_handle: NULL pointer");
    Comm *c = (Comm *)comm;
    TIMET_CHECK_ARG(c->world_size <= P2P_MAX_RANKS, "comm_p2p: world size %d > %d", c->world_size, P2P_MAX_RANKS);
    if (!c->p2p_local) {
        TIMET_CUDA(cudaMalloc(&c->p2p_local, 2 * sizeof(P2PBuf)));          // two exchange channels
        TIMET_CUDA(cudaMemset(c->p2p_local, 0, 2 * sizeof(P2PBuf)));
        TIMET_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    TIMET_CUDA(cudaIpcGetMemHandle(&h, c->p2p_local));
    static_assert(sizeof(h) == TIMET_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &h, sizeof(h));
    return TIMET_OK;
}

int timet_comm_p2p_connect(timet_comm_t comm, const void *all_handles) {
    TIMET_CHECK_ARG(comm && all_handles, "comm_p2p_connect: NULL pointer");
    Comm *c = (Comm *)comm;
    TIMET_CHECK_ARG(c->p2p_local != nullptr, "comm_p2p_connect: call timet_comm_p2p_handle first");
    for (int r = 0; r < c->world_size; ++r) {
        if (r == c->rank) { c->p2p_peers_host[r] = c->p2p_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)all_handles + (size_t)r * TIMET_IPC_HANDLE_BYTES, sizeof(h));
        TIMET_CUDA(cudaIpcOpenMemHandle(&c->p2p_peers_host[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    TIMET_CUDA(cudaMalloc((void **)&c->p2p_peers_dev, sizeof(void *) * P2P_MAX_RANKS));
    TIMET_CUDA(cudaMemcpy(c->p2p_peers_dev, c->p2p_peers_host, sizeof(void *) * P2P_MAX_RANKS, cudaMemcpyHostToDevice));
    c->p2p_epoch[0] = c->p2p_epoch[1] = 0;
    c->p2p_ready = true;
    return TIMET_OK;
}

int timet_comm_p2p_disable(timet_comm_t comm) {
    TIMET_CHECK_ARG(comm != nullptr, "comm_p2p_disable: NULL communicator");
    ((Comm *)comm)->p2p_ready = false;          // Sinkhorn falls back to the NCCL all-reduce path
    return TIMET_OK;
}

int timet_comm_destroy(timet_comm_t comm) {
    if (!comm) return TIMET_OK;
    NcclApi *api = nccl();
    if (!api) return TIMET_ERR_NCCL;
    Comm *c = (Comm *)comm;
    if (c->p2p_ready) {
        cudaDeviceSynchronize();
        for (int r = 0; r < c->world_size; ++r)
            if (r != c->rank && c->p2p_peers_host[r]) cudaIpcCloseMemHandle(c->p2p_peers_host[r]);
        cudaFree(c->p2p_peers_dev);
    }
    if (c->p2p_local) cudaFree(c->p2p_local);
    TIMET_NCCL(api->CommDestroy(c->nccl));
    delete c;
    return TIMET_OK;
}

int timet_comm_allreduce_f32(timet_comm_t comm, float *buf, int64_t n, timet_stream_t stream) {
    TIMET_CHECK_ARG(buf != nullptr && n >= 0, "allreduce: bad buffer");
    return comm_allreduce_f32(comm, buf, n, (cudaStream_t)stream);
}

}
