// NCCL plumbing of the z-slab decomposition.  libnccl is opened with dlopen at the first
// flip_set_slab / flip_get_nccl_unique_id call, so single-GPU users do not need it and a process
// that already holds an NCCL (torch's bundled copy) shares that one.
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <string>
#include "flip_internal.h"

namespace flip {

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_api;

void load_api() {
    if (g_api.lib) return;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw ApiError(FLIP_ERR_UNSUPPORTED, std::string("cannot load libnccl: ") + dlerror());
    auto sym = [&](const char *n) {
        void *p = dlsym(h, n);
        if (!p) throw ApiError(FLIP_ERR_UNSUPPORTED, std::string("libnccl lacks ") + n);
        return p;
    };
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
    g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
    g_api.AllGather = (decltype(g_api.AllGather))sym("ncclAllGather");
    g_api.Send = (decltype(g_api.Send))sym("ncclSend");
    g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
    g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
    g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
    g_api.lib = h;
}

void check(ncclResult_t r, const char *what) {
    if (r != ncclSuccess) throw CudaError(std::string(what) + ": " + g_api.GetErrorString(r));
}
}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};

int comm_unique_id(void *out, int idBytes) {
    load_api();
    if (idBytes < (int)sizeof(ncclUniqueId)) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "unique id buffer too small (128 bytes)");
    ncclUniqueId id;
    check(g_api.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out, &id, sizeof(id));
    return (int)sizeof(id);
}

Comm *comm_create(int rank, int nranks, const void *uniqueId, int idBytes) {
    load_api();
    if (idBytes < (int)sizeof(ncclUniqueId)) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "unique id too short (128 bytes)");
    ncclUniqueId id;
    memcpy(&id, uniqueId, sizeof(id));
    Comm *c = new Comm();
    c->rank = rank;
    c->nranks = nranks;
    check(g_api.CommInitRank(&c->comm, nranks, id, rank), "ncclCommInitRank");
    return c;
}

void comm_destroy(Comm *c) {
    if (!c) return;
    if (c->comm) g_api.CommDestroy(c->comm);
    delete c;
}

void comm_group_begin(Comm *) { check(g_api.GroupStart(), "ncclGroupStart"); }
void comm_group_end(Comm *) { check(g_api.GroupEnd(), "ncclGroupEnd"); }

void comm_send(Comm *c, const void *buf, size_t bytes, int peer, cudaStream_t st) {
    if (bytes == 0) return;
    check(g_api.Send(buf, bytes, ncclChar, peer, c->comm, st), "ncclSend");
}
void comm_recv(Comm *c, void *buf, size_t bytes, int peer, cudaStream_t st) {
    if (bytes == 0) return;
    check(g_api.Recv(buf, bytes, ncclChar, peer, c->comm, st), "ncclRecv");
}

void comm_allgather_f32(Comm *c, const float *send, float *recv, size_t countPerRank, cudaStream_t st) {
    check(g_api.AllGather(send, recv, countPerRank, ncclFloat32, c->comm, st), "ncclAllGather");
}

void comm_allreduce(Comm *c, void *buf, size_t count, int kind, cudaStream_t st) {
    ncclDataType_t dt = ncclFloat64;
    ncclRedOp_t op = ncclSum;
    switch (kind) {
        case COMM_SUM_F64: dt = ncclFloat64; op = ncclSum; break;
        case COMM_MAX_U64: dt = ncclUint64; op = ncclMax; break;
        case COMM_SUM_I32: dt = ncclInt32; op = ncclSum; break;
        case COMM_MAX_U32: dt = ncclUint32; op = ncclMax; break;
    }
    check(g_api.AllReduce(buf, buf, count, dt, op, c->comm, st), "ncclAllReduce");
}

}  // namespace flip
