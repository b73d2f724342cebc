// Peer-memory exchanges between the z-slab ranks of one NVSwitch box: halo planes and the PCG
// scalars go straight into the neighbour's HBM over NVLink from a kernel on the solver's stream, instead of an
// NCCL send/recv group (~30-45 us each; a multigrid-preconditioned PCG iteration needs ~16 of them, which
// made the iteration latency-bound at 2/4/8 GPUs).  One process per GPU: every rank exports an "inbox"
// allocation with cudaIpcGetMemHandle, the handles travel once through an NCCL all-gather, and every rank maps
// the other inboxes with cudaIpcOpenMemHandle.
//
//   plane exchange  (one kernel per rank):  copy my boundary plane(s) into the neighbours' inbox slot
//                   -> __threadfence_system -> the last CTA publishes the epoch in the neighbour's flag
//                   -> spin on my own flag(s) until the neighbours' epoch arrives -> copy inbox -> my halo plane
//   scalar all-reduce (one CTA):            store my value into every rank's inbox -> publish -> wait for all
//                   ranks -> reduce in rank order (the same bits on every rank)
//
// Inbox slots alternate with the epoch parity.  Slot reuse is safe: before a rank writes epoch e+2 into a
// neighbour's slot it has completed its own exchange e+1, which waited for the neighbour's epoch e+1 flag, and
// the neighbour published that only after its exchange e kernel (which emptied the slot) had finished.
// Every rank issues the same sequence of exchanges (the PCG control flow depends on all-reduced scalars only).
//
// FLIP_PEER=0 in the environment keeps everything on NCCL (also the fallback when IPC mapping fails).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "flip_internal.h"

namespace flip {

static constexpr int PEER_FLAG_BYTES = 4096;
static constexpr int PEER_SCALAR_BYTES = 2 * 64 * 8 * (int)sizeof(double);   // [slot][rank<=64][8]
static constexpr int PEER_MAX_RANKS = 64;

// flags (u64 each), inside the first PEER_FLAG_BYTES of an inbox
//   [0] planes from the lower neighbour   [1] planes from the upper neighbour   [8 + r] scalars from rank r
// then the scalar area, then plane data: [side 0|1][slot 0|1][slotBytes]
struct PeerState {
    bool on = false;
    int rank = 0, nranks = 1;
    size_t slotBytes = 0, inboxBytes = 0;
    unsigned char *inbox = nullptr;                 // mine
    std::vector<unsigned char *> peer;              // every rank's inbox as mapped here (peer[rank] == inbox)
    unsigned char **peerDev = nullptr;              // the same table on the device
    unsigned long long planeEpoch = 0, scalarEpoch = 0;
    unsigned int *pushCounter = nullptr;            // CTA counter of the exchange kernel
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <class T>
struct PlaneXchg {
    const T *srcLo, *srcHi;            // my boundary planes (null: no neighbour on that side)
    T *outLo, *outHi;                  // slot in the lower / upper neighbour's inbox
    unsigned long long *flagLo, *flagHi;   // their flags
    const T *inLo, *inHi;              // my inbox slots
    T *dstLo, *dstHi;                  // my halo planes
    const unsigned long long *myFlagLo, *myFlagHi;
    unsigned long long epoch;
    size_t n;                          // words of T per plane
    unsigned int *counter;
};

template <class T>
__global__ void __launch_bounds__(256) k_peer_plane_exchange(PlaneXchg<T> x) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x.srcLo) for (size_t t = t0; t < x.n; t += stride) x.outLo[t] = x.srcLo[t];
    if (x.srcHi) for (size_t t = t0; t < x.n; t += stride) x.outHi[t] = x.srcHi[t];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(x.counter, 1u);
        if (done == gridDim.x - 1) {
            *x.counter = 0u;
            __threadfence_system();
            if (x.srcLo) st_release_sys(x.flagLo, x.epoch);
            if (x.srcHi) st_release_sys(x.flagHi, x.epoch);
        }
        if (x.dstLo) while (ld_acquire_sys(x.myFlagLo) < x.epoch) __nanosleep(20);
        if (x.dstHi) while (ld_acquire_sys(x.myFlagHi) < x.epoch) __nanosleep(20);
    }
    __syncthreads();
    if (x.dstLo) for (size_t t = t0; t < x.n; t += stride) x.dstLo[t] = __ldcv(x.inLo + t);
    if (x.dstHi) for (size_t t = t0; t < x.n; t += stride) x.dstHi[t] = __ldcv(x.inHi + t);
}

// kind: COMM_SUM_F64 or COMM_MAX_U64; count <= 8 values at `val` (device), reduced in place
__global__ void __launch_bounds__(64) k_peer_allreduce(unsigned char **peer, int rank, int nranks, unsigned long long epoch,
                                                       double *val, int count, int kind) {
    const int r = threadIdx.x;
    const int slot = (int)(epoch & 1ull);
    if (r < nranks) {
        double *dst = reinterpret_cast<double *>(peer[r] + PEER_FLAG_BYTES) + ((size_t)slot * PEER_MAX_RANKS + rank) * 8;
        for (int q = 0; q < count; q++) dst[q] = val[q];
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long *>(peer[r]) + 8 + rank, epoch);
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(peer[rank]) + 8 + r;
        while (ld_acquire_sys(mine) < epoch) __nanosleep(20);
    }
    __syncthreads();
    if (r < count) {
        const volatile double *in = reinterpret_cast<const volatile double *>(peer[rank] + PEER_FLAG_BYTES) +
                                    (size_t)slot * PEER_MAX_RANKS * 8;
        if (kind == COMM_SUM_F64) {
            double s = 0.0;
            for (int q = 0; q < nranks; q++) s += in[q * 8 + r];
            val[r] = s;
        } else {
            const volatile unsigned long long *inu = reinterpret_cast<const volatile unsigned long long *>(in);
            unsigned long long m = 0ull;
            for (int q = 0; q < nranks; q++) { unsigned long long v = inu[q * 8 + r]; m = v > m ? v : m; }
            reinterpret_cast<unsigned long long *>(val)[r] = m;
        }
    }
}

static PeerState *state(flip_ctx *c) { return (PeerState *)c->peer; }

bool peer_on(const flip_ctx *c) { return c->peer && ((PeerState *)c->peer)->on; }

// After the grids exist (the slot size follows the plane size).  Collective: every rank calls it.
void peer_setup(flip_ctx *c) {
    if (c->peer || !slab_on(c)) return;
    PeerState *P = new PeerState();
    c->peer = P;
    P->rank = c->rank; P->nranks = c->nranks;
    const char *env = getenv("FLIP_PEER");
    int want = !(env && env[0] == '0') && c->nranks <= PEER_MAX_RANKS;
    cudaStream_t st = c->stream;
    P->slotBytes = (((size_t)c->d.I * c->d.J * sizeof(double)) + 255) & ~(size_t)255;
    P->inboxBytes = PEER_FLAG_BYTES + PEER_SCALAR_BYTES + 4 * P->slotBytes;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (want) {
        if (cudaMalloc(&P->inbox, P->inboxBytes) != cudaSuccess || cudaMemset(P->inbox, 0, P->inboxBytes) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine, P->inbox) != cudaSuccess) {
            cudaGetLastError();
            want = 0;
        }
    }
    // handles (and whether every rank could export one) through NCCL: [ok flag | handle] per rank
    const size_t rec = 128;
    static_assert(sizeof(cudaIpcMemHandle_t) <= 96, "handle record");
    std::vector<unsigned char> host(rec * c->nranks, 0);
    unsigned char *dev = nullptr;
    FLIP_CUDA_CHECK(cudaMalloc(&dev, rec * (c->nranks + 1)));
    unsigned char my[rec];
    memset(my, 0, rec);
    my[0] = (unsigned char)want;
    memcpy(my + 16, &mine, sizeof(mine));
    FLIP_CUDA_CHECK(cudaMemcpyAsync(dev + rec * c->nranks, my, rec, cudaMemcpyHostToDevice, st));
    comm_allgather_f32(c->comm, reinterpret_cast<const float *>(dev + rec * c->nranks), reinterpret_cast<float *>(dev),
                       rec / sizeof(float), st);
    FLIP_CUDA_CHECK(cudaMemcpyAsync(host.data(), dev, rec * c->nranks, cudaMemcpyDeviceToHost, st));
    FLIP_CUDA_CHECK(cudaStreamSynchronize(st));
    int all = 1;
    for (int r = 0; r < c->nranks; r++) all &= host[rec * r];
    P->peer.assign(c->nranks, nullptr);
    if (all) {
        for (int r = 0; r < c->nranks && all; r++) {
            if (r == c->rank) { P->peer[r] = P->inbox; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, host.data() + rec * r + 16, sizeof(h));
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                all = 0;
            }
            P->peer[r] = (unsigned char *)p;
        }
    }
    // agree on the outcome: a rank that could not map a peer switches everybody back to NCCL
    int *flag = reinterpret_cast<int *>(dev);
    FLIP_CUDA_CHECK(cudaMemcpyAsync(flag, &all, sizeof(int), cudaMemcpyHostToDevice, st));
    // min over ranks == 1  <=>  sum == nranks
    comm_allreduce(c->comm, flag, 1, COMM_SUM_I32, st);
    int sum = 0;
    FLIP_CUDA_CHECK(cudaMemcpyAsync(&sum, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    FLIP_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dev);
    if (sum == c->nranks) {
        FLIP_CUDA_CHECK(cudaMalloc(&P->peerDev, sizeof(unsigned char *) * c->nranks));
        FLIP_CUDA_CHECK(cudaMemcpy(P->peerDev, P->peer.data(), sizeof(unsigned char *) * c->nranks, cudaMemcpyHostToDevice));
        FLIP_CUDA_CHECK(cudaMalloc(&P->pushCounter, sizeof(unsigned int)));
        FLIP_CUDA_CHECK(cudaMemset(P->pushCounter, 0, sizeof(unsigned int)));
        P->on = true;
    } else {
        if (c->rank == 0 && !(env && env[0] == '0'))
            fprintf(stderr, "flip_b200: peer-memory exchange unavailable (CUDA IPC), z-slab exchanges stay on NCCL\n");
        for (int r = 0; r < c->nranks; r++)
            if (r != c->rank && P->peer[r]) cudaIpcCloseMemHandle(P->peer[r]);
        P->peer.assign(c->nranks, nullptr);
    }
}

void peer_free(flip_ctx *c) {
    PeerState *P = state(c);
    if (!P) return;
    for (int r = 0; r < (int)P->peer.size(); r++)
        if (r != P->rank && P->peer[r]) cudaIpcCloseMemHandle(P->peer[r]);
    cudaFree(P->peerDev);
    cudaFree(P->pushCounter);
    cudaFree(P->inbox);
    delete P;
    c->peer = nullptr;
}

template <class T>
static void launch_plane_exchange(flip_ctx *c, PeerState *P, const void *sendLo, void *recvLo, const void *sendHi, void *recvHi,
                                  size_t bytes, unsigned long long e) {
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    const size_t slot = (size_t)(e & 1ull);
    auto data = [&](unsigned char *inbox, int side) {
        return inbox + PEER_FLAG_BYTES + PEER_SCALAR_BYTES + ((size_t)side * 2 + slot) * P->slotBytes;
    };
    PlaneXchg<T> x;
    memset(&x, 0, sizeof(x));
    x.epoch = e; x.n = bytes / sizeof(T); x.counter = P->pushCounter;
    if (hasLo) {
        unsigned char *nb = P->peer[c->rank - 1];
        x.srcLo = (const T *)sendLo; x.outLo = (T *)data(nb, 1);           // I am its upper neighbour
        x.flagLo = reinterpret_cast<unsigned long long *>(nb) + 1;
        x.inLo = (const T *)data(P->inbox, 0); x.dstLo = (T *)recvLo;
        x.myFlagLo = reinterpret_cast<const unsigned long long *>(P->inbox) + 0;
    }
    if (hasHi) {
        unsigned char *nb = P->peer[c->rank + 1];
        x.srcHi = (const T *)sendHi; x.outHi = (T *)data(nb, 0);           // I am its lower neighbour
        x.flagHi = reinterpret_cast<unsigned long long *>(nb) + 0;
        x.inHi = (const T *)data(P->inbox, 1); x.dstHi = (T *)recvHi;
        x.myFlagHi = reinterpret_cast<const unsigned long long *>(P->inbox) + 1;
    }
    int blocks = (int)((bytes + 32767) / 32768);
    blocks = blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks);
    k_peer_plane_exchange<T><<<blocks, 256, 0, c->stream>>>(x);
    c->launches++;
}

// One plane of `bytes` bytes each way: sendLo -> lower neighbour (arrives there as its "from upper" plane),
// sendHi -> upper neighbour; recvLo / recvHi are my halo planes.  Both sides of a pair pass the same size.
void peer_exchange_planes(flip_ctx *c, const void *sendLo, void *recvLo, const void *sendHi, void *recvHi, size_t bytes) {
    PeerState *P = state(c);
    if (bytes > P->slotBytes || (bytes & 3)) throw CudaError("peer plane exchange: plane size");
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    if (!hasLo && !hasHi) return;
    const unsigned long long e = ++P->planeEpoch;
    // the inbox slots are 256-byte aligned; my own planes decide the copy width (a local matter)
    uintptr_t al = bytes;
    if (hasLo) al |= (uintptr_t)sendLo | (uintptr_t)recvLo;
    if (hasHi) al |= (uintptr_t)sendHi | (uintptr_t)recvHi;
    if ((al & 15) == 0) launch_plane_exchange<uint4>(c, P, sendLo, recvLo, sendHi, recvHi, bytes, e);
    else launch_plane_exchange<unsigned int>(c, P, sendLo, recvLo, sendHi, recvHi, bytes, e);
}

// `count` (<= 8) fp64 sums or u64 maxima at the device address `val`, over all ranks, in place
void peer_allreduce(flip_ctx *c, void *val, int count, int kind) {
    PeerState *P = state(c);
    const unsigned long long e = ++P->scalarEpoch;
    k_peer_allreduce<<<1, 64, 0, c->stream>>>(P->peerDev, c->rank, c->nranks, e, (double *)val, count, kind);
    c->launches++;
}

}  // namespace flip
