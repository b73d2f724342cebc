// z-slab decomposition across the GPUs of one box.  The reference has no multi-process
// path; this is new.  One process per GPU; rank r owns the global cell planes [K0_r, K1_r) and keeps
// `halo` extra planes towards each neighbour, in which every grid stage is computed redundantly from
// ghost particles, so the only exchanges per substep are:
//   1. ghost particles (the neighbour's particles within `halo` planes of the shared face) before the
//      liquid-SDF / P2G gather;
//   2. one plane of the PCG search vector per iteration + the scalar all-reduces (pressure.cu);
//   3. the projected velocity (and valid mask) halo planes after the pressure update;
//   4. migrating particles after the RK3 advection.
// k is the slowest array index, so every range below is contiguous: halo planes of a grid, and — the
// particle store being sorted by cell — the particles of a range of planes.
#include <algorithm>
#include "device_math.cuh"
#include "flip_internal.h"

namespace flip {

static constexpr int TPB = 256;

__global__ void k_slab_ghost_counts(const int *__restrict__ cellStart, int IJ, int kOwn0, int kOwn1, int halo, int hasLo,
                                    int hasHi, DeviceScalars *S) {
    S->sendCount[0] = hasLo ? cellStart[(kOwn0 + halo) * IJ] - cellStart[kOwn0 * IJ] : 0;
    S->sendCount[1] = hasHi ? cellStart[kOwn1 * IJ] - cellStart[(kOwn1 - halo) * IJ] : 0;
    S->recvCount[0] = 0;
    S->recvCount[1] = 0;
    S->slabError[0] = S->slabError[1] = S->slabError[2] = 0;
}

static void exchange_counts(flip_ctx *c) {
    cudaStream_t st = c->stream;
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    comm_group_begin(c->comm);
    if (hasLo) {
        comm_send(c->comm, &c->dS->sendCount[0], sizeof(int), c->rank - 1, st);
        comm_recv(c->comm, &c->dS->recvCount[0], sizeof(int), c->rank - 1, st);
    }
    if (hasHi) {
        comm_send(c->comm, &c->dS->sendCount[1], sizeof(int), c->rank + 1, st);
        comm_recv(c->comm, &c->dS->recvCount[1], sizeof(int), c->rank + 1, st);
    }
    comm_group_end(c->comm);
    // an error seen by one rank (slabError != 0) must stop ALL ranks at the same point of the exchange sequence:
    // a rank that threw on its own would leave its neighbours waiting in the next send/recv group for ever
    comm_allreduce(c->comm, c->dS->slabError, 3, COMM_SUM_I32, st);
    scalars_to_host(c);
}

struct SoAPtrs {
    float *a[6];
    int *id;
};
static SoAPtrs ptrs_of(flip_ctx *c, int buf) {
    ParticleSoA &p = c->P[buf];
    SoAPtrs q;
    q.a[0] = p.px; q.a[1] = p.py; q.a[2] = p.pz; q.a[3] = p.vx; q.a[4] = p.vy; q.a[5] = p.vz;
    q.id = c->pid[buf];
    return q;
}

// The cell index of the ghost-extended store, without sorting anything: cells are ordered plane by plane (k is
// the slowest index), the neighbours' boundary layers arrive cell-sorted and lie entirely below / above the owned
// planes, so the merged order is [ghosts from below][owned][ghosts from above] and the new cellStart is the
// neighbours' slices and my own, rebased.  Also rebuilds the 4x4x4 occupancy blocks of the gather.
__global__ void k_slab_merge_cellstart(int nC, int IJ, int kOwn0, int kOwn1, const int *__restrict__ oldStart,
                                       const int *__restrict__ csLo, const int *__restrict__ csHi, int recvLo, int nOwned,
                                       int total, int *__restrict__ newStart) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nC) return;
    const int lo0 = kOwn0 * IJ, hi0 = kOwn1 * IJ;
    int v;
    if (c == nC) v = total;
    else if (c < lo0) v = csLo ? csLo[c] - csLo[0] : 0;
    else if (c < hi0) v = recvLo + (oldStart[c] - oldStart[lo0]);
    else v = recvLo + nOwned + (csHi ? csHi[c - hi0] - csHi[0] : 0);
    newStart[c] = v;
}
__global__ void k_slab_mark_occ(int nC, const int *__restrict__ start, unsigned char *__restrict__ occ, int I, int J, int oI,
                                int oJ) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    if (start[c + 1] > start[c]) {
        int i = c % I, j = (c / I) % J, k = c / (I * J);
        occ[(i >> 2) + oI * ((j >> 2) + oJ * (k >> 2))] = 1;
    }
}

void slab_exchange_ghosts(flip_ctx *c) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    const int IJ = d.I * d.J;
    // the store holds exactly the owned particles, sorted: the two boundary layers are its head and tail
    k_slab_ghost_counts<<<1, 1, 0, st>>>(c->cellStart, IJ, d.kOwn0, d.kOwn1, c->halo, hasLo ? 1 : 0, hasHi ? 1 : 0, c->dS);
    c->launches++;
    exchange_counts(c);
    const int sendLo = c->hS->sendCount[0], sendHi = c->hS->sendCount[1];
    const int recvLo = c->hS->recvCount[0], recvHi = c->hS->recvCount[1];
    const int n = c->np;
    const int total = n + recvLo + recvHi;
    if (c->ownedBegin != 0) throw CudaError("z-slab ghost exchange expects the owned particles at the head of the store");
    c->npStore = n;
    particles_alloc(c, total);
    SoAPtrs S = ptrs_of(c, c->cur_buf), D = ptrs_of(c, 1 - c->cur_buf);
    // the neighbours' cell-start slices of the exchanged planes land in the (idle) count array
    const size_t sliceInts = (size_t)c->halo * IJ + 1;
    int *csLo = hasLo ? c->cellCount : nullptr;
    int *csHi = hasHi ? c->cellCount + (hasLo ? sliceInts : 0) : nullptr;
    for (int a = 0; a < 6; a++)
        FLIP_CUDA_CHECK(cudaMemcpyAsync(D.a[a] + recvLo, S.a[a], sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    if (c->trackIds) FLIP_CUDA_CHECK(cudaMemcpyAsync(D.id + recvLo, S.id, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    comm_group_begin(c->comm);
    for (int a = 0; a < 6; a++) {
        if (hasLo) {
            comm_send(c->comm, S.a[a], sizeof(float) * (size_t)sendLo, c->rank - 1, st);
            comm_recv(c->comm, D.a[a], sizeof(float) * (size_t)recvLo, c->rank - 1, st);
        }
        if (hasHi) {
            comm_send(c->comm, S.a[a] + (n - sendHi), sizeof(float) * (size_t)sendHi, c->rank + 1, st);
            comm_recv(c->comm, D.a[a] + recvLo + n, sizeof(float) * (size_t)recvHi, c->rank + 1, st);
        }
    }
    if (c->trackIds) {
        if (hasLo) {
            comm_send(c->comm, S.id, sizeof(int) * (size_t)sendLo, c->rank - 1, st);
            comm_recv(c->comm, D.id, sizeof(int) * (size_t)recvLo, c->rank - 1, st);
        }
        if (hasHi) {
            comm_send(c->comm, S.id + (n - sendHi), sizeof(int) * (size_t)sendHi, c->rank + 1, st);
            comm_recv(c->comm, D.id + recvLo + n, sizeof(int) * (size_t)recvHi, c->rank + 1, st);
        }
    }
    if (hasLo) {
        comm_send(c->comm, c->cellStart + (size_t)d.kOwn0 * IJ, sizeof(int) * sliceInts, c->rank - 1, st);
        comm_recv(c->comm, csLo, sizeof(int) * sliceInts, c->rank - 1, st);
    }
    if (hasHi) {
        comm_send(c->comm, c->cellStart + (size_t)(d.kOwn1 - c->halo) * IJ, sizeof(int) * sliceInts, c->rank + 1, st);
        comm_recv(c->comm, csHi, sizeof(int) * sliceInts, c->rank + 1, st);
    }
    comm_group_end(c->comm);
    k_slab_merge_cellstart<<<cdiv(d.nC + 1, TPB), TPB, 0, st>>>(d.nC, IJ, d.kOwn0, d.kOwn1, c->cellStart, csLo, csHi, recvLo, n,
                                                               total, c->cellStartA);
    c->launches++;
    std::swap(c->cellStart, c->cellStartA);
    {
        const int oI = (d.I + 3) >> 2, oJ = (d.J + 3) >> 2, oK = (d.K + 3) >> 2;
        FLIP_CUDA_CHECK(cudaMemsetAsync(c->occ, 0, (size_t)oI * oJ * oK, st));
        k_slab_mark_occ<<<cdiv(d.nC, TPB), TPB, 0, st>>>(d.nC, c->cellStart, c->occ, d.I, d.J, oI, oJ);
        c->launches++;
    }
    c->occBitsValid = false;         // the one-bit occupancy map of the last sort does not know the ghosts
    c->stepCounter++;
    c->cur_buf = 1 - c->cur_buf;
    c->npStore = total;              // owned + ghosts
    c->ownedBegin = recvLo;
    c->ownedEnd = recvLo + n;
    c->np = n;
    c->ghostsPresent = true;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

// Owned particles whose new plane belongs to a neighbour are appended (unordered) to the send buffers.
// reachLo / reachHi: thickness of the lower / upper neighbour's slab -- a particle that lands beyond it would have to
// go to a rank this exchange does not talk to; that is reported (slabError[0]), never dropped silently.
__global__ void k_slab_pack_emigrants(ParticleSoA p, const int *__restrict__ ids, int n, double invdx, int kOff, int kOwn0,
                                      int kOwn1, int reachLo, int reachHi, float *lo, float *hi, int cap, DeviceScalars *S) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int kl = pos2idx(p.pz[t], invdx) - kOff;
    int side = (kl < kOwn0) ? 0 : (kl >= kOwn1 ? 1 : -1);
    if (side < 0) return;
    if ((side == 0 && kl < kOwn0 - reachLo) || (side == 1 && kl >= kOwn1 + reachHi)) S->slabError[0] = 1;
    int slot = atomicAdd(&S->sendCount[side], 1);
    if (slot >= cap) { S->slabError[1] = 1; return; }
    float *b = side ? hi : lo;
    b[slot] = p.px[t];
    b[(size_t)cap + slot] = p.py[t];
    b[2ull * cap + slot] = p.pz[t];
    b[3ull * cap + slot] = p.vx[t];
    b[4ull * cap + slot] = p.vy[t];
    b[5ull * cap + slot] = p.vz[t];
    if (ids) ((int *)b)[6ull * cap + slot] = ids[t];
}

__global__ void k_slab_reset_counts(DeviceScalars *S) {
    S->sendCount[0] = S->sendCount[1] = 0;
    S->recvCount[0] = S->recvCount[1] = 0;
    S->slabError[0] = S->slabError[1] = 0;      // [2] is set by the advance kernels that ran just before
}

void slab_drop_ghosts_and_migrate(flip_ctx *c) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    const int n = c->np;      // owned, at [ownedBegin, ownedEnd)
    int wantCap = std::max(std::max(1 << 16, c->capacity / 8), c->sendCap);
    // thickness of the neighbouring slabs (global planes): how far an emigrant may land
    int reachLo = 0, reachHi = 0;
    {
        int a = 0, b = 0;
        if (hasLo) { flip_slab_range(d.Kg, c->nranks, c->rank - 1, &a, &b); reachLo = b - a; }
        if (hasHi) { flip_slab_range(d.Kg, c->nranks, c->rank + 1, &a, &b); reachHi = b - a; }
    }
    int sendLo = 0, sendHi = 0, recvLo = 0, recvHi = 0;
    for (int attempt = 0;; attempt++) {
        if (wantCap > c->sendCap) {
            for (int s = 0; s < 2; s++) {
                cudaFree(c->sendBuf[s]);
                c->sendBuf[s] = nullptr;
                FLIP_CUDA_CHECK(cudaMalloc(&c->sendBuf[s], sizeof(float) * 7ull * wantCap));
            }
            c->sendCap = wantCap;
        }
        k_slab_reset_counts<<<1, 1, 0, st>>>(c->dS); c->launches++;
        if (n > 0) {
            ParticleSoA p = c->P[c->cur_buf];
            p.px += c->ownedBegin; p.py += c->ownedBegin; p.pz += c->ownedBegin;
            p.vx += c->ownedBegin; p.vy += c->ownedBegin; p.vz += c->ownedBegin;
            k_slab_pack_emigrants<<<cdiv(n, TPB), TPB, 0, st>>>(p, c->trackIds ? c->pid[c->cur_buf] + c->ownedBegin : nullptr, n,
                                                               1.0 / d.dx, d.kOff, d.kOwn0, d.kOwn1, reachLo, reachHi,
                                                               c->sendBuf[0], c->sendBuf[1], c->sendCap, c->dS);
            c->launches++;
        }
        exchange_counts(c);     // also sums slabError over the ranks: every rank takes the same branch below
        sendLo = c->hS->sendCount[0]; sendHi = c->hS->sendCount[1];
        recvLo = c->hS->recvCount[0]; recvHi = c->hS->recvCount[1];
        if (c->hS->slabError[2])
            throw ApiError(FLIP_ERR_DOMAIN, "z-slab advection: a particle's RK3 sample left the halo planes (substep longer than the "
                                            "CFL bound the halo was sized for): raise flip_set_halo or lower the frame time");
        const int jumped = c->hS->slabError[0], overflowed = c->hS->slabError[1];
        if (!jumped && !overflowed) break;
        // some rank's buffer overflowed (and nothing worse): every rank grows its buffers and packs again
        if (!jumped && attempt < 8) { wantCap = c->sendCap * 2; continue; }
        throw ApiError(FLIP_ERR_DOMAIN, "z-slab migration: a particle crossed more than one slab in a substep (slabs thinner than "
                                        "the particle reach), or the migration buffers could not be grown");
    }
    const int total = n + recvLo + recvHi;
    particles_alloc(c, c->ownedBegin + total);
    SoAPtrs P = ptrs_of(c, c->cur_buf);
    const int at = c->ownedEnd;     // immigrants overwrite the upper ghosts, which are no longer needed
    const size_t cap = c->sendCap;
    comm_group_begin(c->comm);
    for (int a = 0; a < 6; a++) {
        if (hasLo) {
            comm_send(c->comm, c->sendBuf[0] + a * cap, sizeof(float) * (size_t)sendLo, c->rank - 1, st);
            comm_recv(c->comm, P.a[a] + at, sizeof(float) * (size_t)recvLo, c->rank - 1, st);
        }
        if (hasHi) {
            comm_send(c->comm, c->sendBuf[1] + a * cap, sizeof(float) * (size_t)sendHi, c->rank + 1, st);
            comm_recv(c->comm, P.a[a] + at + recvLo, sizeof(float) * (size_t)recvHi, c->rank + 1, st);
        }
    }
    if (c->trackIds) {
        if (hasLo) {
            comm_send(c->comm, c->sendBuf[0] + 6 * cap, sizeof(int) * (size_t)sendLo, c->rank - 1, st);
            comm_recv(c->comm, P.id + at, sizeof(int) * (size_t)recvLo, c->rank - 1, st);
        }
        if (hasHi) {
            comm_send(c->comm, c->sendBuf[1] + 6 * cap, sizeof(int) * (size_t)sendHi, c->rank + 1, st);
            comm_recv(c->comm, P.id + at + recvLo, sizeof(int) * (size_t)recvHi, c->rank + 1, st);
        }
    }
    comm_group_end(c->comm);
    // removal rules + sort over (owned - emigrants + immigrants); emigrants are dropped by the owned-plane test
    particles_sort(c, true, c->frameDt, c->ownedBegin, total, true);
    c->npStore = c->np;
}

// Halo planes of a grid whose plane kl holds `planeElems` entries.  shift = 0 for cell-plane grids
// (U, V, masks of U/V), 1 for the face-plane grid W: face plane kOwn0 / kOwn1 is computed identically
// on both sides, so W halos start one plane further out.
template <class T>
static void exchange_planes(flip_ctx *c, T *f, int planeElems, int shift) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    const int H = c->halo;
    const size_t pe = (size_t)planeElems;
    const size_t bytes = sizeof(T) * pe * H;
    comm_group_begin(c->comm);
    if (hasLo) {
        comm_send(c->comm, f + pe * (d.kOwn0 + shift), bytes, c->rank - 1, st);
        comm_recv(c->comm, f + pe * (d.kOwn0 - H), bytes, c->rank - 1, st);
    }
    if (hasHi) {
        comm_send(c->comm, f + pe * (d.kOwn1 - H), bytes, c->rank + 1, st);
        comm_recv(c->comm, f + pe * (d.kOwn1 + shift), bytes, c->rank + 1, st);
    }
    comm_group_end(c->comm);
}

void slab_exchange_planes(flip_ctx *c, float *f, int planeElems, int facePlanes) { exchange_planes(c, f, planeElems, facePlanes); }
void slab_exchange_planes_u8(flip_ctx *c, unsigned char *f, int planeElems, int facePlanes) { exchange_planes(c, f, planeElems, facePlanes); }

// One cell plane of a float array each way, for a (multigrid) level whose owned planes are [k0,k1) and whose
// plane holds planeElems entries.
void slab_exchange_cell_plane_f32(flip_ctx *c, float *v, int planeElems, int k0, int k1) {
    cudaStream_t st = c->stream;
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    const size_t pe = (size_t)planeElems;
    const size_t bytes = sizeof(float) * pe;
    if (peer_on(c)) {
        peer_exchange_planes(c, v + pe * k0, v + pe * (k0 - 1), v + pe * (k1 - 1), v + pe * k1, bytes);
        return;
    }
    comm_group_begin(c->comm);
    if (hasLo) {
        comm_send(c->comm, v + pe * k0, bytes, c->rank - 1, st);
        comm_recv(c->comm, v + pe * (k0 - 1), bytes, c->rank - 1, st);
    }
    if (hasHi) {
        comm_send(c->comm, v + pe * (k1 - 1), bytes, c->rank + 1, st);
        comm_recv(c->comm, v + pe * k1, bytes, c->rank + 1, st);
    }
    comm_group_end(c->comm);
}

// One cell plane of a dense fp64 vector each way: my first owned plane -> lower neighbour's plane
// kOwn1, my last owned plane -> upper neighbour's plane kOwn0-1.
void slab_exchange_vector_halo(flip_ctx *c, double *v) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    const bool hasLo = c->rank > 0, hasHi = c->rank < c->nranks - 1;
    const size_t pe = (size_t)d.I * d.J;
    const size_t bytes = sizeof(double) * pe;
    if (peer_on(c)) {
        peer_exchange_planes(c, v + pe * d.kOwn0, v + pe * (d.kOwn0 - 1), v + pe * (d.kOwn1 - 1), v + pe * d.kOwn1, bytes);
        return;
    }
    comm_group_begin(c->comm);
    if (hasLo) {
        comm_send(c->comm, v + pe * d.kOwn0, bytes, c->rank - 1, st);
        comm_recv(c->comm, v + pe * (d.kOwn0 - 1), bytes, c->rank - 1, st);
    }
    if (hasHi) {
        comm_send(c->comm, v + pe * (d.kOwn1 - 1), bytes, c->rank + 1, st);
        comm_recv(c->comm, v + pe * d.kOwn1, bytes, c->rank + 1, st);
    }
    comm_group_end(c->comm);
}

// One scalar of the distributed PCG (dot product or residual maximum) over all slabs, in place on the device.
void slab_allreduce_scalar(flip_ctx *c, void *val, int kind) {
    if (peer_on(c)) peer_allreduce(c, val, 1, kind);
    else comm_allreduce(c->comm, val, 1, kind, c->stream);
}

}  // namespace flip
