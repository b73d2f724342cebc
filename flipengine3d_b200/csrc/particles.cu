// Particle-side kernels of the FLIP step: per-step cell sort (+ the reference's particle removal
// rules), liquid SDF + particle-to-grid gather, PIC/FLIP grid-to-particle update and RK3 advection
// with solid collision.  Each kernel cites the reference code whose results it reproduces.
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include "device_math.cuh"
#include "flip_internal.h"

namespace flip {

static constexpr int TPB = 256;

// ------------------------------------------------------------------------------------------------
// allocation
// ------------------------------------------------------------------------------------------------
// One allocation per SoA set: six float arrays of `cap` entries back to back (24*cap bytes), so the
// idle half of the ping-pong pair doubles as the AoS staging area of uploads and downloads.
static void soa_alloc(ParticleSoA &p, int cap) {
    float *blk = nullptr;
    FLIP_CUDA_CHECK(cudaMalloc(&blk, sizeof(float) * 6ull * cap));
    // the gather reads aligned 4-particle windows that may reach past the last particle: keep the slack finite
    FLIP_CUDA_CHECK(cudaMemset(blk, 0, sizeof(float) * 6ull * cap));
    p.px = blk; p.py = blk + (size_t)cap; p.pz = blk + 2ull * cap;
    p.vx = blk + 3ull * cap; p.vy = blk + 4ull * cap; p.vz = blk + 5ull * cap;
}
static void soa_free(ParticleSoA &p) {
    cudaFree(p.px);
    p = ParticleSoA();
}

void particles_free(flip_ctx *c) {
    soa_free(c->P[0]);
    soa_free(c->P[1]);
    cudaFree(c->cellOfParticle); c->cellOfParticle = nullptr;
    cudaFree(c->sortIdx); c->sortIdx = nullptr;
    cudaFree(c->srcIdx); c->srcIdx = nullptr;
    cudaFree(c->pid[0]); cudaFree(c->pid[1]); c->pid[0] = c->pid[1] = nullptr;
    c->capacity = 0;
}

void particles_alloc(flip_ctx *c, int capacity) {
    if (capacity <= c->capacity) return;
    // keep live particles when growing
    ParticleSoA old[2] = {c->P[0], c->P[1]};
    int oldCap = c->capacity;
    int cap = (capacity + capacity / 16 + 1024 + 3) & ~3;   // multiple of 4: every component array stays 16-byte aligned
    ParticleSoA nw[2];
    soa_alloc(nw[0], cap);
    soa_alloc(nw[1], cap);
    int *npid[2] = {nullptr, nullptr};
    FLIP_CUDA_CHECK(cudaMalloc(&npid[0], sizeof(int) * cap));
    FLIP_CUDA_CHECK(cudaMalloc(&npid[1], sizeof(int) * cap));
    const int live = c->npStore > c->np ? c->npStore : c->np;   // owned + ghosts
    if (oldCap > 0 && live > 0)
        FLIP_CUDA_CHECK(cudaMemcpyAsync(npid[c->cur_buf], c->pid[c->cur_buf], sizeof(int) * live, cudaMemcpyDeviceToDevice, c->stream));
    if (oldCap > 0 && live > 0) {
        const ParticleSoA &s = old[c->cur_buf];
        ParticleSoA &t = nw[c->cur_buf];
        size_t b = sizeof(float) * live;
        FLIP_CUDA_CHECK(cudaMemcpyAsync(t.px, s.px, b, cudaMemcpyDeviceToDevice, c->stream));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(t.py, s.py, b, cudaMemcpyDeviceToDevice, c->stream));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(t.pz, s.pz, b, cudaMemcpyDeviceToDevice, c->stream));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(t.vx, s.vx, b, cudaMemcpyDeviceToDevice, c->stream));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(t.vy, s.vy, b, cudaMemcpyDeviceToDevice, c->stream));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(t.vz, s.vz, b, cudaMemcpyDeviceToDevice, c->stream));
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    if (oldCap > 0) {
        soa_free(old[0]);
        soa_free(old[1]);
        cudaFree(c->cellOfParticle); cudaFree(c->sortIdx); cudaFree(c->srcIdx);
        cudaFree(c->pid[0]); cudaFree(c->pid[1]);
    }
    c->P[0] = nw[0];
    c->P[1] = nw[1];
    c->pid[0] = npid[0];
    c->pid[1] = npid[1];
    FLIP_CUDA_CHECK(cudaMalloc(&c->cellOfParticle, sizeof(int) * cap));
    FLIP_CUDA_CHECK(cudaMalloc(&c->sortIdx, sizeof(int) * cap));
    FLIP_CUDA_CHECK(cudaMalloc(&c->srcIdx, sizeof(int) * cap));
    c->capacity = cap;
}

// ------------------------------------------------------------------------------------------------
// AoS <-> SoA at the API boundary (MarkerParticle, markerparticle.h:30-42)
// ------------------------------------------------------------------------------------------------
__global__ void k_aos_to_soa(const float *__restrict__ aos, int n, ParticleSoA p) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float *a = aos + 6ll * t;
    p.px[t] = a[0]; p.py[t] = a[1]; p.pz[t] = a[2];
    p.vx[t] = a[3]; p.vy[t] = a[4]; p.vz[t] = a[5];
}
__global__ void k_split_to_soa(const float *__restrict__ pos, const float *__restrict__ vel, int n, ParticleSoA p) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    p.px[t] = pos[3ll * t]; p.py[t] = pos[3ll * t + 1]; p.pz[t] = pos[3ll * t + 2];
    p.vx[t] = vel[3ll * t]; p.vy[t] = vel[3ll * t + 1]; p.vz[t] = vel[3ll * t + 2];
}
__global__ void k_soa_to_aos(ParticleSoA p, int n, float *__restrict__ aos) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float *a = aos + 6ll * t;
    a[0] = p.px[t]; a[1] = p.py[t]; a[2] = p.pz[t];
    a[3] = p.vx[t]; a[4] = p.vy[t]; a[5] = p.vz[t];
}
__global__ void k_soa_to_xyz(const float *x, const float *y, const float *z, int n, float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    out[3ll * t] = x[t]; out[3ll * t + 1] = y[t]; out[3ll * t + 2] = z[t];
}

// ------------------------------------------------------------------------------------------------
// Cell sort + removal rules.
//
// Reference semantics reproduced (fluidsimulation.cpp:4298-4353):
//   * particles whose solid SDF sample is < 0 are deleted and never counted;
//   * per cell, in particle-index order, the first maxMarkerParticlesPerCell are kept;
//   * of those, particles faster than the histogram speed limit are deleted (but still counted).
// The load-time filter is the in-domain test of _loadMarkerParticles / _addMarkerParticle
// (:2773-2785, :2637-2642).  Order inside a cell = previous index order (stable), so the whole
// pipeline is deterministic.
// ------------------------------------------------------------------------------------------------
struct SortParams {
    int n;
    int I, J, K;          // local grid
    int kOff;             // global k of local plane 0
    int kOwn0, kOwn1;     // owned local planes
    int ownedOnly;        // z-slab: silently drop particles outside the owned planes (they were sent to a neighbour)
    double dx, invdx;
    int applyRules;
    int maxPerCell;
    float farFromSolid;   // 3dx
};

__device__ __forceinline__ bool sort_keeps(const SortParams &sp, float z) {
    if (!sp.ownedOnly) return true;
    int kl = pos2idx(z, sp.invdx) - sp.kOff;
    return kl >= sp.kOwn0 && kl < sp.kOwn1;
}

// _getMarkerParticleSpeedLimit, first loop (:4300-4305): histogram of |v| in CFL*dx/dt_frame bins.
__global__ void k_speed_hist(ParticleSoA p, SortParams sp, double speedLimitStep, int nbins, DeviceScalars *S) {
    __shared__ int sh[8];
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int n = sp.n;
    if (t < n && sort_keeps(sp, p.pz[t])) {
        float len = length3(p.vx[t], p.vy[t], p.vz[t]);
        double speed = (double)len;
        double b = fmin(floor(speed / speedLimitStep), (double)(nbins - 1));
        int bi = (int)b;
        if (bi > 0) atomicAdd(&sh[bi], 1);   // bin 0 is never read by the walk below
    }
    __syncthreads();
    if (threadIdx.x > 0 && threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(&S->speedHist[threadIdx.x], sh[threadIdx.x]);
}

// _getMarkerParticleSpeedLimit, second loop (:4307-4320)
__global__ void k_speed_limit(int n, double speedLimitStep, int nbins, double maxpct, int maxabs, int enabled, DeviceScalars *S) {
    int maxRemovalCount = (int)fmin((double)(int)((double)n * maxpct), (double)maxabs);
    double maxspeed = nbins * speedLimitStep;
    int cur = 0;
    for (int i = nbins - 1; i > 0; i--) {
        if (cur + S->speedHist[i] > maxRemovalCount) break;
        cur += S->speedHist[i];
        maxspeed = i * speedLimitStep;
    }
    // disabled (disableExtremeVelocityRemoval): no speed exceeds the limit (its square is +inf in k_classify)
    S->maxSpeedLimit = enabled ? (float)maxspeed : 3.0e38f;
    for (int i = 0; i < 8; i++) S->speedHist[i] = 0;
}

// cellOf[t] = destination cell (or -1), with the extreme-velocity flag in bit 30; rankOf[t] = arrival number of
// the particle in its cell (any order: k_cell_finalize orders the cell by previous index afterwards).  Particles
// are cell-sorted from the previous step and move less than a cell per substep on average, so a warp holds few
// distinct destination cells: one counter atomic per distinct cell of the warp instead of one per particle.
static constexpr int FAST_BIT = 1 << 30;
__global__ void k_classify(ParticleSoA p, SortParams sp, const float *__restrict__ phiS, int *__restrict__ cellOf,
                           int *__restrict__ rankOf, int *__restrict__ cellCount, DeviceScalars *S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < sp.n;
    const int tt = live ? t : 0;
    float x = p.px[tt], y = p.py[tt], z = p.pz[tt];
    int i = pos2idx(x, sp.invdx), j = pos2idx(y, sp.invdx), k = pos2idx(z, sp.invdx) - sp.kOff;
    int cell = -1;
    bool inRange = (i >= 0 && j >= 0 && k >= 0 && i < sp.I && j < sp.J && k < sp.K);
    bool gone = sp.ownedOnly && (k < sp.kOwn0 || k >= sp.kOwn1);   // now lives on a neighbouring slab
    unsigned char f = 0;
    if (!live) {
        cell = -1;
    } else if (gone) {
        cell = -1;
    } else if (inRange) {
        cell = i + sp.I * (j + sp.J * k);
        if (sp.applyRules) {
            // MeshLevelSet::trilinearInterpolateSolidPoints  meshlevelset.h:325-332.  Only its sign is used: where the
            // lower corner node of the cell is at least 3dx outside the solid, the other seven nodes of the cell are
            // outside too (a distance field changes by at most sqrt(3)dx across a cell) and so is their blend.
            float phi = 1.0f;
            const float corner = __ldg(phiS + (size_t)i + (size_t)(sp.I + 1) * ((size_t)j + (size_t)(sp.J + 1) * (size_t)k));
            if (!(corner >= sp.farFromSolid))
                phi = sample_scalar(phiS, sp.I + 1, sp.J + 1, sp.K + 1, sp.dx, sp.invdx, x, y, z, sp.kOff);
            if (phi < 0.0f) {
                cell = -1;
                atomicAdd(&S->removedSolid, 1);
            } else {
                float vx = p.vx[tt], vy = p.vy[tt], vz = p.vz[tt];
                float ms = S->maxSpeedLimit;
                double maxspeedsq = (double)fmul(ms, ms);   // float*float then widened (:4328)
                if ((double)lengthsq3(vx, vy, vz) > maxspeedsq) f = 1;
            }
        }
    } else if (sp.applyRules) {
        // cannot happen after _resolveCollision (positions are clamped into the boundary box); counted as solid
        atomicAdd(&S->removedSolid, 1);
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, cell);
    int rank = 0;
    if (cell >= 0) {
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cellCount[cell], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        rank = base + __popc(peers & ((1u << lane) - 1u));
    }
    if (live) {
        cellOf[t] = (cell >= 0 && f) ? (cell | FAST_BIT) : cell;
        rankOf[t] = rank;
    }
}

// sortIdx entry = previous index, with the extreme-velocity flag in bit 30
__global__ void k_scatter_idx(int n, const int *__restrict__ cellOf, const int *__restrict__ rankOf,
                              const int *__restrict__ startA, int *__restrict__ sortIdx) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int cf = cellOf[t];
    if (cf < 0) return;
    const int cell = cf & ~FAST_BIT;
    sortIdx[startA[cell] + rankOf[t]] = t | (cf & FAST_BIT);
}

// One thread per cell: order the cell's candidates by previous index, apply the per-cell cap and
// the speed rule, leave the kept ones at the front of the cell's range, publish the kept count.
__global__ void k_cell_finalize(int nC, const int *__restrict__ startA, int *__restrict__ sortIdx,
                                int *__restrict__ keptCount, int applyRules, int maxPerCell, DeviceScalars *S) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    int b = startA[c], e = startA[c + 1];
    int n = e - b;
    if (n == 0) { keptCount[c] = 0; return; }
    constexpr int SMALL = 16;
    if (n <= SMALL && n <= maxPerCell) {
        // the usual cell (~8 particles): every load is independent, the order comes from a rank count in registers.
        // final slot of a kept particle = number of kept particles of the cell with a smaller previous index
        int v[SMALL];
        bool drop[SMALL];
#pragma unroll
        for (int a = 0; a < SMALL; a++) {
            const int w = (a < n) ? sortIdx[b + a] : 0x3fffffff;
            v[a] = w & ~FAST_BIT;
            drop[a] = (a < n) ? (applyRules && (w & FAST_BIT)) : true;
        }
        int kept = 0;
#pragma unroll
        for (int a = 0; a < SMALL; a++) {
            if (a < n && !drop[a]) {
                int r = 0;
#pragma unroll
                for (int q = 0; q < SMALL; q++) r += (!drop[q] && v[q] < v[a]) ? 1 : 0;
                sortIdx[b + r] = v[a];
                kept++;
            }
        }
        if (kept != n) atomicAdd(&S->removedFast, n - kept);
        keptCount[c] = kept;
        return;
    }
    // insertion sort by previous index (the cap is 250)
    for (int a = b + 1; a < e; a++) {
        int w = sortIdx[a];
        int v = w & ~FAST_BIT;
        int q = a - 1;
        while (q >= b && (sortIdx[q] & ~FAST_BIT) > v) { sortIdx[q + 1] = sortIdx[q]; q--; }
        sortIdx[q + 1] = w;
    }
    int kept = 0, crowded = 0, fastc = 0;
    for (int a = b; a < e; a++) {
        const int w = sortIdx[a];
        if (applyRules && a - b >= maxPerCell) { crowded++; continue; }
        if (applyRules && (w & FAST_BIT)) { fastc++; continue; }
        sortIdx[b + kept] = w & ~FAST_BIT;
        kept++;
    }
    if (crowded) atomicAdd(&S->removedCrowded, crowded);
    if (fastc) atomicAdd(&S->removedFast, fastc);
    keptCount[c] = kept;
}

// also marks the 4x4x4-cell blocks that hold particles (the gather kernel skips empty space with it)
// and, when the rows are whole words long (bits != nullptr), the one-bit-per-cell occupancy map of k_occ_bits
__global__ void k_build_src(int nC, const int *__restrict__ startA, const int *__restrict__ start,
                            const int *__restrict__ sortIdx, int *__restrict__ srcIdx, DeviceScalars *S,
                            unsigned char *__restrict__ occ, int I, int J, int oI, int oJ, unsigned int *__restrict__ bits) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (bits) {      // nC is a multiple of 32 here: whole warps pass or return together
        const unsigned int m = __ballot_sync(0xffffffffu, c < nC && start[c + 1] > start[c]);
        if ((threadIdx.x & 31) == 0 && c < nC) bits[c >> 5] = m;
    }
    if (c >= nC) return;
    int b = start[c], e = start[c + 1];
    if (e > b) {
        int i = c % I, j = (c / I) % J, k = c / (I * J);
        occ[(i >> 2) + oI * ((j >> 2) + oJ * (k >> 2))] = 1;
    }
    int a = startA[c];
    for (int q = b; q < e; q++) srcIdx[q] = sortIdx[a + (q - b)];
    if (c == nC - 1) S->numParticles = e;
}

__global__ void k_iota(int *p, int n, int base) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = base + t;
}

__global__ void k_gather(ParticleSoA src, ParticleSoA dst, const int *__restrict__ srcIdx, int nmax, DeviceScalars *S,
                         const int *__restrict__ idSrc, int *__restrict__ idDst) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int n = S->numParticles;
    float sp2 = 0.0f;
    if (t < n && t < nmax) {
        int s = srcIdx[t];
        if (idSrc) idDst[t] = idSrc[s];
        dst.px[t] = src.px[s]; dst.py[t] = src.py[s]; dst.pz[t] = src.pz[s];
        float vx = src.vx[s], vy = src.vy[s], vz = src.vz[s];
        dst.vx[t] = vx; dst.vy[t] = vy; dst.vz[t] = vz;
        sp2 = lengthsq3(vx, vy, vz);   // vmath::dot(mp.velocity, mp.velocity)  :5558
    }
    // one atomic per block at most, and only when it can raise the maximum (hundreds of thousands of
    // same-address atomics would serialise in L2)
    __shared__ float shmax[TPB / 32];
    sp2 = warp_max(sp2);
    if ((threadIdx.x & 31) == 0) shmax[threadIdx.x >> 5] = sp2;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < TPB / 32) ? shmax[threadIdx.x] : 0.0f;
        v = warp_max(v);
        if (threadIdx.x == 0 && v > 0.0f && __float_as_uint(v) > __ldcg(&S->maxSpeedSqBits))
            atomicMax(&S->maxSpeedSqBits, __float_as_uint(v));
    }
}

__global__ void k_reset_sort_scalars(DeviceScalars *S) {
    S->removedSolid = 0; S->removedCrowded = 0; S->removedFast = 0;
    S->maxSpeedSqBits = 0u;
    S->numParticles = 0;
}

static void ensure_scan_temp(flip_ctx *c, int n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (int *)nullptr, (int *)nullptr, n, c->stream);
    if (bytes > c->scanTempBytes) {
        cudaFree(c->scanTemp);
        FLIP_CUDA_CHECK(cudaMalloc(&c->scanTemp, bytes));
        c->scanTempBytes = bytes;
    }
}

void scalars_to_host(flip_ctx *c) {
    FLIP_CUDA_CHECK(cudaMemcpyAsync(c->hS, c->dS, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream));
    FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

static ParticleSoA soa_offset(const ParticleSoA &p, int off) {
    ParticleSoA q = p;
    q.px += off; q.py += off; q.pz += off; q.vx += off; q.vy += off; q.vz += off;
    return q;
}

// Sorts `count` particles of P[cur] starting at `srcOffset` into P[1-cur] (from index 0), grouped by
// cell, and rebuilds cellStart.  ownedOnly (z-slabs): particles outside the owned planes are dropped.
void particles_sort(flip_ctx *c, bool applyRules, double frameDt, int srcOffset, int count, bool ownedOnly) {
    const Dims &d = c->d;
    int n = count;
    cudaStream_t st = c->stream;
    ParticleSoA src = soa_offset(c->P[c->cur_buf], srcOffset);
    ParticleSoA &dst = c->P[1 - c->cur_buf];
    int nC = d.nC;
    if (nC >= FAST_BIT || n >= FAST_BIT) throw ApiError(FLIP_ERR_UNSUPPORTED, "more than 2^30 cells or particles per GPU");
    size_t ktSort = kt_begin(c);
    k_reset_sort_scalars<<<1, 1, 0, st>>>(c->dS); c->launches++;
    FLIP_CUDA_CHECK(cudaMemsetAsync(c->cellCount, 0, sizeof(int) * (nC + 1), st));
    SortParams sp;
    sp.n = n; sp.I = d.I; sp.J = d.J; sp.K = d.K; sp.dx = d.dx; sp.invdx = 1.0 / d.dx;
    sp.kOff = d.kOff; sp.kOwn0 = d.kOwn0; sp.kOwn1 = d.kOwn1; sp.ownedOnly = ownedOnly ? 1 : 0;
    sp.applyRules = applyRules ? 1 : 0; sp.maxPerCell = c->maxParticlesPerCell;
    sp.farFromSolid = (float)(3.0 * d.dx);
    if (applyRules) {
        // _removeMarkerParticles(_currentFrameDeltaTime): bins are CFL*dx/dt_FRAME wide
        double speedLimitStep = c->CFL * d.dx / frameDt;
        int nGlobal = n;
        if (n > 0 && !c->speedHistReady) { k_speed_hist<<<cdiv(n, TPB), TPB, 0, st>>>(src, sp, speedLimitStep, c->maxSubsteps, c->dS); c->launches++; }
        c->speedHistReady = false;
        if (slab_on(c)) {
            // the rule is global: histogram and particle count summed over the slabs
            comm_allreduce(c->comm, c->dS->speedHist, 8, COMM_SUM_I32, st);
            nGlobal = c->np_global;
        }
        k_speed_limit<<<1, 1, 0, st>>>(nGlobal, speedLimitStep, c->maxSubsteps, c->maxExtremeVelocityRemovalPercent,
                                       c->maxExtremeVelocityRemovalAbsolute, c->extremeVelocityRemoval ? 1 : 0, c->dS); c->launches++;
    }
    if (n > 0) {
        // the arrival ranks live in srcIdx, which is only rebuilt (k_build_src) after the index scatter has used them
        k_classify<<<cdiv(n, TPB), TPB, 0, st>>>(src, sp, c->phiS, c->cellOfParticle, c->srcIdx, c->cellCount, c->dS);
        c->launches++;
    }
    ensure_scan_temp(c, nC + 1);
    cub::DeviceScan::ExclusiveSum(c->scanTemp, c->scanTempBytes, c->cellCount, c->cellStartA, nC + 1, st); c->launches++;
    if (n > 0) {
        k_scatter_idx<<<cdiv(n, TPB), TPB, 0, st>>>(n, c->cellOfParticle, c->srcIdx, c->cellStartA, c->sortIdx);
        c->launches++;
    }
    // overwrites the candidate counts with the kept counts, cell by cell
    k_cell_finalize<<<cdiv(nC, TPB), TPB, 0, st>>>(nC, c->cellStartA, c->sortIdx, c->cellCount,
                                                   applyRules ? 1 : 0, c->maxParticlesPerCell, c->dS);
    c->launches++;
    FLIP_CUDA_CHECK(cudaMemsetAsync(c->cellCount + nC, 0, sizeof(int), st));
    cub::DeviceScan::ExclusiveSum(c->scanTemp, c->scanTempBytes, c->cellCount, c->cellStart, nC + 1, st); c->launches++;
    {
        int oI = (d.I + 3) >> 2, oJ = (d.J + 3) >> 2, oK = (d.K + 3) >> 2;
        size_t need = (size_t)oI * oJ * oK;
        if (need > c->occBytes) {
            cudaFree(c->occ);
            FLIP_CUDA_CHECK(cudaMalloc(&c->occ, need));
            c->occBytes = need;
        }
        FLIP_CUDA_CHECK(cudaMemsetAsync(c->occ, 0, need, st));
        c->occBitsValid = c->occBits && (d.I % 32 == 0);
        k_build_src<<<cdiv(nC, TPB), TPB, 0, st>>>(nC, c->cellStartA, c->cellStart, c->sortIdx, c->srcIdx, c->dS, c->occ, d.I,
                                                   d.J, oI, oJ, c->occBitsValid ? c->occBits : nullptr);
        c->launches++;
    }
    if (n > 0) {
        k_gather<<<cdiv(n, TPB), TPB, 0, st>>>(src, dst, c->srcIdx, n, c->dS, c->trackIds ? c->pid[c->cur_buf] + srcOffset : nullptr,
                                               c->pid[1 - c->cur_buf]); c->launches++;
    }
    if (slab_on(c) && ownedOnly) {
        // CFL uses the maximum speed over all slabs; particle and removal counts of the whole domain
        comm_allreduce(c->comm, &c->dS->maxSpeedSqBits, 1, COMM_MAX_U32, st);
        FLIP_CUDA_CHECK(cudaMemcpyAsync(&c->dS->globalParticles, &c->dS->numParticles, sizeof(int), cudaMemcpyDeviceToDevice, st));
        comm_allreduce(c->comm, &c->dS->globalParticles, 1, COMM_SUM_I32, st);
        comm_allreduce(c->comm, &c->dS->removedSolid, 3, COMM_SUM_I32, st);
    }
    kt_end(c, FLIP_KERNEL_SORT, ktSort);
    scalars_to_host(c);
    c->np = c->hS->numParticles;
    c->npStore = c->np;
    if (!slab_on(c)) c->np_global = c->np;
    else if (ownedOnly) c->np_global = c->hS->globalParticles;
    c->cur_buf = 1 - c->cur_buf;
    c->stepCounter++;
    c->ownedBegin = 0;
    c->ownedEnd = c->np;
    c->ghostsPresent = false;
}

void particles_upload_aos(flip_ctx *c, const float *aos6, int n) {
    particles_alloc(c, n);
    c->np = n;
    if (n > 0) {
        // host AoS -> idle SoA block (as raw AoS) -> transpose into the live block
        float *tmp = c->P[1 - c->cur_buf].px;
        FLIP_CUDA_CHECK(cudaMemcpyAsync(tmp, aos6, sizeof(float) * 6ll * n, cudaMemcpyHostToDevice, c->stream));
        k_aos_to_soa<<<cdiv(n, TPB), TPB, 0, c->stream>>>(tmp, n, c->P[c->cur_buf]); c->launches++;
        if (c->trackIds) { k_iota<<<cdiv(n, TPB), TPB, 0, c->stream>>>(c->pid[c->cur_buf], n, c->particleIdBase); c->launches++; }
    }
    particles_sort(c, false, 0.0, 0, n, slab_on(c));
}

void particles_upload_split(flip_ctx *c, const float *pos, const float *vel, int n) {
    particles_alloc(c, n);
    c->np = n;
    if (n > 0) {
        float *tp = c->P[1 - c->cur_buf].px;
        float *tv = tp + 3ull * n;
        FLIP_CUDA_CHECK(cudaMemcpyAsync(tp, pos, sizeof(float) * 3ll * n, cudaMemcpyHostToDevice, c->stream));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(tv, vel, sizeof(float) * 3ll * n, cudaMemcpyHostToDevice, c->stream));
        k_split_to_soa<<<cdiv(n, TPB), TPB, 0, c->stream>>>(tp, tv, n, c->P[c->cur_buf]); c->launches++;
        if (c->trackIds) { k_iota<<<cdiv(n, TPB), TPB, 0, c->stream>>>(c->pid[c->cur_buf], n, c->particleIdBase); c->launches++; }
    }
    particles_sort(c, false, 0.0, 0, n, slab_on(c));
}

void particles_download_aos(flip_ctx *c, float *aos6) {
    int n = c->np;
    if (n == 0) return;
    float *tmp = c->P[1 - c->cur_buf].px;
    k_soa_to_aos<<<cdiv(n, TPB), TPB, 0, c->stream>>>(soa_offset(c->P[c->cur_buf], c->ownedBegin), n, tmp); c->launches++;
    FLIP_CUDA_CHECK(cudaMemcpyAsync(aos6, tmp, sizeof(float) * 6ll * n, cudaMemcpyDeviceToHost, c->stream));
    FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void particles_download_ids(flip_ctx *c, int *ids) {
    if (c->np == 0) return;
    FLIP_CUDA_CHECK(cudaMemcpyAsync(ids, c->pid[c->cur_buf] + c->ownedBegin, sizeof(int) * c->np, cudaMemcpyDeviceToHost, c->stream));
    FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void particles_download_component(flip_ctx *c, float *xyz, int which) {
    int n = c->np;
    if (n == 0) return;
    float *tmp = c->P[1 - c->cur_buf].px;
    const ParticleSoA p = soa_offset(c->P[c->cur_buf], c->ownedBegin);
    if (which == 0) k_soa_to_xyz<<<cdiv(n, TPB), TPB, 0, c->stream>>>(p.px, p.py, p.pz, n, tmp);
    else k_soa_to_xyz<<<cdiv(n, TPB), TPB, 0, c->stream>>>(p.vx, p.vy, p.vz, n, tmp);
    c->launches++;
    FLIP_CUDA_CHECK(cudaMemcpyAsync(xyz, tmp, sizeof(float) * 3ll * n, cudaMemcpyDeviceToHost, c->stream));
    FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// ------------------------------------------------------------------------------------------------
// Liquid SDF + P2G gather over cell-sorted particles.
//
// One thread per point (i,j,k) of the extended index space (I+1)x(J+1)x(K+1).  The thread owns
//   phi(i,j,k)                (if i<I, j<J, k<K)   ParticleLevelSet min-splat  particlelevelset.cpp:582-630
//   U(i,j,k), V(i,j,k), W(i,j,k) (where in range)  VelocityAdvector radial splat velocityadvector.cpp:390-459
// and visits the particles of the 3x3x3 cells around (i,j,k) (every particle that can reach the
// three faces lies there because r = 0.866dx < dx); the SDF's 5x5x5 box is completed in a second
// phase only when the nearest particle so far is farther than 1.5dx (then particles outside 3x3x3
// could be nearer).  Summation order per face = cell order (k,j,i) then index order inside a cell:
// fixed, hence deterministic (the reference sums in particle-index order inside a 10^3 block; the
// difference is float summation order only, SURVEY §7 hard part 3).
//
// Arithmetic restated literally: the reference works in coordinates local to a
// 10-cell block; `p - offset`, `- blockOrigin`, node position `(float)l*dx` are evaluated here in
// the same order and precision, so results are bit-identical per term for any dx.
// ------------------------------------------------------------------------------------------------
struct GatherParams {
    int I, J, K;          // local grid
    int kOff;             // global k of local plane 0 (block origins and node positions use global indices)
    double dx, invdx;
    float hw;             // (float)(0.5*dx): _getDirectionOffset, velocityadvector.cpp:153-164
    double blockdxP2G;    // _chunkWidth * _dx (double)            velocityadvector.cpp:411
    double blockdxSDF;    // _blockwidth * _dx (double)            particlelevelset.cpp:593
    double invBlockdxSDF; // 1.0 / (double)(float)(_blockwidth*_dx) particlelevelset.cpp:483,487
    float r, rsq, coef1, coef2, coef3, eps;   // velocityadvector.cpp:393-399
    float rS, srS;        // SDF radius (float) and search radius 2r particlelevelset.cpp:592-594
    float maxDist;        // 3dx  particlelevelset.cpp:295
    float hwS;            // cell-centre half width as added in double (GridIndexToCellCenter grid3d.h:101)
    float halfDxSolid;    // post-process thresholds
    int packed;           // FLIP_SAMPLING_FAST: weights through the packed-f32x2 row; literal weights otherwise
};

__device__ __forceinline__ float kernel_weight(float d2, const GatherParams &g) {
    // 1.0f - coef1*d2*d2*d2 + coef2*d2*d2 - coef3*d2, left to right (velocityadvector.cpp:437)
    float a = fmul(fmul(fmul(g.coef1, d2), d2), d2);
    float b = fmul(fmul(g.coef2, d2), d2);
    float cc = fmul(g.coef3, d2);
    return fsub(fadd(fsub(1.0f, a), b), cc);
}

// Geometry of one node/cell in the reference's block-local frame (10-cell blocks).
struct NodeFrame {
    int bi, bj, bk, li, lj, lk;
    float ox, oy, oz;      // block origin
    float gx, gy, gz;      // node position in the block
    float cx, cy, cz;      // cell centre in the block
};

__device__ __forceinline__ NodeFrame node_frame(const GatherParams &g, int i, int j, int kg) {
    NodeFrame f;
    f.bi = i / 10; f.bj = j / 10; f.bk = kg / 10;
    f.li = i - f.bi * 10; f.lj = j - f.bj * 10; f.lk = kg - f.bk * 10;
    // block origins  (float)bi * blockdx  -> float  (grid3d.h:81-83)
    f.ox = (float)dmul((double)(float)f.bi, g.blockdxP2G);
    f.oy = (float)dmul((double)(float)f.bj, g.blockdxP2G);
    f.oz = (float)dmul((double)(float)f.bk, g.blockdxP2G);
    // node position (float)l*dx -> float   (grid3d.h:81)
    f.gx = (float)dmul((double)(float)f.li, g.dx);
    f.gy = (float)dmul((double)(float)f.lj, g.dx);
    f.gz = (float)dmul((double)(float)f.lk, g.dx);
    // cell centre (float)l*dx + hw in double -> float  (grid3d.h:101-104)
    const double hwd = 0.5 * g.dx;
    f.cx = (float)dadd(dmul((double)(float)f.li, g.dx), hwd);
    f.cy = (float)dadd(dmul((double)(float)f.lj, g.dx), hwd);
    f.cz = (float)dadd(dmul((double)(float)f.lk, g.dx), hwd);
    return f;
}

// dist = length(gpos - p) - r ; phi = min(3dx, dist) (particlelevelset.cpp:615-621, :295), then
// postProcessSignedDistanceField (particlelevelset.cpp:172-196)
__device__ __forceinline__ float finish_phi(const GatherParams &g, const float *__restrict__ phiS, float best2, int i, int j,
                                            int k) {
    const int I = g.I, J = g.J;
    float phi = g.maxDist;
    if (best2 < 1.0e38f) {
        float dist = fsub(__fsqrt_rn(best2), g.rS);
        if (dist < phi) phi = dist;
    }
    double dxd = g.dx;
    if ((double)phi < 0.5 * dxd) {
        // MeshLevelSet::getDistanceAtCellCenter  meshlevelset.cpp:152-162
        int ni = I + 1;
        long long nj = (long long)(I + 1) * (J + 1);
        long long n0 = (long long)i + (long long)ni * j + nj * k;
        float s = __ldg(phiS + n0);
        s = fadd(s, __ldg(phiS + n0 + 1));
        s = fadd(s, __ldg(phiS + n0 + ni));
        s = fadd(s, __ldg(phiS + n0 + ni + 1));
        s = fadd(s, __ldg(phiS + n0 + nj));
        s = fadd(s, __ldg(phiS + n0 + nj + 1));
        s = fadd(s, __ldg(phiS + n0 + nj + ni));
        s = fadd(s, __ldg(phiS + n0 + nj + ni + 1));
        if (fmul(0.125f, s) < 0.0f) phi = (float)dmul((double)-0.5f, dxd);
    }
    float epsf = (float)(0.005 * dxd);
    if (fabsf(phi) < epsf) phi = (phi > 0.0f) ? epsf : -epsf;
    return phi;
}

// One row (cj,ck) of the 3x3 cell rows around a node: particles [qb,qe).  DOV / DOW: whether particles of this
// row can reach the V / W face of the node (they cannot from row j+1 / plane k+1: r < dx).
template <bool DYADIC, bool DOV, bool DOW>
__device__ __forceinline__ void gather_row(const ParticleSoA &p, const GatherParams &g, const NodeFrame &f, int qb, int qe,
                                           float Xn, float Yn, float Zn, float Xh, float Yh, float Zh, float &su,
                                           float &wu, float &sv, float &wv, float &sw, float &ww, float &best2) {
    // four particles per step through 16-byte loads of each coordinate array (16-byte aligned, padded): lanes of
    // a warp read windows ~8 particles apart, so this cuts the L1 wavefronts per particle by four.  Summation
    // order stays q ascending.
    for (int q4 = qb & ~3; q4 < qe; q4 += 4) {
        const float4 X4 = __ldg(reinterpret_cast<const float4 *>(p.px + q4));
        const float4 Y4 = __ldg(reinterpret_cast<const float4 *>(p.py + q4));
        const float4 Z4 = __ldg(reinterpret_cast<const float4 *>(p.pz + q4));
        const float xs[4] = {X4.x, X4.y, X4.z, X4.w}, ys[4] = {Y4.x, Y4.y, Y4.z, Y4.w}, zs[4] = {Z4.x, Z4.y, Z4.z, Z4.w};
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int q = q4 + m;
            if (q < qb || q >= qe) continue;
            const float x = xs[m], y = ys[m], z = zs[m];
            float ax, ay, az, bx, by, bz, d2c;
            if (DYADIC) {
                ax = fsub(Xn, x); ay = fsub(Yn, y); az = fsub(Zn, z);
                bx = fsub(Xh, x); by = fsub(Yh, y); bz = fsub(Zh, z);
            } else {
                // block-local particle coordinates for the P2G components: (p - offset) - origin
                float xl = fsub(x, f.ox), yl = fsub(y, f.oy), zl = fsub(z, f.oz);                 // offset 0 on that axis
                float xh = fsub(fsub(x, g.hw), f.ox), yh = fsub(fsub(y, g.hw), f.oy), zh = fsub(fsub(z, g.hw), f.oz);
                ax = fsub(f.gx, xl); ay = fsub(f.gy, yl); az = fsub(f.gz, zl);                   // v = gpos - p
                bx = fsub(f.gx, xh); by = fsub(f.gy, yh); bz = fsub(f.gz, zh);
                d2c = lengthsq3(fsub(f.cx, xl), fsub(f.cy, yl), fsub(f.cz, zl));
            }
            const float ax2 = fmul(ax, ax), bx2 = fmul(bx, bx), by2 = fmul(by, by), bz2 = fmul(bz, bz);
            const float d2u = fadd(fadd(ax2, by2), bz2);
            if (d2u < g.rsq) { float w = kernel_weight(d2u, g); su = fadd(su, fmul(w, __ldg(p.vx + q))); wu = fadd(wu, w); }
            if (DOV) {
                const float d2v = fadd(fadd(bx2, fmul(ay, ay)), bz2);
                if (d2v < g.rsq) { float w = kernel_weight(d2v, g); sv = fadd(sv, fmul(w, __ldg(p.vy + q))); wv = fadd(wv, w); }
            }
            if (DOW) {
                const float d2w = fadd(fadd(bx2, by2), fmul(az, az));
                if (d2w < g.rsq) { float w = kernel_weight(d2w, g); sw = fadd(sw, fmul(w, __ldg(p.vz + q))); ww = fadd(ww, w); }
            }
            // SDF: all particles of the 3x3x3 neighbourhood are inside the reference's search box
            // (|c-p|_inf < 1.5dx < 2r+0.5dx); block origin is the same number for both subsystems
            if (DYADIC) d2c = fadd(fadd(bx2, by2), bz2);
            best2 = fminf(best2, d2c);
        }
    }
}

// Packed single precision (sm_100a FADD2/FMUL2/FFMA2: two particles per instruction).  ptxas contracts packed
// mul+add pairs into FFMA2 even with explicit .rn, so only quantities that tolerate it go through these.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// Constants of one node for gather_row_pk (both halves hold the same number).
struct PkNode {
    f32x2 Xn, Yn, Zn, Xh, Yh, Zh;    // node coordinates and node + dx/2
    f32x2 c1, c2, c3, one;           // Horner form of the weight: ((c1*d2 + c2)*d2 + c3)*d2 + 1
};

// The same row for power-of-two dx, two particles per instruction.  The kernel is bound by instruction issue
// (ncu: issue slots 80 %, DRAM 6 %); here the coordinate differences, the three squared face distances and the
// three weight polynomials of a particle PAIR are 23 packed instructions.  What must match the reference bit for
// bit stays scalar and literal: the squared distance to the cell centre (SDF minimum).  The packed face distances
// and the Horner weights differ from the literal ones by a few ulp: the support test can only differ where the
// weight is ~1e-7, the weights by <1e-6, i.e. float summation noise for the face value (tolerance 1e-5) -- except
// for the `valid = sum(w) > 1e-6` decision of a face whose total weight is that small, and the caller recomputes
// those faces with gather_row.  Velocities are loaded under the support predicate.
template <bool DOV, bool DOW>
__device__ __forceinline__ void gather_row_pk(const ParticleSoA &p, const GatherParams &g, const PkNode &N, int qb, int qe,
                                              float Xh, float Yh, float Zh, float &su, float &wu, float &sv, float &wv,
                                              float &sw, float &ww, float &best2) {
    const float rsq = g.rsq;
    for (int q4 = qb & ~3; q4 < qe; q4 += 4) {
        const float4 X4 = __ldg(reinterpret_cast<const float4 *>(p.px + q4));
        const float4 Y4 = __ldg(reinterpret_cast<const float4 *>(p.py + q4));
        const float4 Z4 = __ldg(reinterpret_cast<const float4 *>(p.pz + q4));
        // a particle outside [qb,qe) is moved far away: no support hit, no effect on the minimum
        const unsigned lo = (unsigned)max(qb - q4, 0), span = (unsigned)min(qe - q4, 4) - lo;
        const float xs[4] = {(0u - lo) < span ? X4.x : 1.0e18f, (1u - lo) < span ? X4.y : 1.0e18f,
                             (2u - lo) < span ? X4.z : 1.0e18f, (3u - lo) < span ? X4.w : 1.0e18f};
        const float ys[4] = {Y4.x, Y4.y, Y4.z, Y4.w}, zs[4] = {Z4.x, Z4.y, Z4.z, Z4.w};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const f32x2 x = pk2(xs[2 * h], xs[2 * h + 1]), y = pk2(ys[2 * h], ys[2 * h + 1]), z = pk2(zs[2 * h], zs[2 * h + 1]);
            const f32x2 ax = sub2(N.Xn, x), bx = sub2(N.Xh, x), by = sub2(N.Yh, y), bz = sub2(N.Zh, z);
            const f32x2 bz2 = mul2(bz, bz);
            const f32x2 d2u = fma2(ax, ax, fma2(by, by, bz2));
            const f32x2 wU = fma2(fma2(fma2(N.c1, d2u, N.c2), d2u, N.c3), d2u, N.one);
            f32x2 d2v = 0, wV = 0, d2w = 0, wW = 0;
            if (DOV) {
                const f32x2 ay = sub2(N.Yn, y);
                d2v = fma2(bx, bx, fma2(ay, ay, bz2));
                wV = fma2(fma2(fma2(N.c1, d2v, N.c2), d2v, N.c3), d2v, N.one);
            }
            if (DOW) {
                const f32x2 az = sub2(N.Zn, z);
                d2w = fma2(bx, bx, fma2(by, by, mul2(az, az)));
                wW = fma2(fma2(fma2(N.c1, d2w, N.c2), d2w, N.c3), d2w, N.one);
            }
            float du[2], wu2[2], dv[2], wv2[2], dw[2], ww2[2];
            upk2(d2u, du[0], du[1]); upk2(wU, wu2[0], wu2[1]);
            upk2(d2v, dv[0], dv[1]); upk2(wV, wv2[0], wv2[1]);
            upk2(d2w, dw[0], dw[1]); upk2(wW, ww2[0], ww2[1]);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int m = 2 * h + e, q = q4 + m;
                if (du[e] < rsq) { const float w = fmaxf(wu2[e], 1.0e-20f); su = fmaf(w, __ldg(p.vx + q), su); wu += w; }
                if (DOV) { if (dv[e] < rsq) { const float w = fmaxf(wv2[e], 1.0e-20f); sv = fmaf(w, __ldg(p.vy + q), sv); wv += w; } }
                if (DOW) { if (dw[e] < rsq) { const float w = fmaxf(ww2[e], 1.0e-20f); sw = fmaf(w, __ldg(p.vz + q), sw); ww += w; } }
                // SDF: literal (x*x + y*y) + z*z of the exact differences
                const float ex = fsub(Xh, xs[m]), ey = fsub(Yh, ys[m]), ez = fsub(Zh, zs[m]);
                best2 = fminf(best2, fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez)));
            }
        }
    }
}

template <bool DYADIC>
__global__ void __launch_bounds__(128) k_sdf_p2g(ParticleSoA p, const int *__restrict__ cellStart, GatherParams g,
                                                 float *__restrict__ U, float *__restrict__ V, float *__restrict__ W,
                                                 unsigned char *__restrict__ validU, unsigned char *__restrict__ validV,
                                                 unsigned char *__restrict__ validW, float *__restrict__ phiL,
                                                 const float *__restrict__ phiS, const unsigned char *__restrict__ occ,
                                                 int *__restrict__ farCells, float *__restrict__ farBest, int *farCount) {
    const int I = g.I, J = g.J, K = g.K;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y, k = blockIdx.z;      // k: local plane
    if (i > I) return;
    const int kg = k + g.kOff;               // global plane
    const bool hasU = (j < J && k < K);
    const bool hasV = (i < I && k < K);
    const bool hasW = (i < I && j < J);
    const bool hasC = (i < I && j < J && k < K);

    // Empty space (most of the box): no particle within the 5x5x5 cells around the node, per the 4^3-cell
    // occupancy blocks -> faces are 0 / invalid and phi keeps its "no particles" value.
    {
        const int oI = (I + 3) >> 2, oJ = (J + 3) >> 2, oK = (K + 3) >> 2;
        const int bi0 = max(i - 2, 0) >> 2, bi1 = min(min(i + 2, I - 1) >> 2, oI - 1);
        const int bj0 = max(j - 2, 0) >> 2, bj1 = min(min(j + 2, J - 1) >> 2, oJ - 1);
        const int bk0 = max(k - 2, 0) >> 2, bk1 = min(min(k + 2, K - 1) >> 2, oK - 1);
        bool any = false;
        for (int bk = bk0; bk <= bk1; bk++)
            for (int bj = bj0; bj <= bj1; bj++)
                for (int bi = bi0; bi <= bi1; bi++) any |= __ldg(occ + bi + oI * (bj + oJ * bk)) != 0;
        if (!any) {
            if (hasU) { long long idx = (long long)i + (long long)(I + 1) * (j + (long long)J * k); U[idx] = 0.0f; validU[idx] = 0; }
            if (hasV) { long long idx = (long long)i + (long long)I * (j + (long long)(J + 1) * k); V[idx] = 0.0f; validV[idx] = 0; }
            if (hasW) { long long idx = (long long)i + (long long)I * (j + (long long)J * k); W[idx] = 0.0f; validW[idx] = 0; }
            if (hasC) phiL[(long long)i + (long long)I * (j + (long long)J * k)] = g.maxDist;
            return;
        }
    }

    const NodeFrame f = node_frame(g, i, j, kg);
    float su = 0.f, wu = 0.f, sv = 0.f, wv = 0.f, sw = 0.f, ww = 0.f;
    float best2 = 3.0e38f;   // min squared distance cell centre <-> particle (block-local arithmetic)

    const int jlo = max(j - 1, 0), jhi = min(j + 1, J - 1);
    const int klo = max(k - 1, 0), khi = min(k + 1, K - 1);
    const int ilo = max(i - 1, 0), ihi = min(i + 1, I - 1);
    // DYADIC (dx a power of two): every subtraction of the block-local formulation is exact, so
    //   gpos - ((p - offset) - origin)  ==  (node + offset) - p   bit for bit,
    // and the cell-centre distance of the SDF is the fully staggered one.  Xn.. = global node coordinates.
    const float Xn = (float)dmul((double)(float)i, g.dx), Yn = (float)dmul((double)(float)j, g.dx),
                Zn = (float)dmul((double)(float)kg, g.dx);
    const float Xh = fadd(Xn, g.hw), Yh = fadd(Yn, g.hw), Zh = fadd(Zn, g.hw);
    PkNode N;
    N.Xn = pk2(Xn, Xn); N.Yn = pk2(Yn, Yn); N.Zn = pk2(Zn, Zn); N.Xh = pk2(Xh, Xh); N.Yh = pk2(Yh, Yh); N.Zh = pk2(Zh, Zh);
    N.c1 = pk2(-g.coef1, -g.coef1); N.c2 = pk2(g.coef2, g.coef2); N.c3 = pk2(-g.coef3, -g.coef3); N.one = pk2(1.0f, 1.0f);
    for (int ck = klo; ck <= khi; ck++) {
        for (int cj = jlo; cj <= jhi; cj++) {
            int rowBase = I * (cj + J * ck);
            const int qb = __ldg(cellStart + rowBase + ilo);
            const int qe = __ldg(cellStart + rowBase + ihi + 1);
            if (qb == qe) continue;
            const bool doV = (cj <= j), doW = (ck <= k);     // uniform over the block
            if (DYADIC && g.packed) {
                if (doV && doW) gather_row_pk<true, true>(p, g, N, qb, qe, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
                else if (doV) gather_row_pk<true, false>(p, g, N, qb, qe, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
                else if (doW) gather_row_pk<false, true>(p, g, N, qb, qe, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
                else gather_row_pk<false, false>(p, g, N, qb, qe, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
                continue;
            }
            if (doV && doW) gather_row<DYADIC, true, true>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
            else if (doV) gather_row<DYADIC, true, false>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
            else if (doW) gather_row<DYADIC, false, true>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
            else gather_row<DYADIC, false, false>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, best2);
        }
    }

    if (DYADIC && g.packed) {
        // faces whose total weight is within the rounding band of the validity threshold (there were hits, yet the
        // sum is tiny): redo with the literal weights so that the valid mask is the reference's.  Practically never.
        const float band = 6.0e-5f;
        if ((wu > 0.0f && wu < band) || (wv > 0.0f && wv < band) || (ww > 0.0f && ww < band)) {
            su = wu = sv = wv = sw = ww = 0.0f;
            float b2 = 3.0e38f;
            for (int ck = klo; ck <= khi; ck++) {
                for (int cj = jlo; cj <= jhi; cj++) {
                    int rowBase = I * (cj + J * ck);
                    const int qb = __ldg(cellStart + rowBase + ilo);
                    const int qe = __ldg(cellStart + rowBase + ihi + 1);
                    if (qb == qe) continue;
                    if (cj <= j && ck <= k) gather_row<true, true, true>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, b2);
                    else if (cj <= j) gather_row<true, true, false>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, b2);
                    else if (ck <= k) gather_row<true, false, true>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, b2);
                    else gather_row<true, false, false>(p, g, f, qb, qe, Xn, Yn, Zn, Xh, Yh, Zh, su, wu, sv, wv, sw, ww, b2);
                }
            }
        }
    }

    if (hasU) {
        long long idx = (long long)i + (long long)(I + 1) * (j + (long long)J * k);
        bool ok = wu > g.eps;
        U[idx] = ok ? __fdiv_rn(su, wu) : su;     // scalar /= weight only if weight > eps (:448-453)
        validU[idx] = ok ? 1 : 0;
    }
    if (hasV) {
        long long idx = (long long)i + (long long)I * (j + (long long)(J + 1) * k);
        bool ok = wv > g.eps;
        V[idx] = ok ? __fdiv_rn(sv, wv) : sv;
        validV[idx] = ok ? 1 : 0;
    }
    if (hasW) {
        long long idx = (long long)i + (long long)I * (j + (long long)J * k);
        bool ok = ww > g.eps;
        W[idx] = ok ? __fdiv_rn(sw, ww) : sw;
        validW[idx] = ok ? 1 : 0;
    }
    if (!hasC) return;

    // Particles outside the 3x3x3 neighbourhood can only matter if nothing nearer than 1.5dx was found (they are
    // at least 1.5dx away along one axis).  Those few cells (a shell around the liquid) are queued for
    // k_sdf_far, which searches the 5x5x5 box with all lanes busy.
    const float thr = fmul(fmul(1.45f, (float)g.dx), fmul(1.45f, (float)g.dx));
    const int cell = i + I * (j + J * k);
    if (!(best2 < thr)) {
        int slot = warp_append_slot(farCount);
        farCells[slot] = cell;
        farBest[slot] = best2;
        return;
    }
    phiL[cell] = finish_phi(g, phiS, best2, i, j, k);
}

// Liquid SDF of the queued cells: the whole 5x5x5 search box (particlelevelset.cpp:476-513, :596-621).
// ONE WARP PER CELL.  (One thread per cell walks its rows through chains of dependent loads, ~0.4 us of L2 latency
// per particle: measured 209 us for ~1e5 queued cells, nearly all of it the slowest thread's chain.)  Lane r < 25
// fetches the particle range of row r of the 5x5 rows around the cell, all at once; the rows are then taken nearest
// first -- a row whose nearest point is farther than the best particle so far ends the search -- with the lanes on
// consecutive particles of the row (coalesced), and the lane minima are folded after every row.
// A particle of a cell two away along some axis lies inside the reference's search box for roughly 3/4 of that
// cell; the decision is taken in float with a margin far above the rounding of the reference's block-local
// arithmetic, and only the borderline cases run the literal double-precision index computation.
// the 25 (dj,dk) row offsets ordered by the distance of the row from the cell, packed as r = (dj+2) + 5 (dk+2)
__constant__ unsigned char c_far_order[25] = {12, 7, 11, 13, 17, 6, 8, 16, 18, 2, 10, 14, 22, 1, 3, 5, 9, 15, 19, 21, 23, 0, 4, 20, 24};

__global__ void __launch_bounds__(256) k_sdf_far(ParticleSoA p, const int *__restrict__ cellStart, GatherParams g,
                                                 float *__restrict__ phiL, const float *__restrict__ phiS,
                                                 const int *__restrict__ farCells, const float *__restrict__ farBest,
                                                 const int *__restrict__ farCount) {
    const int I = g.I, J = g.J, K = g.K;
    const int n = *farCount;
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const float sr = g.srS;
    const double invdx = g.invdx;
    const float dxf = (float)g.dx, margin = 1.0e-3f * dxf;
    const float loIn = -sr + margin, hiIn = dxf + sr - margin, loOut = -sr - margin, hiOut = dxf + sr + margin;
    // q2[d]: squared distance (in dx^2) from the centre to the nearest point of the cells at offset d along one axis,
    // less a margin against the rounding of the computed distances
    const float dx2 = fmul(dxf, dxf);
    auto q2 = [&](int d) { return d == 0 ? 0.0f : (d == 1 || d == -1) ? 0.2499f * dx2 : 2.2499f * dx2; };
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n; t += nw) {
        const int cell = farCells[t];
        float best2 = farBest[t];
        const int i = cell % I, j = (cell / I) % J, k = cell / (I * J);
        const NodeFrame f = node_frame(g, i, j, k + g.kOff);
        const float Xn = (float)dmul((double)(float)i, g.dx), Yn = (float)dmul((double)(float)j, g.dx),
                    Zn = (float)dmul((double)(float)(k + g.kOff), g.dx);
        const int i2lo = max(i - 2, 0), i2hi = min(i + 2, I - 1);
        int qb = 0, qe = 0;
        if (lane < 25) {
            const int cj = j + (lane % 5) - 2, ck = k + (lane / 5) - 2;
            if (cj >= 0 && cj < J && ck >= 0 && ck < K) {
                const int rowBase = I * (cj + J * ck);
                qb = __ldg(cellStart + rowBase + i2lo);
                qe = __ldg(cellStart + rowBase + i2hi + 1);
            }
        }
        const unsigned int rows = __ballot_sync(0xffffffffu, qe > qb);
        for (int o = 0; o < 25; o++) {
            const int r = c_far_order[o];
            const int dj = (r % 5) - 2, dk = (r / 5) - 2;
            if (!(q2(dj) + q2(dk) < best2)) break;       // the rows that follow are no nearer
            if (!((rows >> r) & 1u)) continue;
            const int rb = __shfl_sync(0xffffffffu, qb, r), re = __shfl_sync(0xffffffffu, qe, r);
            float mine = best2;
            for (int q = rb + lane; q < re; q += 32) {
                const float x = __ldg(p.px + q), y = __ldg(p.py + q), z = __ldg(p.pz + q);
                // position relative to the cell's lower corner: inside the search box iff -sr <= u < dx + sr
                const float ux = fsub(x, Xn), uy = fsub(y, Yn), uz = fsub(z, Zn);
                const bool out = ux < loOut || ux >= hiOut || uy < loOut || uy >= hiOut || uz < loOut || uz >= hiOut;
                if (out) continue;
                const bool in = ux >= loIn && ux < hiIn && uy >= loIn && uy < hiIn && uz >= loIn && uz < hiIn;
                const float xl = fsub(x, f.ox), yl = fsub(y, f.oy), zl = fsub(z, f.oz);
                if (!in) {
                    // borderline: the literal test.  Block membership of the particle: the blocks overlapped by
                    // [p-sr, p+sr] in global coordinates; then the block-local search box.
                    int bminx = pos2idx_d((double)fsub(x, sr), g.invBlockdxSDF), bmaxx = pos2idx_d((double)fadd(x, sr), g.invBlockdxSDF);
                    int bminy = pos2idx_d((double)fsub(y, sr), g.invBlockdxSDF), bmaxy = pos2idx_d((double)fadd(y, sr), g.invBlockdxSDF);
                    int bminz = pos2idx_d((double)fsub(z, sr), g.invBlockdxSDF), bmaxz = pos2idx_d((double)fadd(z, sr), g.invBlockdxSDF);
                    if (f.bi < bminx || f.bi > bmaxx || f.bj < bminy || f.bj > bmaxy || f.bk < bminz || f.bk > bmaxz) continue;
                    int gminx = pos2idx(fsub(xl, sr), invdx), gmaxx = pos2idx(fadd(xl, sr), invdx);
                    int gminy = pos2idx(fsub(yl, sr), invdx), gmaxy = pos2idx(fadd(yl, sr), invdx);
                    int gminz = pos2idx(fsub(zl, sr), invdx), gmaxz = pos2idx(fadd(zl, sr), invdx);
                    if (f.li < gminx || f.li > gmaxx || f.lj < gminy || f.lj > gmaxy || f.lk < gminz || f.lk > gmaxz) continue;
                }
                mine = fminf(mine, lengthsq3(fsub(f.cx, xl), fsub(f.cy, yl), fsub(f.cz, zl)));
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) mine = fminf(mine, __shfl_xor_sync(0xffffffffu, mine, s));
            best2 = mine;
        }
        if (lane == 0) phiL[cell] = finish_phi(g, phiS, best2, i, j, k);
    }
}

// ------------------------------------------------------------------------------------------------
// P2G + liquid SDF as a shared-memory SCATTER (default for power-of-two dx in FLIP_SAMPLING_FAST; the gather above
// stays the literal path of FLIP_SAMPLING_EXACT and of non-dyadic cell widths).
//
// A particle reaches, per component, at most the 2x2x2 faces around it (r = 0.866 dx < dx): 24 candidate faces,
// ~8 hits, against ~650 (particle, face) evaluations per node of the 3x3x3-cell gather.  What a scatter needs is an
// accumulation that many threads can add to; shared-memory float atomics are CAS loops on sm_100a and would make
// the sums depend on the arrival order, so the sums are kept in FIXED POINT with native 32-bit integer atomics
// (ATOMS.ADD; measured: 0.34 ms for the whole particle set against 0.81 ms with float CAS atomics and 1.5 ms with
// global REDs, scripts/micro/atomics_bench.cu): integer addition is associative, hence the result is bit-reproducible
// from run to run whatever the order, inside a tile and across tiles.  Per face and component TWO words:
//     W = sum wq,  wq = round(w * 2^22)       u32: no overflow below sum(w) = 1024, i.e. 56 particles in each of the
//                                             18 cells that can reach a face (denser cells: see below)
//     S = sum round(wq v 2^(q-22))            i32: 2^q = the largest power of two with vmax 2^q <= 2^21, vmax = the
//                                             largest magnitude of that velocity component among the TILE's particles
//                                             (a pass over the tile's velocities first; a few fast particles elsewhere
//                                             in the domain cost this tile no digits)
// in TWO TIERS: a contribution of weight below 1/16 goes to a second pair of words scaled 16 times finer (same bounds:
// the terms are 16 times smaller).  S is formed from the ROUNDED weight, so u = S/W is a weighted mean with weights
// wq/2^22 (2^26): the rounding of the weights only enters through the velocity differences of the contributors, and
// the momentum terms are rounded at vmax 2^-22 -- vmax 2^-26 on the faces at the rim of the liquid, whose
// contributions are all small: < 3e-6 vmax on a face of total weight 0.02, ~1e-7 of the face value in the bulk.
// (Measured on the B200: the kernel is bound by the shared-memory atomic unit, ~4 cycles per warp-wide ATOMS; a third
// word per component -- the same precision from one scale -- costs 682 us against 411 us for two.)
// The liquid SDF rides along: a particle also takes the minimum of its squared distance to the 2x2x2 cell centres
// around it (literal arithmetic: (x*x + y*y) + z*z of the exact differences; ATOMS.MAX on the inverted float bits),
// where that distance is below dx -- every particle NOT among those a centre receives is at least dx away from it
// along some axis, so a minimum below dx is final.  (That settles every liquid cell and most of the first air layer;
// the other cells within reach of a particle are queued for k_sdf_far.)
// One CTA owns the particles of a tile of 8x8x8 cells, spread evenly over its threads (thread t takes M consecutive
// particles of the tile's 64 rows laid end to end: a thread per cell leaves 44 % of the lanes idle, E[max of 32
// Poisson(8)] = 14; and neighbouring lanes M particles apart work on different cells, so their atomics rarely
// collide), and accumulates into the (8+2)^3 slots around it; the tile's partial sums are then added to global
// accumulators with one 64-bit RED per face (W in the low word, S in the high word: W cannot carry) where the
// contributions of neighbouring tiles meet, and k_p2g_finish turns the sums into u = S/W and valid = W > 1e-6 (and
// clears them).  Faces whose total weight is below P2G_LITERAL_BAND -- few particles near the rim of the support:
// the fixed-point weight has too few significant digits there, and the valid-mask decision sum(w) > 1e-6 lies in
// that range -- are queued and recomputed by k_p2g_literal with the literal weights in the fixed particle order of
// the gather, so those values and the valid mask are what the gather produces.
// ------------------------------------------------------------------------------------------------
static constexpr int P2G_T = 8;                       // tile edge, cells
static constexpr int P2G_F = P2G_T + 2;               // slots per axis
static constexpr int P2G_SLOTS = P2G_F * P2G_F * P2G_F;
static constexpr int P2G_WORDS = 13;                  // per slot: (W, S, W fine, S fine) of U, V, W; ~min squared distance
static constexpr int P2G_THREADS = 512;
static constexpr int P2G_ROWS = P2G_T * P2G_T;
static constexpr float P2G_LITERAL_BAND = 0.02f;
static constexpr float P2G_WSCALE = 4194304.0f;       // 2^22
static constexpr int P2G_MAX_PER_CELL = 56;           // 18 cells x 56 particles x (w <= 1) < 1024: W and S cannot overflow;
                                                      // the particles of a denser cell stay out of the sums and the 81 faces
                                                      // around it are recomputed by k_p2g_literal

// ---- occupancy bitmaps: one bit per cell, 32 cells of a row per word.  occ: the cell holds particles; near3 / near5:
// some cell of the 3x3x3 / 5x5x5 neighbourhood does.  A face of node (i,j,k) can only receive weight from the 3x3x3
// cells around cell (i,j,k), and a cell whose 5x5x5 neighbourhood is empty lies in no particle's SDF search box
// (particlelevelset.cpp:476-513: a box reaches at most two cells from the particle's cell).
__global__ void k_occ_bits(const int *__restrict__ cellStart, int I, int rows, int WR, unsigned int *__restrict__ bits) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= rows * WR) return;
    const int row = w / WR, i = (w % WR) * 32 + lane;
    bool on = false;
    if (i < I) { const int c = row * I + i; on = __ldg(cellStart + c + 1) > __ldg(cellStart + c); }
    const unsigned int m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[w] = m;
}
// surf: the cell holds particles and some cell of its 5x5x5 neighbourhood (inside the grid) holds none
__global__ void k_occ_dilate(const unsigned int *__restrict__ bits, int I, int J, int K, int WR, unsigned int *__restrict__ near3,
                             unsigned int *__restrict__ near5, unsigned int *__restrict__ surf, int *__restrict__ tileFlags) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= J * K * WR) return;
    const int wr = w % WR, j = (w / WR) % J, k = w / (WR * J);
    // cells of the row that exist, per word
    auto exist = [&](int q) { return (q < 0 || q >= WR) ? 0u : ((q + 1) * 32 <= I ? 0xffffffffu : ((1u << (I - q * 32)) - 1u)); };
    const unsigned int xp = exist(wr - 1), xc = exist(wr), xn = exist(wr + 1);
    unsigned int p3 = 0, c3 = 0, n3 = 0, p5 = 0, c5 = 0, n5 = 0, pe = 0, ce = 0, ne = 0;
    for (int dk = -2; dk <= 2; dk++) {
        const int kk = k + dk;
        if (kk < 0 || kk >= K) continue;
        for (int dj = -2; dj <= 2; dj++) {
            const int jj = j + dj;
            if (jj < 0 || jj >= J) continue;
            const unsigned int *r = bits + (size_t)WR * (jj + (size_t)J * kk);
            const unsigned int pv = wr > 0 ? __ldg(r + wr - 1) : 0u, cv = __ldg(r + wr), nv = wr + 1 < WR ? __ldg(r + wr + 1) : 0u;
            p5 |= pv; c5 |= cv; n5 |= nv;
            pe |= ~pv & xp; ce |= ~cv & xc; ne |= ~nv & xn;
            if (dk >= -1 && dk <= 1 && dj >= -1 && dj <= 1) { p3 |= pv; c3 |= cv; n3 |= nv; }
        }
    }
    // along the row: bit i of the result looks at bits i-d .. i+d across the word boundaries
    const unsigned int d3 = c3 | (c3 << 1) | (p3 >> 31) | (c3 >> 1) | (n3 << 31);
    const unsigned int d5 = c5 | (c5 << 1) | (p5 >> 31) | (c5 >> 1) | (n5 << 31) | (c5 << 2) | (p5 >> 30) | (c5 >> 2) | (n5 << 30);
    const unsigned int e5 = ce | (ce << 1) | (pe >> 31) | (ce >> 1) | (ne << 31) | (ce << 2) | (pe >> 30) | (ce >> 2) | (ne << 30);
    near3[w] = d3;
    near5[w] = d5;
    const unsigned int own = __ldg(bits + w);
    surf[w] = own & e5;
    // the 8x8x8-cell tiles this word touches: holds particles (P2G scatter) / holds empty cells within reach of a
    // particle (SDF shell search).  Everybody stores the same 1.
    const unsigned int shell = d5 & ~own & xc;
    const int tX = (I + 7) >> 3, tY = (J + 7) >> 3;
    const int tBase = tX * ((j >> 3) + tY * (k >> 3)) + wr * 4;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if ((own >> (8 * b)) & 0xffu) tileFlags[2 * (tBase + b)] = 1;
        if ((shell >> (8 * b)) & 0xffu) tileFlags[2 * (tBase + b) + 1] = 1;
    }
}

// flags -> lists of tiles (and the flags are cleared for the next step); counts[0/1]: list lengths
__global__ void k_tile_lists(int nTiles, int *__restrict__ tileFlags, int *__restrict__ listP, int *__restrict__ listS,
                             int *__restrict__ counts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTiles) return;
    const int2 f = reinterpret_cast<int2 *>(tileFlags)[t];
    if (f.x | f.y) reinterpret_cast<int2 *>(tileFlags)[t] = make_int2(0, 0);
    if (f.x) listP[warp_append_slot(&counts[0])] = t;
    if (f.y) listS[warp_append_slot(&counts[1])] = t;
}

__device__ __forceinline__ unsigned int smem_addr(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

// round-to-nearest of |t| < 2^22 through the mantissa: t + 1.5*2^23 has the integer in its low bits (one FFMA
// instead of a conversion)
#define P2G_MAGIC 12582912.0f
#define P2G_MAGIC_BITS 0x4B400000

// One candidate face: weight (Horner form of velocityadvector.cpp:437; zero outside the support), the two
// fixed-point words, two REDs.  The REDs are unconditional -- a miss adds zeros: ptxas turns a predicated ATOMS into a
// branch around it with a convergence barrier (BSSY / BRA / BSYNC, three issue slots per atomic; some lane of a warp
// hits nearly every candidate anyway).  vs = v * 2^(q-22).
template <int OFF>
__device__ __forceinline__ void p2g_candidate(float d2, float rsq, float c1, float c2, float c3, float vs, unsigned int addr) {
    float w = fmaf(fmaf(fmaf(c1, d2, c2), d2, c3), d2, 1.0f);
    // (the polynomial is a few 1e-8 below zero at the very rim of the support; the sums are unsigned)
    w = d2 < rsq ? fmaxf(w, 0.0f) : 0.0f;
    const bool fine = w < 0.0625f;
    const float t = fmaf(w, fine ? 16.0f * P2G_WSCALE : P2G_WSCALE, P2G_MAGIC);
    const int iw = __float_as_int(t) - P2G_MAGIC_BITS;
    const int is = __float_as_int(fmaf(t - P2G_MAGIC, vs, P2G_MAGIC)) - P2G_MAGIC_BITS;      // t - magic = (float)iw, exact
    const unsigned int a = addr + (fine ? 8u : 0u);
    asm volatile("red.shared.add.u32 [%0+%3], %1;\n\tred.shared.add.u32 [%0+%4], %2;"
                 :: "r"(a), "r"(iw), "r"(is), "n"(OFF), "n"(OFF + 4) : "memory");
}

// the 2x2x2 faces of component COMP around a particle; s0/s1/s2: squared offsets to the two planes along x/y/z
template <int COMP>
__device__ __forceinline__ void p2g_component(unsigned int accAddr, int fi, int fj, int fk, const float *s0, const float *s1,
                                              const float *s2, float vs, float rsq, float c1, float c2, float c3) {
    const unsigned int a = accAddr + (unsigned int)((fi + P2G_F * (fj + P2G_F * fk)) * (P2G_WORDS * 4) + COMP * 16);
#define P2G_CAND(DI, DJ, DK) \
    p2g_candidate<((DI) + P2G_F * ((DJ) + P2G_F * (DK))) * (P2G_WORDS * 4)>(s0[DI] + (s1[DJ] + s2[DK]), rsq, c1, c2, c3, vs, a);
    P2G_CAND(0, 0, 0) P2G_CAND(1, 0, 0) P2G_CAND(0, 1, 0) P2G_CAND(1, 1, 0)
    P2G_CAND(0, 0, 1) P2G_CAND(1, 0, 1) P2G_CAND(0, 1, 1) P2G_CAND(1, 1, 1)
#undef P2G_CAND
}

// squared distance to one of the 2x2x2 cell centres, literal ((x*x + y*y) + z*z); kept where below `thr`
template <int OFF>
__device__ __forceinline__ void sdf_candidate(float xy, float zz, float thr, unsigned int addr) {
    const float d2 = fadd(xy, zz);
    // non-negative floats order as integers: max of ~bits = min; 0 = nothing
    const unsigned int inv = d2 < thr ? ~__float_as_uint(d2) : 0u;
    asm volatile("red.shared.max.u32 [%0+%2], %1;" :: "r"(addr), "r"(inv), "n"(OFF) : "memory");
}

struct P2GGlobal {
    ulonglong2 *accU, *accV, *accW;           // per face: see below
    unsigned int *minC;                       // per cell: ~bits of the minimum squared distance, 0 = none
};
// global accumulators of a face: x = sum of weights in units of 2^-26, y = sum of momenta in units of 2^-(qg+24), qg:
// the scale exponent that the largest speed of the whole domain would get (every tile's own exponent lies in
// [qg, qg+20])
struct SdfQueues {
    int *farCells; float *farBest; int *farCount;   // (gather path only) no particle within 1.45 dx: the 5x5x5 box (k_sdf_far)
    int *litFaces, *litCount; int litCap;           // faces below P2G_LITERAL_BAND or next to a dense cell (k_p2g_literal)
};

__global__ void __launch_bounds__(P2G_THREADS, 3) k_p2g_scatter(ParticleSoA p, const int *__restrict__ cellStart, GatherParams g,
                                                               P2GGlobal G, int tilesX, int tilesY, int qg, SdfQueues Q,
                                                               const int *__restrict__ tileList, int *__restrict__ tileCounts) {
    extern __shared__ int acc[];         // [slot][P2G_WORDS]
    __shared__ int rowBeg[P2G_ROWS], rowPre[P2G_ROWS + 1];
    __shared__ unsigned int rowDense[P2G_ROWS];    // bit i: cell i of the row is too dense for the fixed-point sums
    __shared__ int nextTile;
    __shared__ float vmaxWarp[P2G_THREADS / 32][3];
    __shared__ int qTile[3];
    const int tid = threadIdx.x, lane = tid & 31;
    const int I = g.I, J = g.J, K = g.K;
    const float dxf = (float)g.dx, invf = (float)g.invdx, hw = g.hw, rsq = g.rsq;
    const float c1 = -g.coef1, c2 = g.coef2, c3 = -g.coef3;
    const float thr = 0.999f * dxf * dxf;
    const unsigned int accAddr = smem_addr(acc);
    // persistent CTAs take the tiles that hold particles (most tiles of the box are air) from the list, one ticket each
    const int nTiles = tileCounts[0];
    for (;;) {
    __syncthreads();                                   // the previous tile's flush has read acc / the row tables
    if (tid == 0) nextTile = atomicAdd(&tileCounts[2], 1);
    __syncthreads();
    if (nextTile >= nTiles) break;
    const int tile = tileList[nextTile];
    const int tx = tile % tilesX, ty = (tile / tilesX) % tilesY, tz = tile / (tilesX * tilesY);
    const int i0 = tx * P2G_T, j0 = ty * P2G_T, k0 = tz * P2G_T;     // first cell of the tile (local planes)
    // the tile's 64 rows of 8 cells: particle ranges and their running total (two warps, shuffle scan)
    if (tid < P2G_ROWS) {
        const int j = j0 + (tid % P2G_T), k = k0 + tid / P2G_T;
        int b = 0, n = 0;
        unsigned int dense = 0u;
        if (j < J && k < K) {
            const int c = i0 + I * (j + J * k);
            const int nc = min(P2G_T, I - i0);
            int prev = __ldg(cellStart + c);
            b = prev;
#pragma unroll
            for (int m = 1; m <= P2G_T; m++) {
                if (m <= nc) {
                    const int nx = __ldg(cellStart + c + m);
                    if (nx - prev > P2G_MAX_PER_CELL) dense |= 1u << (m - 1);
                    prev = nx;
                }
            }
            n = prev - b;
            // the faces that a dense cell's particles can reach: the three faces of the 3x3x3 nodes from the cell's own
            for (unsigned int dm = dense; dm; dm &= dm - 1u) {
                const int ci = i0 + __ffs(dm) - 1;
                const int slot = atomicAdd(Q.litCount, 81);
                int w = 0;
                for (int dk = 0; dk <= 2; dk++)
                    for (int dj = 0; dj <= 2; dj++)
                        for (int di = 0; di <= 2; di++) {
                            const int ni = ci - 1 + di, nj = j - 1 + dj, nk = k - 1 + dk;
                            // (nodes outside the grid: the cell's own node instead, a harmless duplicate)
                            const bool in = ni >= 0 && nj >= 0 && nk >= 0 && ni <= I && nj <= J && nk <= K;
                            const int node = in ? ni + (I + 1) * (nj + (J + 1) * nk) : ci + (I + 1) * (j + (J + 1) * k);
                            for (int comp = 0; comp < 3; comp++, w++)
                                if (slot + w < Q.litCap) Q.litFaces[slot + w] = (int)(((unsigned int)comp << 30) | (unsigned int)node);
                        }
            }
        }
        rowBeg[tid] = b;
        rowDense[tid] = dense;
        int v = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        rowPre[tid + 1] = v;     // inclusive within the warp
    }
    for (int q = tid; q < P2G_SLOTS * P2G_WORDS; q += P2G_THREADS) acc[q] = 0;
    __syncthreads();
    if (tid >= 32 && tid < P2G_ROWS) rowPre[tid + 1] += rowPre[32];
    if (tid == 0) rowPre[0] = 0;
    __syncthreads();
    const int total = rowPre[P2G_ROWS];
    const float org[3] = {(float)(i0 - 1), (float)(j0 - 1), (float)(k0 - 1) + (float)g.kOff};   // z-slab: local planes
    // thread t: the M consecutive particles [t M, (t+1) M) of the rows laid end to end
    const int M = (total + P2G_THREADS - 1) / P2G_THREADS;
    int idx = tid * M;
    const int idxEnd = min(idx + M, total);
    int row = 0;
    if (idx < total) {
#pragma unroll
        for (int s = P2G_ROWS / 2; s > 0; s >>= 1)
            if (rowPre[row + s] <= idx) row += s;
    }
    // ---- the momentum scales of this tile: largest |vx|, |vy|, |vz| of its particles
    {
        // (warp w reads the rows w, w + 16, ...: coalesced, and the lines are in L1 when the scatter below asks for them)
        float mx = 0.0f, my = 0.0f, mz = 0.0f;
        for (int r2 = tid >> 5; r2 < P2G_ROWS; r2 += P2G_THREADS / 32) {
            const int qb = rowBeg[r2], qe = qb + (rowPre[r2 + 1] - rowPre[r2]);
            for (int q = qb + lane; q < qe; q += 32) {
                mx = fmaxf(mx, fabsf(__ldg(p.vx + q))); my = fmaxf(my, fabsf(__ldg(p.vy + q))); mz = fmaxf(mz, fabsf(__ldg(p.vz + q)));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); my = fmaxf(my, __shfl_xor_sync(0xffffffffu, my, o));
            mz = fmaxf(mz, __shfl_xor_sync(0xffffffffu, mz, o));
        }
        if (lane == 0) { vmaxWarp[tid >> 5][0] = mx; vmaxWarp[tid >> 5][1] = my; vmaxWarp[tid >> 5][2] = mz; }
        __syncthreads();
        if (tid < 3) {
            float m = 0.0f;
            for (int w = 0; w < P2G_THREADS / 32; w++) m = fmaxf(m, vmaxWarp[w][tid]);
            // m < 2^e  =>  m 2^(21-e) < 2^21
            const int e = (m > 0.0f && m < 3.0e38f) ? ilogbf(m) + 1 : -1000;
            qTile[tid] = min(max(21 - e, qg), qg + 20);
        }
        __syncthreads();
    }
    const int qx = qTile[0], qy = qTile[1], qz = qTile[2];
    const float sScaleX = __int_as_float((qx - 22 + 127) << 23), sScaleY = __int_as_float((qy - 22 + 127) << 23),
                sScaleZ = __int_as_float((qz - 22 + 127) << 23);
    for (; idx < idxEnd; idx++) {
        while (idx >= rowPre[row + 1]) row++;
        const int q = rowBeg[row] + (idx - rowPre[row]);
        const float pos[3] = {__ldg(p.px + q), __ldg(p.py + q), __ldg(p.pz + q)};
        // v 2^(q-22), at most 1/2 in magnitude (the clamp only matters if qg was computed from too small a speed)
        const float vsx = fminf(fmaxf(__ldg(p.vx + q) * sScaleX, -1.0f), 1.0f), vsy = fminf(fmaxf(__ldg(p.vy + q) * sScaleY, -1.0f), 1.0f),
                    vsz = fminf(fmaxf(__ldg(p.vz + q) * sScaleZ, -1.0f), 1.0f);
        // per axis: the planes below / above the particle (a: through the nodes, b: through the cell centres), the
        // offsets to them -- each ONE rounding of the exact difference, as the literal fsub(plane, p) -- and their
        // squares.  dx is a power of two and the indices are far below 2^24: plane coordinates are exact.
        int ia[3], ib[3];
        float sqa[3][2], sqb[3][2];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const float u = pos[ax] * invf;
            const float fa = floorf(u), fb = floorf(u - 0.5f);
            const float ca = fa * dxf, cb = fb * dxf + hw;
            const float da0 = fsub(pos[ax], ca), da1 = fsub(pos[ax], ca + dxf);
            const float db0 = fsub(pos[ax], cb), db1 = fsub(pos[ax], cb + dxf);
            sqa[ax][0] = fmul(da0, da0); sqa[ax][1] = fmul(da1, da1);
            sqb[ax][0] = fmul(db0, db0); sqb[ax][1] = fmul(db1, db1);
            ia[ax] = (int)(fa - org[ax]); ib[ax] = (int)(fb - org[ax]);    // slot coordinates inside the (8+2)^3 block
        }
        // component c is staggered on the two other axes
        if (!((rowDense[row] >> (ia[0] - 1)) & 1u)) {
            p2g_component<0>(accAddr, ia[0], ib[1], ib[2], sqa[0], sqb[1], sqb[2], vsx, rsq, c1, c2, c3);
            p2g_component<1>(accAddr, ib[0], ia[1], ib[2], sqb[0], sqa[1], sqb[2], vsy, rsq, c1, c2, c3);
            p2g_component<2>(accAddr, ib[0], ib[1], ia[2], sqb[0], sqb[1], sqa[2], vsz, rsq, c1, c2, c3);
        }
        // liquid SDF: the 2x2x2 cell centres
        {
            const unsigned int a = accAddr + (unsigned int)((ib[0] + P2G_F * (ib[1] + P2G_F * ib[2])) * (P2G_WORDS * 4) + 48);
            const float xy00 = fadd(sqb[0][0], sqb[1][0]), xy10 = fadd(sqb[0][1], sqb[1][0]);
            const float xy01 = fadd(sqb[0][0], sqb[1][1]), xy11 = fadd(sqb[0][1], sqb[1][1]);
#define SDF_CAND(DI, DJ, DK, XY) sdf_candidate<((DI) + P2G_F * ((DJ) + P2G_F * (DK))) * (P2G_WORDS * 4)>(XY, sqb[2][DK], thr, a);
            SDF_CAND(0, 0, 0, xy00) SDF_CAND(1, 0, 0, xy10) SDF_CAND(0, 1, 0, xy01) SDF_CAND(1, 1, 0, xy11)
            SDF_CAND(0, 0, 1, xy00) SDF_CAND(1, 0, 1, xy10) SDF_CAND(0, 1, 1, xy01) SDF_CAND(1, 1, 1, xy11)
#undef SDF_CAND
        }
    }
    __syncthreads();
    // the tile's partial sums -> global accumulators (faces outside the grid cannot exist in the reference's loops)
    for (int s = tid; s < P2G_SLOTS; s += P2G_THREADS) {
        const int *a = acc + s * P2G_WORDS;
        const int fi = i0 - 1 + s % P2G_F, fj = j0 - 1 + (s / P2G_F) % P2G_F, fk = k0 - 1 + s / (P2G_F * P2G_F);
        if (fi < 0 || fj < 0 || fk < 0) continue;
        const unsigned int mn = a[12];
        auto flush = [&](int comp, ulonglong2 *dst, size_t idx) {
            const unsigned int wc = a[4 * comp], wf = a[4 * comp + 2];
            const int sc = a[4 * comp + 1], sf = a[4 * comp + 3];
            if (!(wc | wf | (unsigned int)sc | (unsigned int)sf)) return;
            const int qt = comp == 0 ? qx : comp == 1 ? qy : qz;
            atomicAdd(&dst[idx].x, (unsigned long long)wc * 16ull + wf);
            atomicAdd(&dst[idx].y, (unsigned long long)(((long long)sc * 16 + (long long)sf) * (1ll << (qg + 20 - qt))));
        };
        if (fi <= I && fj < J && fk < K) flush(0, G.accU, (size_t)fi + (size_t)(I + 1) * ((size_t)fj + (size_t)J * (size_t)fk));
        if (fi < I && fj <= J && fk < K) flush(1, G.accV, (size_t)fi + (size_t)I * ((size_t)fj + (size_t)(J + 1) * (size_t)fk));
        if (fi < I && fj < J && fk <= K) flush(2, G.accW, (size_t)fi + (size_t)I * ((size_t)fj + (size_t)J * (size_t)fk));
        if (mn && fi < I && fj < J && fk < K) atomicMax(G.minC + ((size_t)fi + (size_t)I * ((size_t)fj + (size_t)J * (size_t)fk)), mn);
    }
    }   // tiles
}

// ---- liquid SDF of the cells the 2x2x2 scatter cannot settle (no particle within dx of the centre).  Such a cell is
// EMPTY (a particle of the cell itself is within 0.87 dx of its centre), so the particles whose search box
// (particlelevelset.cpp:476-513: at most two cells from the particle's cell) contains it lie in cells that have an
// empty cell in their 5x5x5 neighbourhood: the SURFACE cells (bitmap from k_occ_dilate) -- a shell two cells thick
// along the free surface and the walls.  One CTA per tile of 8x8x8 cells that holds unsettled cells: the particle
// positions of the surface cells of the (8+4)^3 region go to shared memory once (a shell through the region: a few
// thousand particles), then one thread per unsettled cell searches the 5x5x5 cells around it nearest first and stops
// at the first cell that cannot hold anything nearer -- the dependent loads of the search cost a shared-memory
// access instead of an L2 round trip (measured for ~4e5 such cells: 431 us with a warp per cell on global memory).
// The box test is taken in float with a margin far above the rounding of the reference's block-local arithmetic;
// only the borderline cases run the literal double-precision index computation.  A region with more surface
// particles than the staging area holds sends its cells to k_sdf_far instead.
static constexpr int SHELL_F = P2G_T + 4;
static constexpr int SHELL_ROWS = SHELL_F * SHELL_F;
static constexpr int SHELL_THREADS = 256;
static constexpr int SHELL_CAP = 4096;      // particles staged per tile (48 KB)
// the 125 cell offsets ordered by the distance of the cell from the centre of the middle one:
// (di+2) | (dj+2) << 3 | (dk+2) << 6 | n1 << 9 | n2 << 11, n1 / n2 = number of axes with |d| = 1 / 2
__constant__ unsigned short c_shell_order[125] = {
    146, 657, 659, 650, 666, 594, 722, 1161, 1163, 1177, 1179, 1105, 1107, 1233, 1235, 1098, 1114, 1226, 1242, 1609, 1611, 1625,
    1627, 1737, 1739, 1753, 1755, 2192, 2196, 2178, 2210, 2066, 2322, 2696, 2700, 2712, 2716, 2689, 2691, 2721, 2723, 2640, 2644,
    2768, 2772, 2626, 2658, 2754, 2786, 2577, 2579, 2833, 2835, 2570, 2586, 2826, 2842, 3144, 3148, 3160, 3164, 3272, 3276, 3288,
    3292, 3137, 3139, 3169, 3171, 3265, 3267, 3297, 3299, 3081, 3083, 3097, 3099, 3337, 3339, 3353, 3355, 4224, 4228, 4256, 4260,
    4112, 4116, 4368, 4372, 4098, 4130, 4354, 4386, 4672, 4676, 4704, 4708, 4800, 4804, 4832, 4836, 4616, 4620, 4632, 4636, 4872,
    4876, 4888, 4892, 4609, 4611, 4641, 4643, 4865, 4867, 4897, 4899, 6144, 6148, 6176, 6180, 6400, 6404, 6432, 6436};

__device__ __noinline__ bool sdf_box_literal(const GatherParams &g, float x, float y, float z, int ci, int cj, int ckg) {
    const float sr = g.srS;
    const NodeFrame f = node_frame(g, ci, cj, ckg);
    int bminx = pos2idx_d((double)fsub(x, sr), g.invBlockdxSDF), bmaxx = pos2idx_d((double)fadd(x, sr), g.invBlockdxSDF);
    int bminy = pos2idx_d((double)fsub(y, sr), g.invBlockdxSDF), bmaxy = pos2idx_d((double)fadd(y, sr), g.invBlockdxSDF);
    int bminz = pos2idx_d((double)fsub(z, sr), g.invBlockdxSDF), bmaxz = pos2idx_d((double)fadd(z, sr), g.invBlockdxSDF);
    if (f.bi < bminx || f.bi > bmaxx || f.bj < bminy || f.bj > bmaxy || f.bk < bminz || f.bk > bmaxz) return false;
    const float xl = fsub(x, f.ox), yl = fsub(y, f.oy), zl = fsub(z, f.oz);
    int gminx = pos2idx(fsub(xl, sr), g.invdx), gmaxx = pos2idx(fadd(xl, sr), g.invdx);
    int gminy = pos2idx(fsub(yl, sr), g.invdx), gmaxy = pos2idx(fadd(yl, sr), g.invdx);
    int gminz = pos2idx(fsub(zl, sr), g.invdx), gmaxz = pos2idx(fadd(zl, sr), g.invdx);
    return !(f.li < gminx || f.li > gmaxx || f.lj < gminy || f.lj > gmaxy || f.lk < gminz || f.lk > gmaxz);
}

__global__ void __launch_bounds__(SHELL_THREADS) k_sdf_shell(ParticleSoA p, const int *__restrict__ cellStart, GatherParams g,
                                                            const unsigned int *__restrict__ occ, const unsigned int *__restrict__ near5,
                                                            const unsigned int *__restrict__ surf, int WR, unsigned int *__restrict__ minC,
                                                            int tilesX, int tilesY, SdfQueues Q, const int *__restrict__ tileList,
                                                            int *__restrict__ tileCounts) {
    extern __shared__ float stage[];              // [3][SHELL_CAP] particle positions
    __shared__ int off[SHELL_ROWS * (SHELL_F + 1)];   // staged range of every cell of the region, row by row
    __shared__ int rowG[SHELL_ROWS], rowS[SHELL_ROWS + 1];   // first particle of the row's span (global / staged)
    __shared__ __align__(16) int warpSum[SHELL_THREADS / 32];   // (aligned: the compiler reads it with 128-bit loads, which would otherwise take in the last word of rowS as well -- harmless, but compute-sanitizer racecheck reports it)
    __shared__ short needList[P2G_T * P2G_T * P2G_T];        // the unsettled cells of the tile, compacted
    __shared__ int nextTile, needCount;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int I = g.I, J = g.J, K = g.K;
    const float dxf = (float)g.dx, hw = g.hw, sr = g.srS, margin = 1.0e-3f * dxf, dx2 = dxf * dxf;
    const float loIn = -sr + margin, hiIn = dxf + sr - margin, loOut = -sr - margin, hiOut = dxf + sr + margin;
    float *sx = stage, *sy = stage + SHELL_CAP, *sz = stage + 2 * SHELL_CAP;
    const int nTiles = tileCounts[1];
    for (;;) {
        __syncthreads();
        if (tid == 0) { nextTile = atomicAdd(&tileCounts[3], 1); needCount = 0; }
        __syncthreads();
        if (nextTile >= nTiles) break;
        const int tile = tileList[nextTile];
        const int tx = tile % tilesX, ty = (tile / tilesX) % tilesY, tz = tile / (tilesX * tilesY);
        const int i0 = tx * P2G_T, j0 = ty * P2G_T, k0 = tz * P2G_T;
        // unsettled: an empty cell within reach of a particle that the 2x2x2 scatter left without a minimum
        for (int t = tid; t < P2G_T * P2G_T * P2G_T; t += SHELL_THREADS) {
            const int i = i0 + (t % P2G_T), j = j0 + (t / P2G_T) % P2G_T, k = k0 + t / (P2G_T * P2G_T);
            bool need = false;
            if (i < I && j < J && k < K) {
                const size_t word = (size_t)(i >> 5) + (size_t)WR * (j + (size_t)J * k);
                if (((__ldg(near5 + word) & ~__ldg(occ + word)) >> (i & 31)) & 1u) need = minC[i + I * (j + J * k)] == 0u;
            }
            if (need) needList[atomicAdd(&needCount, 1)] = (short)t;
        }
        __syncthreads();
        const int nNeed = needCount;
        if (nNeed == 0) continue;
        // ---- stage, row by row of the (8+4)^2 rows of 12 cells: the span from the first to the last surface cell
        int cnt = 0;
        if (tid < SHELL_ROWS) {
            const int cj = j0 - 2 + tid % SHELL_F, ck = k0 - 2 + tid / SHELL_F;
            int first = -1, last = -1, b = 0;
            if (cj >= 0 && ck >= 0 && cj < J && ck < K) {
                const unsigned int *sw = surf + (size_t)WR * (cj + (size_t)J * ck);
                // the twelve surface bits of the row (cells i0-2 .. i0+9: at most two words)
                unsigned int bits12 = 0u;
                {
                    const int lo = i0 - 2;
                    const int w0 = lo >> 5;                      // (arithmetic shift: -1 for lo < 0)
                    const unsigned long long two = ((w0 >= 0 && w0 < WR) ? (unsigned long long)__ldg(sw + w0) : 0ull) |
                                                   ((w0 + 1 >= 0 && w0 + 1 < WR) ? (unsigned long long)__ldg(sw + w0 + 1) << 32 : 0ull);
                    bits12 = (unsigned int)(two >> (lo - w0 * 32)) & 0xfffu;
                }
                if (bits12) {
                    first = __ffs(bits12) - 1;
                    last = 31 - __clz(bits12);
                    const int c = (i0 - 2) + I * (cj + J * ck);
                    // where each cell of the row begins inside the span (cells outside it: empty ranges at its ends)
                    int cs[SHELL_F + 1];
#pragma unroll
                    for (int m = 0; m <= SHELL_F; m++) cs[m] = (m >= first && m <= last + 1) ? __ldg(cellStart + c + m) : 0;
                    b = __ldg(cellStart + c + first);
                    cnt = __ldg(cellStart + c + last + 1) - b;
#pragma unroll
                    for (int m = 0; m <= SHELL_F; m++) off[tid * (SHELL_F + 1) + m] = (m <= first) ? 0 : (m > last ? cnt : cs[m] - b);
                }
            }
            if (first < 0) {
#pragma unroll
                for (int m = 0; m <= SHELL_F; m++) off[tid * (SHELL_F + 1) + m] = 0;
            }
            rowG[tid] = b;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) warpSum[wid] = incl;
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SHELL_THREADS / 32; w++) { if (w < wid) base += warpSum[w]; total += warpSum[w]; }
        if (total > SHELL_CAP) {       // (uniform) too crowded for the staging area
            for (int t = tid; t < nNeed; t += SHELL_THREADS) {
                const int q = needList[t];
                const int slot = warp_append_slot(Q.farCount);
                Q.farCells[slot] = (i0 + (q % P2G_T)) + I * ((j0 + (q / P2G_T) % P2G_T) + J * (k0 + q / (P2G_T * P2G_T)));
                Q.farBest[slot] = 3.0e38f;
            }
            continue;
        }
        if (tid < SHELL_ROWS) rowS[tid] = base + incl - cnt;
        if (tid == 0) rowS[SHELL_ROWS] = total;
        __syncthreads();
        for (int r = wid; r < SHELL_ROWS; r += SHELL_THREADS / 32) {
            const int gb = rowG[r], sb = rowS[r], n = rowS[r + 1] - sb;
            for (int q = lane; q < n; q += 32) {
                sx[sb + q] = __ldg(p.px + gb + q); sy[sb + q] = __ldg(p.py + gb + q); sz[sb + q] = __ldg(p.pz + gb + q);
            }
        }
        __syncthreads();
        // ---- the search, one thread per unsettled cell
        for (int t = tid; t < nNeed; t += SHELL_THREADS) {
            const int q0 = needList[t];
            const int li = q0 % P2G_T, lj = (q0 / P2G_T) % P2G_T, lk = q0 / (P2G_T * P2G_T);
            const int i = i0 + li, j = j0 + lj, k = k0 + lk, kg = k + g.kOff;
            const float Xn = (float)i * dxf, Yn = (float)j * dxf, Zn = (float)kg * dxf;       // exact (power-of-two dx)
            const float Xh = Xn + hw, Yh = Yn + hw, Zh = Zn + hw;
            const int row0 = (lj + 2) + SHELL_F * (lk + 2), x0 = li + 2;
            float best2 = 3.0e38f;
            for (int o = 0; o < 125; o++) {
                const unsigned int e = c_shell_order[o];
                // squared distance to the nearest point of that cell, less a margin against the rounding of the computed ones
                const float cellMin2 = ((float)((e >> 9) & 3u) * 0.2499f + (float)(e >> 11) * 2.2499f) * dx2;
                if (!(cellMin2 < best2)) break;          // the cells that follow are no nearer
                const int di = (int)(e & 7u) - 2, dj = (int)((e >> 3) & 7u) - 2, dk = (int)((e >> 6) & 7u) - 2;
                const int row = row0 + dj + SHELL_F * dk;
                const int sb = rowS[row];
                const int qb = sb + off[row * (SHELL_F + 1) + x0 + di], qe = sb + off[row * (SHELL_F + 1) + x0 + di + 1];
                for (int q = qb; q < qe; q++) {
                    const float x = sx[q], y = sy[q], z = sz[q];
                    // position relative to the cell's lower corner: inside the search box iff -sr <= u < dx + sr
                    const float ux = fsub(x, Xn), uy = fsub(y, Yn), uz = fsub(z, Zn);
                    if (ux < loOut || ux >= hiOut || uy < loOut || uy >= hiOut || uz < loOut || uz >= hiOut) continue;
                    const bool in = ux >= loIn && ux < hiIn && uy >= loIn && uy < hiIn && uz >= loIn && uz < hiIn;
                    if (!in && !sdf_box_literal(g, x, y, z, i, j, kg)) continue;
                    const float ex = fsub(Xh, x), ey = fsub(Yh, y), ez = fsub(Zh, z);
                    best2 = fminf(best2, fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez)));
                }
            }
            if (best2 < 1.0e38f) minC[i + I * (j + J * k)] = ~__float_as_uint(best2);
        }
    }
}

// One thread per point (i,j,k) of the extended index space: turns the accumulated sums of the three faces of a node
// into velocities and valid flags and the accumulated minimum of the cell into the liquid SDF (clearing the
// accumulators for the next step).  A cell that received no minimum lies in no particle's search box.  Five nodes in
// six are far from any particle and take the short path: zeros and the "no particles" SDF value (finish_phi of an
// infinite distance: 3 dx, particlelevelset.cpp:295), in 32-bit index arithmetic (the scatter path is limited to
// 2^30 nodes).  (ncu on the first version: 210 instructions per node, issue-bound at 225 us; the loads were not the
// limit -- two nodes per thread with all loads hoisted took 240 us.)
__global__ void k_p2g_finish(GatherParams g, P2GGlobal G, double invSScale,
                             float *__restrict__ U, float *__restrict__ V, float *__restrict__ W,
                             unsigned char *__restrict__ validU, unsigned char *__restrict__ validV,
                             unsigned char *__restrict__ validW, float *__restrict__ phiL,
                             const float *__restrict__ phiS, const unsigned int *__restrict__ near3,
                             const unsigned int *__restrict__ near5, int WR, SdfQueues Q) {
    const int I = g.I, J = g.J, K = g.K;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, k = blockIdx.z;      // k: local plane
    if (i > I) return;
    const bool hasU = (j < J && k < K), hasV = (i < I && k < K), hasW = (i < I && j < J), hasC = (i < I && j < J && k < K);
    const int iu = i + (I + 1) * (j + J * k), iv = i + I * (j + (J + 1) * k), iw = i + I * (j + J * k);
    // the neighbourhood bits of the cell the node belongs to (nodes past the last cell of an axis: the last cell)
    const int ci = min(i, I - 1);
    const int word = (ci >> 5) + WR * (min(j, J - 1) + J * min(k, K - 1));
    const unsigned int bit = 1u << (ci & 31);
    const bool n3 = __ldg(near3 + word) & bit;
    if (!n3) {
        // no particle can reach the faces of this node, nor lies in the 3x3x3 cells around the cell
        if (hasU) { U[iu] = 0.0f; validU[iu] = 0; }
        if (hasV) { V[iv] = 0.0f; validV[iv] = 0; }
        if (hasW) { W[iw] = 0.0f; validW[iw] = 0; }
        if (hasC) {
            float phi = g.maxDist;
            if (__ldg(near5 + word) & bit) {           // within reach of a particle two cells away (k_sdf_shell)
                const unsigned int mn = G.minC[iw];
                if (mn) { G.minC[iw] = 0u; phi = finish_phi(g, phiS, __uint_as_float(~mn), i, j, k); }
            }
            phiL[iw] = phi;
        }
        return;
    }
    // (all loads first)
    ulonglong2 aU = make_ulonglong2(0ull, 0ull), aV = aU, aW = aU;
    if (hasU) aU = G.accU[iu];
    if (hasV) aV = G.accV[iv];
    if (hasW) aW = G.accW[iw];
    const unsigned int mn = hasC ? G.minC[iw] : 0u;
    auto finish = [&](ulonglong2 a, ulonglong2 *acc, int idx, float *field, unsigned char *valid, int comp) {
        float val = 0.0f;
        unsigned char ok = 0;
        if ((a.x | a.y) != 0ull) {
            acc[idx] = make_ulonglong2(0ull, 0ull);
            const float wsum = (float)((double)a.x * (1.0 / (16.0 * (double)P2G_WSCALE)));
            const float ssum = (float)((double)(long long)a.y * invSScale);
            int slot = -1;
            if (wsum < P2G_LITERAL_BAND) slot = warp_append_slot(Q.litCount);
            if (slot >= 0 && slot < Q.litCap) Q.litFaces[slot] = (int)(((unsigned int)comp << 30) | (unsigned int)(i + (I + 1) * (j + (J + 1) * k)));
            else { ok = wsum > g.eps; val = ok ? __fdiv_rn(ssum, wsum) : ssum; }   // (a full queue costs digits, nothing else)
        }
        field[idx] = val;
        valid[idx] = ok;
    };
    if (hasU) finish(aU, G.accU, iu, U, validU, 0);
    if (hasV) finish(aV, G.accV, iv, V, validV, 1);
    if (hasW) finish(aW, G.accW, iw, W, validW, 2);
    if (!hasC) return;
    // ---- liquid SDF of cell (i,j,k)
    if (mn) {
        G.minC[iw] = 0u;
        phiL[iw] = finish_phi(g, phiS, __uint_as_float(~mn), i, j, k);
    } else {
        phiL[iw] = g.maxDist;
    }
}

// The queued faces (total weight below P2G_LITERAL_BAND, or next to a cell too dense for the scatter), ONE WARP PER
// FACE: the literal weights of the reference (velocityadvector.cpp:436-441), the lanes on consecutive particles of a
// cell row (coalesced; one thread per face walks ~150 particles through chains of dependent loads), the lane terms
// folded by a fixed shuffle tree and the rows added in (k, j) order: a fixed summation order, hence deterministic.
template <int COMP>
__device__ __forceinline__ void literal_face(const ParticleSoA &p, const int *__restrict__ cellStart, const GatherParams &g, int i,
                                             int j, int k, float &sum, float &wsum) {
    const int I = g.I, J = g.J, K = g.K;
    const int lane = threadIdx.x & 31;
    const float Xn = (float)dmul((double)(float)i, g.dx), Yn = (float)dmul((double)(float)j, g.dx),
                Zn = (float)dmul((double)(float)(k + g.kOff), g.dx);
    // face position: the node, shifted by dx/2 on the two staggered axes
    const float fx = COMP == 0 ? Xn : fadd(Xn, g.hw), fy = COMP == 1 ? Yn : fadd(Yn, g.hw), fz = COMP == 2 ? Zn : fadd(Zn, g.hw);
    const float *vel = COMP == 0 ? p.vx : COMP == 1 ? p.vy : p.vz;
    const int ilo = max(i - 1, 0), ihi = min(COMP == 0 ? i : i + 1, I - 1);
    const int jlo = max(j - 1, 0), jhi = min(COMP == 1 ? j : j + 1, J - 1);
    const int klo = max(k - 1, 0), khi = min(COMP == 2 ? k : k + 1, K - 1);
    sum = 0.0f; wsum = 0.0f;
    // the (at most nine) row ranges, fetched together by the first lanes
    int qb = 0, qe = 0;
    const int nj = jhi - jlo + 1, nk = khi - klo + 1;
    if (lane < nj * nk) {
        const int rowBase = I * ((jlo + lane % nj) + J * (klo + lane / nj));
        qb = __ldg(cellStart + rowBase + ilo);
        qe = __ldg(cellStart + rowBase + ihi + 1);
    }
    for (int r = 0; r < nj * nk; r++) {
        const int rb = __shfl_sync(0xffffffffu, qb, r), re = __shfl_sync(0xffffffffu, qe, r);
        for (int q0 = rb; q0 < re; q0 += 32) {
            const int q = q0 + lane;
            float ws = 0.0f, w = 0.0f;
            if (q < re) {
                const float ax = fsub(fx, __ldg(p.px + q)), ay = fsub(fy, __ldg(p.py + q)), az = fsub(fz, __ldg(p.pz + q));
                const float d2 = fadd(fadd(fmul(ax, ax), fmul(ay, ay)), fmul(az, az));
                if (d2 < g.rsq) {
                    w = kernel_weight(d2, g);
                    ws = fmul(w, __ldg(vel + q));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ws = fadd(ws, __shfl_xor_sync(0xffffffffu, ws, o));
                w = fadd(w, __shfl_xor_sync(0xffffffffu, w, o));
            }
            sum = fadd(sum, ws);
            wsum = fadd(wsum, w);
        }
    }
}

__global__ void __launch_bounds__(256) k_p2g_literal(ParticleSoA p, const int *__restrict__ cellStart, GatherParams g,
                                                     float *__restrict__ U, float *__restrict__ V, float *__restrict__ W,
                                                     unsigned char *__restrict__ validU, unsigned char *__restrict__ validV,
                                                     unsigned char *__restrict__ validW, const int *__restrict__ litFaces,
                                                     const int *__restrict__ litCount, int litCap) {
    const int I = g.I, J = g.J, K = g.K;
    const int n = min(*litCount, litCap);
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n; t += nw) {
        const int e = litFaces[t];
        const int comp = (int)((unsigned int)e >> 30), node = e & 0x3fffffff;
        const int i = node % (I + 1), j = (node / (I + 1)) % (J + 1), k = node / ((I + 1) * (J + 1));
        // (entries queued around a dense cell may name faces that do not exist on the upper borders)
        if ((comp == 0 && (j >= J || k >= K)) || (comp == 1 && (i >= I || k >= K)) || (comp == 2 && (i >= I || j >= J))) continue;
        float sum, wsum;
        long long idx;
        float *field;
        unsigned char *valid;
        if (comp == 0) { literal_face<0>(p, cellStart, g, i, j, k, sum, wsum); idx = (long long)i + (long long)(I + 1) * (j + (long long)J * k); field = U; valid = validU; }
        else if (comp == 1) { literal_face<1>(p, cellStart, g, i, j, k, sum, wsum); idx = (long long)i + (long long)I * (j + (long long)(J + 1) * k); field = V; valid = validV; }
        else { literal_face<2>(p, cellStart, g, i, j, k, sum, wsum); idx = (long long)i + (long long)I * (j + (long long)J * k); field = W; valid = validW; }
        if ((threadIdx.x & 31) == 0) {
            const bool ok = wsum > g.eps;
            field[idx] = ok ? __fdiv_rn(sum, wsum) : sum;     // scalar /= weight only if weight > eps (:448-453)
            valid[idx] = ok ? 1 : 0;
        }
    }
}

static GatherParams make_gather_params(const flip_ctx *c) {
    const Dims &d = c->d;
    GatherParams g;
    g.I = d.I; g.J = d.J; g.K = d.K; g.kOff = d.kOff;
    g.dx = d.dx; g.invdx = 1.0 / d.dx;
    g.hw = (float)(0.5 * d.dx);
    g.blockdxP2G = 10 * d.dx;
    g.blockdxSDF = 10 * d.dx;
    float blockdxf = (float)(10 * d.dx);
    g.invBlockdxSDF = 1.0 / (double)blockdxf;
    float r = (float)c->liquidRadius;
    g.r = r; g.rsq = r * r;
    g.eps = 1e-6f;
    g.coef1 = (4.0f / 9.0f) * (1.0f / (r * r * r * r * r * r));
    g.coef2 = (17.0f / 9.0f) * (1.0f / (r * r * r * r));
    g.coef3 = (22.0f / 9.0f) * (1.0f / (r * r));
    g.rS = r;
    g.srS = 2.0f * r;
    g.maxDist = (float)(3.0 * d.dx);
    g.hwS = (float)(0.5 * d.dx);
    g.halfDxSolid = 0;
    g.packed = c->samplingMode == FLIP_SAMPLING_FAST ? 1 : 0;
    return g;
}

// The SDF and the P2G are produced together; the stage that runs first computes both.
static void run_sdf_p2g(flip_ctx *c) {
    const Dims &d = c->d;
    GatherParams g = make_gather_params(c);
    dim3 block(128, 1, 1);
    dim3 grid(cdiv(d.I + 1, 128), d.J + 1, d.K + 1);
    size_t kt = kt_begin(c);
    // power-of-two dx (every BASELINE config: 0.125) and moderate extents: the exact fast path
    int ex = 0;
    const bool dyadic = (frexp(d.dx, &ex) == 0.5) && std::max(d.I, std::max(d.J, d.Kg)) <= 4096;
    // far-cell queue in the extrapolation scratch (idle at this point of the step)
    int *farCells = c->frontier[0];
    float *farBest = reinterpret_cast<float *>(c->frontier[1]);
    int *farCount = &c->dS->frontierCount[0];
    FLIP_CUDA_CHECK(cudaMemsetAsync(c->dS->frontierCount, 0, 2 * sizeof(int), c->stream));
    // the scatter: single-precision mode, power-of-two dx
    const bool scatter = dyadic && g.packed && (long long)d.nN < (1ll << 30) && !getenv("FLIP_P2G_GATHER");
    if (scatter) {
        const int WR = cdiv(d.I, 32);
        const size_t words = (size_t)WR * d.J * d.K;
        if (!c->p2gAcc[0]) {
            const size_t n[4] = {2 * (size_t)d.nU, 2 * (size_t)d.nV, 2 * (size_t)d.nW, ((size_t)d.nC + 1) / 2};
            for (int m = 0; m < 4; m++) {
                FLIP_CUDA_CHECK(cudaMalloc(&c->p2gAcc[m], sizeof(unsigned long long) * (n[m] + 16)));
                FLIP_CUDA_CHECK(cudaMemsetAsync(c->p2gAcc[m], 0, sizeof(unsigned long long) * (n[m] + 16), c->stream));
            }
        }
        P2GGlobal G;
        G.accU = (ulonglong2 *)c->p2gAcc[0]; G.accV = (ulonglong2 *)c->p2gAcc[1];
        G.accW = (ulonglong2 *)c->p2gAcc[2]; G.minC = (unsigned int *)c->p2gAcc[3];
        unsigned int *bits = c->occBits, *near3 = bits + words, *near5 = bits + 2 * words;
        const size_t stride = ext_stride(d);
        SdfQueues Q;
        Q.farCells = farCells; Q.farBest = farBest; Q.farCount = farCount;
        Q.litFaces = c->frontier[1] + stride; Q.litCount = &c->dS->frontierCount[1]; Q.litCap = (int)std::min<size_t>(2 * stride, 1u << 30);
        // momentum scale 2^q: the largest power of two with vmax 2^q <= 2^18 (vmax: the maximum particle speed, from the sort)
        // momentum scale 2^q: the largest power of two with vmax 2^q <= 2^21 (vmax: the maximum particle speed, from the sort)
        float vmax2;
        { const unsigned int bitsv = c->hS->maxSpeedSqBits; memcpy(&vmax2, &bitsv, sizeof(float)); }
        const double vmax = std::max(std::sqrt((double)vmax2) * 1.0000002, 1e-30);
        int qg = (int)std::floor(std::log2(2097152.0 / vmax));
        qg = std::max(-60, std::min(60, qg));
        const double invSScale = std::ldexp(1.0, -(qg + 24));
        unsigned int *surf = c->occBits + 3 * words;
        const int tX = cdiv(d.I, P2G_T), tY = cdiv(d.J, P2G_T), tZ = cdiv(d.K, P2G_T);
        // tile bookkeeping: [0,4) counts and tickets, then flags (two per tile; x tiles padded to whole words of the
        // bitmaps), then the two lists
        const int nTiles = tX * tY * tZ, nFlagTiles = nTiles + 8;
        if (!c->p2gTiles) {
            FLIP_CUDA_CHECK(cudaMalloc(&c->p2gTiles, sizeof(int) * (8 + 2 * (size_t)nFlagTiles + 2 * (size_t)nTiles)));
            FLIP_CUDA_CHECK(cudaMemsetAsync(c->p2gTiles, 0, sizeof(int) * (8 + 2 * (size_t)nFlagTiles + 2 * (size_t)nTiles), c->stream));
        }
        int *tileCounts = c->p2gTiles, *tileFlags = c->p2gTiles + 8, *listP = tileFlags + 2 * nFlagTiles, *listS = listP + nTiles;
        FLIP_CUDA_CHECK(cudaMemsetAsync(tileCounts, 0, 4 * sizeof(int), c->stream));
        if (!c->occBitsValid) {
            k_occ_bits<<<cdiv((long long)words * 32, TPB), TPB, 0, c->stream>>>(c->cellStart, d.I, d.J * d.K, WR, bits);
            c->launches++;
        }
        k_occ_dilate<<<cdiv((long long)words, TPB), TPB, 0, c->stream>>>(bits, d.I, d.J, d.K, WR, near3, near5, surf, tileFlags);
        k_tile_lists<<<cdiv(nTiles, TPB), TPB, 0, c->stream>>>(nTiles, tileFlags, listP, listS, tileCounts);
        const size_t smem = (size_t)P2G_SLOTS * P2G_WORDS * sizeof(int);
        {
            static bool attrS = false;
            if (!attrS) {
                FLIP_CUDA_CHECK(cudaFuncSetAttribute(k_p2g_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attrS = true;
            }
        }
        k_p2g_scatter<<<148 * 3, P2G_THREADS, smem, c->stream>>>(c->P[c->cur_buf], c->cellStart, g, G, tX, tY, qg, Q, listP, tileCounts);
        {
            static bool attr = false;
            const size_t shellSmem = 3 * (size_t)SHELL_CAP * sizeof(float);
            if (!attr) {
                FLIP_CUDA_CHECK(cudaFuncSetAttribute(k_sdf_shell, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shellSmem));
                attr = true;
            }
            k_sdf_shell<<<148 * 3, SHELL_THREADS, shellSmem, c->stream>>>(c->P[c->cur_buf], c->cellStart, g, bits, near5, surf, WR,
                                                                         G.minC, tX, tY, Q, listS, tileCounts);
        }
        // (one CTA per row of nodes where the row fits: I + 1 = 257 nodes would leave a third CTA of 128 with one thread)
        const int fT = std::min(1024, ((d.I + 1 + 31) / 32) * 32);
        k_p2g_finish<<<dim3(cdiv(d.I + 1, fT), d.J + 1, d.K + 1), fT, 0, c->stream>>>(g, G, invSScale, c->U, c->V, c->W, c->validU, c->validV, c->validW, c->phiL,
                                                    c->phiS, near3, near5, WR, Q);
        k_p2g_literal<<<148 * 8, 256, 0, c->stream>>>(c->P[c->cur_buf], c->cellStart, g, c->U, c->V, c->W, c->validU, c->validV,
                                                      c->validW, Q.litFaces, Q.litCount, Q.litCap);
        // (cells of a region too crowded for the shell kernel's staging area; it is after k_p2g_finish that their value stands)
        k_sdf_far<<<148 * 8, 256, 0, c->stream>>>(c->P[c->cur_buf], c->cellStart, g, c->phiL, c->phiS, farCells, farBest, farCount);
        c->launches += 7;
        kt_end(c, FLIP_KERNEL_SDF_P2G, kt);
        FLIP_CUDA_CHECK(cudaGetLastError());
        return;
    } else if (dyadic)
        k_sdf_p2g<true><<<grid, block, 0, c->stream>>>(c->P[c->cur_buf], c->cellStart, g, c->U, c->V, c->W, c->validU, c->validV,
                                                       c->validW, c->phiL, c->phiS, c->occ, farCells, farBest, farCount);
    else
        k_sdf_p2g<false><<<grid, block, 0, c->stream>>>(c->P[c->cur_buf], c->cellStart, g, c->U, c->V, c->W, c->validU, c->validV,
                                                        c->validW, c->phiL, c->phiS, c->occ, farCells, farBest, farCount);
    k_sdf_far<<<148 * 8, 256, 0, c->stream>>>(c->P[c->cur_buf], c->cellStart, g, c->phiL, c->phiS, farCells, farBest, farCount);
    c->launches++;
    kt_end(c, FLIP_KERNEL_SDF_P2G, kt);
    c->launches++;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

void stage_liquid_sdf(flip_ctx *c) {
    if (slab_on(c)) slab_exchange_ghosts(c);    // neighbours' particles within the halo planes, then re-sort
    run_sdf_p2g(c);
}

// _advectVelocityField (fluidsimulation.cpp:3253-3278): valid.reset(); MAC.clear(); advect().
// The fused gather of stage_liquid_sdf already produced U,V,W and the valid masks from the same
// particle state, so this stage has nothing left to launch.  (When driven stage by stage the
// arrays can be read back here and compared with VelocityAdvector::advect.)
void stage_p2g(flip_ctx *c) { (void)c; }

// ------------------------------------------------------------------------------------------------
// G2P (PIC/FLIP) and RK3 advection + collision
// ------------------------------------------------------------------------------------------------
struct AdvectParams {
    int n;
    GridGeom G;
    double dx, invdx, hdx;
    float ratioPIC, ratioFLIP;     // (float)_ratioPICFLIP, (float)(1 - _ratioPICFLIP)   :4088
    float c1, c2, c3;              // (float)(0.5*dt), (float)(0.75*dt), (float)(dt/9.0f) :4191-4196
    // boundary box after expand(-3dx-1e-4) and expand(-solidBufferWidth*dx)  (:2834-2839, :4214-4215)
    float bminx, bminy, bminz;     // AABB::position (vec3, float)
    double bw, bh, bd;             // AABB::width/height/depth (double)
    double nearCell, invNearCell;  // _nearSolidGridCellSize
    int nsI, nsJ, nsK;
    float stepDistance;            // _markerParticleStepDistanceFactor * (float)_dx
    float maxResolvedDistance;     // _CFLConditionNumber * _dx
    double pushOut;                // _solidBufferWidth * _dx (float*double -> double)
    FastGeom F;                    // single-precision sampling fast path (device_math.cuh)
};

// Velocity sample for the particle kernels.  FAST (flip_set_sampling_mode, power-of-two dx): single-precision
// blend on the interior fast path of device_math.cuh, the literal double-precision restatement elsewhere.
// FLIP update of one particle from the two sampled velocities: vFLIP = v + vPIC - vOld ;
// v = ratio*vPIC + (1-ratio)*vFLIP   (fluidsimulation.cpp:4084-4090)
__device__ __forceinline__ void picflip_update(ParticleSoA &p, const AdvectParams &a, int t, float nx, float ny, float nz,
                                               float ox, float oy, float oz) {
    float fx = fsub(fadd(p.vx[t], nx), ox);
    float fy = fsub(fadd(p.vy[t], ny), oy);
    float fz = fsub(fadd(p.vz[t], nz), oz);
    p.vx[t] = fadd(fmul(a.ratioPIC, nx), fmul(a.ratioFLIP, fx));
    p.vy[t] = fadd(fmul(a.ratioPIC, ny), fmul(a.ratioFLIP, fy));
    p.vz[t] = fadd(fmul(a.ratioPIC, nz), fmul(a.ratioFLIP, fz));
}

// Literal double-precision sampling (FLIP_SAMPLING_EXACT, non-power-of-two dx, and the particles the
// single-precision kernel deferred).  list == nullptr: every particle; else the listed ones.
__global__ void k_g2p(ParticleSoA p, AdvectParams a, MacField fnew, MacField fold, const int *__restrict__ list,
                      const int *__restrict__ listCount) {
    const int n = list ? *listCount : a.n;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int t = list ? list[q] : q;
        float x = p.px[t], y = p.py[t], z = p.pz[t];
        float nx, ny, nz, ox, oy, oz;
        sample_velocity(fnew, a.G, a.dx, a.invdx, a.hdx, x, y, z, nx, ny, nz);
        sample_velocity(fold, a.G, a.dx, a.invdx, a.hdx, x, y, z, ox, oy, oz);
        picflip_update(p, a, t, nx, ny, nz, ox, oy, oz);
    }
}

// Single-precision blend (FLIP_SAMPLING_FAST): exact indices and weights, see device_math.cuh.  Particles
// outside the interior box (none in a walled domain) are queued for the literal kernel.
__global__ void k_g2p_fast(ParticleSoA p, AdvectParams a, MacField fnew, MacField fold, int *__restrict__ deferred,
                           int *__restrict__ deferredCount) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    float x = p.px[t], y = p.py[t], z = p.pz[t];
    if (!fast_interior(a.F, x, y, z)) {
        deferred[warp_append_slot(deferredCount)] = t;
        return;
    }
    float nx, ny, nz, ox, oy, oz;
    const FastStencil s = fast_stencil(a.F, x, y, z);   // both fields share indices and weights
    fast_sample(fnew, a.F, s, nx, ny, nz);
    fast_sample(fold, a.F, s, ox, oy, oz);
    picflip_update(p, a, t, nx, ny, nz, ox, oy, oz);
}

// AABB::isPointInside  aabb.cpp:130-133
__device__ __forceinline__ bool box_inside(const AdvectParams &a, float x, float y, float z) {
    return x >= a.bminx && y >= a.bminy && z >= a.bminz && (double)x < dadd((double)a.bminx, a.bw) &&
           (double)y < dadd((double)a.bminy, a.bh) && (double)z < dadd((double)a.bminz, a.bd);
}
// AABB::getNearestPointInsideAABB(p, 1e-6)  aabb.cpp:497-518
__device__ __forceinline__ void box_nearest(const AdvectParams &a, float &x, float &y, float &z) {
    if (box_inside(a, x, y, z)) return;
    float maxx = fadd(a.bminx, (float)a.bw), maxy = fadd(a.bminy, (float)a.bh), maxz = fadd(a.bminz, (float)a.bd);
    const double eps = 1e-6;
    x = fmaxf(x, a.bminx); y = fmaxf(y, a.bminy); z = fmaxf(z, a.bminz);
    x = (float)fmin((double)x, dsub((double)maxx, eps));
    y = (float)fmin((double)y, dsub((double)maxy, eps));
    z = (float)fmin((double)z, dsub((double)maxz, eps));
}

__device__ __forceinline__ bool near_solid(const unsigned char *__restrict__ ns, const AdvectParams &a, float x, float y, float z) {
    int i = pos2idx(x, a.invNearCell), j = pos2idx(y, a.invNearCell), k = pos2idx(z, a.invNearCell);
    if (i < 0 || j < 0 || k < 0 || i >= a.nsI || j >= a.nsJ || k >= a.nsK) return false;
    return __ldg(ns + i + a.nsI * (j + a.nsJ * k)) != 0;
}

// FluidSimulation::_resolveCollision  fluidsimulation.cpp:4221-4296
__device__ void resolve_collision(const AdvectParams &a, const float *__restrict__ phiS,
                                  const unsigned char *__restrict__ ns, float ox, float oy, float oz, float &nx,
                                  float &ny, float &nz) {
    const int gi = a.G.I + 1, gj = a.G.J + 1, gk = a.G.K + 1, ko = a.G.kOff;
    int ci = pos2idx(nx, a.invdx), cj = pos2idx(ny, a.invdx), ck = pos2idx(nz, a.invdx);
    if (!(ci >= 0 && cj >= 0 && ck >= 0 && ci < a.G.I && cj < a.G.J && ck < a.G.Kg)) box_nearest(a, nx, ny, nz);
    if (!near_solid(ns, a, ox, oy, oz) && !near_solid(ns, a, nx, ny, nz)) return;

    const float eps = 1e-6f;
    float dxv = fsub(nx, ox), dyv = fsub(ny, oy), dzv = fsub(nz, oz);
    float travel = length3(dxv, dyv, dzv);
    if (travel < eps) return;
    int numSteps = (int)ceilf(__fdiv_rn(travel, a.stepDistance));
    // (newp - oldp).normalize(): v * (float)(1.0/len)   vmath.cpp:100-103
    float inv = (float)(1.0 / (double)travel);
    float sx = fmul(dxv, inv), sy = fmul(dyv, inv), sz = fmul(dzv, inv);

    float lx = ox, ly = oy, lz = oz;   // lastPosition
    float cx = 0.f, cy = 0.f, cz = 0.f;
    bool found = false;
    float collisionPhi = 0.0f;
    ScalarSample smp;
    for (int s = 0; s < numSteps; s++) {
        if (s == numSteps - 1) { cx = nx; cy = ny; cz = nz; }
        else {
            float f = fmul((float)(s + 1), a.stepDistance);
            cx = fadd(ox, fmul(sx, f)); cy = fadd(oy, fmul(sy, f)); cz = fadd(oz, fmul(sz, f));
        }
        fetch_scalar(phiS, gi, gj, gk, a.dx, a.invdx, cx, cy, cz, smp, ko);
        float phi = scalar_value(smp);
        if (phi < 0.0f || !box_inside(a, cx, cy, cz)) { collisionPhi = phi; found = true; break; }
        lx = cx; ly = cy; lz = cz;
    }
    if (!found) return;

    float rx, ry, rz;
    float gx, gy, gz;
    scalar_gradient(smp, gx, gy, gz);
    if (length3(gx, gy, gz) > eps) {
        float glen = length3(gx, gy, gz);
        float ginv = (float)(1.0 / (double)glen);
        gx = fmul(gx, ginv); gy = fmul(gy, ginv); gz = fmul(gz, ginv);
        // currentPosition - (collisionPhi - _solidBufferWidth*_dx) * grad : the scalar is double, narrowed
        // to float by operator*(float, vec3)
        float sc = (float)dsub((double)collisionPhi, a.pushOut);
        rx = fsub(cx, fmul(gx, sc)); ry = fsub(cy, fmul(gy, sc)); rz = fsub(cz, fmul(gz, sc));
        float rphi = sample_scalar(phiS, gi, gj, gk, a.dx, a.invdx, rx, ry, rz, ko);
        float rdist = length3(fsub(rx, cx), fsub(ry, cy), fsub(rz, cz));
        if (rphi < 0 || rdist > a.maxResolvedDistance) { rx = lx; ry = ly; rz = lz; }
    } else {
        rx = lx; ry = ly; rz = lz;
    }
    if (!box_inside(a, rx, ry, rz)) {
        float qx = rx, qy = ry, qz = rz;
        box_nearest(a, rx, ry, rz);
        float rphi = sample_scalar(phiS, gi, gj, gk, a.dx, a.invdx, rx, ry, rz, ko);
        float rdist = length3(fsub(rx, qx), fsub(ry, qy), fsub(rz, qz));
        if (rphi < 0.0f || rdist > a.maxResolvedDistance) { rx = lx; ry = ly; rz = lz; }
    }
    nx = rx; ny = ry; nz = rz;
}

// Ralston RK3 increment from the three samples (_RK3  fluidsimulation.cpp:4191-4198)
__device__ __forceinline__ void rk3_combine(const AdvectParams &a, float x, float y, float z, float k1x, float k1y, float k1z,
                                            float k2x, float k2y, float k2z, float k3x, float k3y, float k3z, float &nx,
                                            float &ny, float &nz) {
    float sx = fadd(fadd(fmul(k1x, 2.0f), fmul(k2x, 3.0f)), fmul(k3x, 4.0f));
    float sy = fadd(fadd(fmul(k1y, 2.0f), fmul(k2y, 3.0f)), fmul(k3y, 4.0f));
    float sz = fadd(fadd(fmul(k1z, 2.0f), fmul(k2z, 3.0f)), fmul(k3z, 4.0f));
    nx = fadd(x, fmul(sx, a.c3)); ny = fadd(y, fmul(sy, a.c3)); nz = fadd(z, fmul(sz, a.c3));
}

// z-slab: a sample position whose stencil would leave the local planes towards a neighbouring slab (not towards the
// wall of the domain) reads zeros there; report it instead (DeviceScalars::slabError[2], acted on after the
// migration exchange), the halo was sized for CFL-bounded substeps
__device__ __forceinline__ bool leaves_halo(const AdvectParams &a, float z) {
    const int kl = pos2idx(z, a.invdx) - a.G.kOff;
    return (kl < 1 && a.G.kOff > 0) || (kl > a.G.K - 2 && a.G.kOff + a.G.K < a.G.Kg);
}

// literal sampling; list as in k_g2p
__global__ void k_advance(ParticleSoA p, AdvectParams a, MacField f, const float *__restrict__ phiS,
                          const unsigned char *__restrict__ ns, const int *__restrict__ list,
                          const int *__restrict__ listCount, DeviceScalars *S) {
    const int n = list ? *listCount : a.n;
    const bool slab = a.G.K != a.G.Kg;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int t = list ? list[q] : q;
        float x = p.px[t], y = p.py[t], z = p.pz[t];
        float k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z;
        sample_velocity(f, a.G, a.dx, a.invdx, a.hdx, x, y, z, k1x, k1y, k1z);
        const float z2 = fadd(z, fmul(k1z, a.c1));
        sample_velocity(f, a.G, a.dx, a.invdx, a.hdx, fadd(x, fmul(k1x, a.c1)), fadd(y, fmul(k1y, a.c1)), z2, k2x, k2y, k2z);
        const float z3 = fadd(z, fmul(k2z, a.c2));
        sample_velocity(f, a.G, a.dx, a.invdx, a.hdx, fadd(x, fmul(k2x, a.c2)), fadd(y, fmul(k2y, a.c2)), z3, k3x, k3y, k3z);
        if (slab && (leaves_halo(a, z) || leaves_halo(a, z2) || leaves_halo(a, z3))) S->slabError[2] = 1;
        float nx, ny, nz;
        rk3_combine(a, x, y, z, k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z, nx, ny, nz);
        resolve_collision(a, phiS, ns, x, y, z, nx, ny, nz);
        p.px[t] = nx; p.py[t] = ny; p.pz[t] = nz;
    }
}

// single-precision blend; a particle whose start or intermediate RK positions leave the interior box is queued
// for the literal kernel (untouched here)
__global__ void k_advance_fast(ParticleSoA p, AdvectParams a, MacField f, const float *__restrict__ phiS,
                               const unsigned char *__restrict__ ns, int *__restrict__ deferred,
                               int *__restrict__ deferredCount) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const float x = p.px[t], y = p.py[t], z = p.pz[t];
    float k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z;
    bool ok = fast_interior(a.F, x, y, z);
    if (ok) {
        fast_sample(f, a.F, fast_stencil(a.F, x, y, z), k1x, k1y, k1z);
        const float x2 = fadd(x, fmul(k1x, a.c1)), y2 = fadd(y, fmul(k1y, a.c1)), z2 = fadd(z, fmul(k1z, a.c1));
        ok = fast_interior(a.F, x2, y2, z2);
        if (ok) {
            fast_sample(f, a.F, fast_stencil(a.F, x2, y2, z2), k2x, k2y, k2z);
            const float x3 = fadd(x, fmul(k2x, a.c2)), y3 = fadd(y, fmul(k2y, a.c2)), z3 = fadd(z, fmul(k2z, a.c2));
            ok = fast_interior(a.F, x3, y3, z3);
            if (ok) fast_sample(f, a.F, fast_stencil(a.F, x3, y3, z3), k3x, k3y, k3z);
        }
    }
    if (!ok) {
        deferred[warp_append_slot(deferredCount)] = t;
        return;
    }
    float nx, ny, nz;
    rk3_combine(a, x, y, z, k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z, nx, ny, nz);
    resolve_collision(a, phiS, ns, x, y, z, nx, ny, nz);
    p.px[t] = nx; p.py[t] = ny; p.pz[t] = nz;
}

// G2P and the RK3 advection in ONE pass over the particles (whole-step path, single precision, single GPU): the grid
// sample of the PIC/FLIP update at the particle's position IS the first Ralston stage (the same field at the same
// point, fluidsimulation.cpp:4084 and :4192), so the fused kernel reads the position once, samples the new field
// three times instead of four and the saved field once, and writes velocity and position; the speed histogram of
// _getMarkerParticleSpeedLimit (:4300-4305), which the sort that follows needs, is taken from the updated velocity
// while it is in registers.  Arithmetic per particle is that of k_g2p_fast followed by k_advance_fast (same device
// functions): results are bit-identical to the two-kernel path.  Particles that leave the interior box at any
// sample go to the literal kernels (k_g2p then k_advance), untouched here.
__global__ void __launch_bounds__(256) k_g2p_advance_fused(ParticleSoA p, AdvectParams a, MacField fnew, MacField fold,
                                                          const float *__restrict__ phiS, const unsigned char *__restrict__ ns,
                                                          int *__restrict__ deferred, int *__restrict__ deferredCount,
                                                          double speedLimitStep, int nbins, DeviceScalars *S) {
    __shared__ int sh[8];
    if (threadIdx.x < 8) sh[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.n) {
        const float x = p.px[t], y = p.py[t], z = p.pz[t];
        float k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z, ox = 0.f, oy = 0.f, oz = 0.f;
        bool ok = fast_interior(a.F, x, y, z);
        if (ok) {
            const FastStencil s = fast_stencil(a.F, x, y, z);
            fast_sample(fnew, a.F, s, k1x, k1y, k1z);
            fast_sample(fold, a.F, s, ox, oy, oz);
            const float x2 = fadd(x, fmul(k1x, a.c1)), y2 = fadd(y, fmul(k1y, a.c1)), z2 = fadd(z, fmul(k1z, a.c1));
            ok = fast_interior(a.F, x2, y2, z2);
            if (ok) {
                fast_sample(fnew, a.F, fast_stencil(a.F, x2, y2, z2), k2x, k2y, k2z);
                const float x3 = fadd(x, fmul(k2x, a.c2)), y3 = fadd(y, fmul(k2y, a.c2)), z3 = fadd(z, fmul(k2z, a.c2));
                ok = fast_interior(a.F, x3, y3, z3);
                if (ok) fast_sample(fnew, a.F, fast_stencil(a.F, x3, y3, z3), k3x, k3y, k3z);
            }
        }
        float vx, vy, vz;
        if (!ok) {
            deferred[warp_append_slot(deferredCount)] = t;
            // (its histogram entry is added by k_speed_hist_list once the literal kernels have updated it)
        } else {
            picflip_update(p, a, t, k1x, k1y, k1z, ox, oy, oz);
            vx = p.vx[t]; vy = p.vy[t]; vz = p.vz[t];
            const double b = fmin(floor((double)length3(vx, vy, vz) / speedLimitStep), (double)(nbins - 1));
            const int bi = (int)b;
            if (bi > 0) atomicAdd(&sh[bi], 1);
            float nx, ny, nz;
            rk3_combine(a, x, y, z, k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z, nx, ny, nz);
            resolve_collision(a, phiS, ns, x, y, z, nx, ny, nz);
            p.px[t] = nx; p.py[t] = ny; p.pz[t] = nz;
        }
    }
    __syncthreads();
    if (threadIdx.x > 0 && threadIdx.x < 8 && sh[threadIdx.x]) atomicAdd(&S->speedHist[threadIdx.x], sh[threadIdx.x]);
}
// the histogram entries of the deferred particles (after the literal G2P)
__global__ void k_speed_hist_list(ParticleSoA p, const int *__restrict__ list, const int *__restrict__ listCount, double speedLimitStep,
                                  int nbins, DeviceScalars *S) {
    const int n = *listCount;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int t = list[q];
        const double b = fmin(floor((double)length3(p.vx[t], p.vy[t], p.vz[t]) / speedLimitStep), (double)(nbins - 1));
        const int bi = (int)b;
        if (bi > 0) atomicAdd(&S->speedHist[bi], 1);
    }
}

// single-precision sampling needs a power-of-two dx (exact index / weight arithmetic in float) and extents
// whose coordinates stay far below 2^24 ulps
static bool sampling_fast(const flip_ctx *c) {
    int ex = 0;
    return c->samplingMode == FLIP_SAMPLING_FAST && frexp(c->d.dx, &ex) == 0.5 &&
           std::max(c->d.I, std::max(c->d.J, c->d.Kg)) <= 4096;
}

static AdvectParams make_advect_params(const flip_ctx *c, double dt) {
    const Dims &d = c->d;
    AdvectParams a;
    a.n = c->np;
    a.G.I = d.I; a.G.J = d.J; a.G.K = d.K; a.G.Kg = d.Kg; a.G.kOff = d.kOff;
    a.dx = d.dx; a.invdx = 1.0 / d.dx; a.hdx = 0.5 * d.dx;
    a.ratioPIC = (float)c->ratioPICFLIP;
    a.ratioFLIP = (float)(1 - c->ratioPICFLIP);
    a.c1 = (float)(0.5 * dt);
    a.c2 = (float)(0.75 * dt);
    a.c3 = (float)(dt / 9.0f);
    // _getBoundaryAABB + expand(-_solidBufferWidth*_dx): AABB::expand (aabb.cpp:122-128) with vec3 float position
    {
        float px = 0.f, py = 0.f, pz = 0.f;
        double w = d.I * d.dx, h = d.J * d.dx, dp = d.Kg * d.dx;
        auto expand = [&](double v) {
            double hh = 0.5 * v;
            float hf = (float)hh;
            px = px - hf; py = py - hf; pz = pz - hf;
            w += v; h += v; dp += v;
        };
        double eps = 1e-4;
        expand(-3 * d.dx - eps);
        expand(-(double)(float)c->solidBufferWidth * d.dx);
        a.bminx = px; a.bminy = py; a.bminz = pz;
        a.bw = w; a.bh = h; a.bd = dp;
    }
    a.nearCell = c->nearSolidFactor * d.dx;
    a.invNearCell = 1.0 / a.nearCell;
    a.nsI = c->nsI; a.nsJ = c->nsJ; a.nsK = c->nsK;
    a.stepDistance = c->markerParticleStepDistanceFactor * (float)d.dx;
    a.maxResolvedDistance = (float)(c->CFL * d.dx);
    a.pushOut = (double)(float)c->solidBufferWidth * d.dx;
    a.F.invdx = (float)(1.0 / d.dx); a.F.hdx = (float)(0.5 * d.dx);
    a.F.lox = a.F.loy = (float)d.dx; a.F.loz = (float)((d.kOff + 1) * d.dx);
    a.F.hix = (float)((d.I - 1) * d.dx); a.F.hiy = (float)((d.J - 1) * d.dx); a.F.hiz = (float)((d.kOff + d.K - 1) * d.dx);
    a.F.I = d.I; a.F.J = d.J; a.F.kOff = d.kOff;
    return a;
}

void stage_g2p(flip_ctx *c) {
    if (c->np == 0) return;
    AdvectParams a = make_advect_params(c, 0.0);
    MacField fn{c->U, c->V, c->W}, fo{c->sU, c->sV, c->sW};
    size_t kt = kt_begin(c);
    const ParticleSoA P = soa_offset(c->P[c->cur_buf], c->ownedBegin);
    if (sampling_fast(c)) {
        int *deferred = c->sortIdx, *cnt = &c->dS->deferredCount;     // the sort scratch is idle between sorts
        FLIP_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int), c->stream));
        k_g2p_fast<<<cdiv(c->np, TPB), TPB, 0, c->stream>>>(P, a, fn, fo, deferred, cnt);
        k_g2p<<<148, TPB, 0, c->stream>>>(P, a, fn, fo, deferred, cnt);
        c->launches++;
    } else {
        k_g2p<<<cdiv(c->np, TPB), TPB, 0, c->stream>>>(P, a, fn, fo, nullptr, nullptr);
    }
    kt_end(c, FLIP_KERNEL_G2P, kt);
    c->launches++;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

// whole-step path: stage_g2p + the kernels of stage_advance as one pass (see k_g2p_advance_fused); the sort that ends
// stage_advance follows as usual and finds its speed histogram ready
bool stage_g2p_advance_fused(flip_ctx *c, double dt) {
    if (c->np == 0 || slab_on(c) || !sampling_fast(c) || getenv("FLIP_NO_FUSED_ADVANCE")) return false;
    AdvectParams a = make_advect_params(c, dt);
    MacField fn{c->U, c->V, c->W}, fo{c->sU, c->sV, c->sW};
    const ParticleSoA P = soa_offset(c->P[c->cur_buf], c->ownedBegin);
    int *deferred = c->sortIdx, *cnt = &c->dS->deferredCount;
    const double speedLimitStep = c->CFL * c->d.dx / c->frameDt;
    size_t kt = kt_begin(c);
    FLIP_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int), c->stream));
    k_g2p_advance_fused<<<cdiv(c->np, 256), 256, 0, c->stream>>>(P, a, fn, fo, c->phiS, c->nearSolid, deferred, cnt, speedLimitStep,
                                                                 c->maxSubsteps, c->dS);
    // the few particles outside the interior box: the literal kernels, in stage order
    k_g2p<<<148, TPB, 0, c->stream>>>(P, a, fn, fo, deferred, cnt);
    k_speed_hist_list<<<148, TPB, 0, c->stream>>>(P, deferred, cnt, speedLimitStep, c->maxSubsteps, c->dS);
    k_advance<<<148, TPB, 0, c->stream>>>(P, a, fn, c->phiS, c->nearSolid, deferred, cnt, c->dS);
    kt_end(c, FLIP_KERNEL_G2P_ADVANCE, kt);
    c->launches += 4;
    FLIP_CUDA_CHECK(cudaGetLastError());
    c->speedHistReady = true;
    return true;
}

void stage_advance(flip_ctx *c, double dt) {
    if (c->fusedAdvanceDone) {
        // (the kernels ran with the G2P stage; only the removal rules and the sort are left)
        c->fusedAdvanceDone = false;
        particles_sort(c, true, c->frameDt, 0, c->np, false);
        return;
    }
    if (c->np > 0) {
        AdvectParams a = make_advect_params(c, dt);
        MacField fn{c->U, c->V, c->W};
        size_t kt = kt_begin(c);
        const ParticleSoA P = soa_offset(c->P[c->cur_buf], c->ownedBegin);
        if (sampling_fast(c)) {
            int *deferred = c->sortIdx, *cnt = &c->dS->deferredCount;
            FLIP_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int), c->stream));
            k_advance_fast<<<cdiv(c->np, TPB), TPB, 0, c->stream>>>(P, a, fn, c->phiS, c->nearSolid, deferred, cnt);
            k_advance<<<148, TPB, 0, c->stream>>>(P, a, fn, c->phiS, c->nearSolid, deferred, cnt, c->dS);
            c->launches++;
        } else {
            k_advance<<<cdiv(c->np, TPB), TPB, 0, c->stream>>>(P, a, fn, c->phiS, c->nearSolid, nullptr, nullptr, c->dS);
        }
        kt_end(c, FLIP_KERNEL_ADVANCE, kt);
        c->launches++;
        FLIP_CUDA_CHECK(cudaGetLastError());
    }
    // _removeMarkerParticles(_currentFrameDeltaTime) + re-sort for the next step
    if (slab_on(c)) slab_drop_ghosts_and_migrate(c);     // ends with the rules sort over the owned particles
    else particles_sort(c, true, c->frameDt, 0, c->np, false);
}

}  // namespace flip
