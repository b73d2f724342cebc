// Seeding of queued fluid objects on the device: FluidSimulation::_updateAddedFluidMeshObjectQueue
// (fluidsimulation.cpp:4724-4759) with _addNewFluidCells / _addNewFluidCellsThread (:4397-4428, :4479-4518) and the
// ParticleMaskGrid (particlemaskgrid.cpp:55-112).
//
// The reference walks the cells of the object on the host, tests the eight sub-cell points (+-dx/4)^3 of every cell
// against the object's signed distance field (trilinear sample of the nodal MeshLevelSet, d <= 0 keeps the point)
// and the solid SDF (> 0 keeps it), and skips points whose sub-cell already holds a particle -- of the simulation or of
// an object seeded earlier in the same queue.  Here: one kernel marks the occupied sub-cells from the particle store,
// one kernel per object does the tests for all cells of the object's cell range at once and appends the survivors to
// the store (one atomic per warp), and the store is cell-sorted again.  The object's SDF comes either as a nodal array
// (what a MeshLevelSet holds, flip_add_fluid_sdf) or as an axis-aligned box whose nodal distances are evaluated in
// place (flip_add_fluid_box).  What is not reproduced: the jitter of the reference (amplitude 0.25 (jitterFactor -
// 1e-3) dx, i.e. 2.5e-4 dx at the default factor 0, drawn from the unseeded rand(): SURVEY §0 fact 9) -- the seeds sit
// exactly on the sub-cell centres.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "device_math.cuh"
#include "flip_internal.h"

namespace flip {

static constexpr int TPB = 256;

struct SeedParams {
    int I, J, K, Kg, kOff, kOwn0, kOwn1;
    double dx, invdx, invsubdx;       // invsubdx: 1 / (0.5 dx)  particlemaskgrid.cpp:35
    int lo[3], hi[3];                 // cell range of the object (global indices), [lo, hi)
    float boxLo[3], boxHi[3];         // analytic box (sdf == nullptr)
    float vel[3];
    int org[3], nodes[3];             // the source's own level-set grid (inflow velocity constraint)
};

// sub-cell bit of a point: ParticleMaskGrid::addParticle / isSubCellSet (particlemaskgrid.cpp:55-101)
__device__ __forceinline__ void subcell_of(const SeedParams &s, float x, float y, float z, int &i, int &j, int &k, unsigned int &bit) {
    const int si = pos2idx(x, s.invsubdx), sj = pos2idx(y, s.invsubdx), sk = pos2idx(z, s.invsubdx);
    i = si >> 1; j = sj >> 1; k = sk >> 1;
    bit = 1u << ((si & 1) | ((sj & 1) << 1) | ((sk & 1) << 2));
}

// mask[cell] |= sub-cell bits of the particles in the store (maskgrid.addParticle over _markerParticles, :4729-4732)
__global__ void k_seed_mask(ParticleSoA p, int n, SeedParams s, unsigned int *__restrict__ mask) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i, j, k;
    unsigned int bit;
    subcell_of(s, p.px[t], p.py[t], p.pz[t], i, j, k, bit);
    k -= s.kOff;
    if (i < 0 || j < 0 || k < 0 || i >= s.I || j >= s.J || k >= s.K) return;
    atomicOr(&mask[((size_t)i + (size_t)s.I * (j + (size_t)s.J * k)) >> 2], bit << (8 * ((i + s.I * (j + s.J * k)) & 3)));
}

// Euclidean signed distance to an axis-aligned box (negative inside)
__device__ __forceinline__ float box_sdf(const SeedParams &s, float x, float y, float z) {
    const float cx = 0.5f * (s.boxLo[0] + s.boxHi[0]), cy = 0.5f * (s.boxLo[1] + s.boxHi[1]), cz = 0.5f * (s.boxLo[2] + s.boxHi[2]);
    const float hx = 0.5f * (s.boxHi[0] - s.boxLo[0]), hy = 0.5f * (s.boxHi[1] - s.boxLo[1]), hz = 0.5f * (s.boxHi[2] - s.boxLo[2]);
    const float qx = fabsf(x - cx) - hx, qy = fabsf(y - cy) - hy, qz = fabsf(z - cz) - hz;
    const float ox = fmaxf(qx, 0.0f), oy = fmaxf(qy, 0.0f), oz = fmaxf(qz, 0.0f);
    return sqrtf(ox * ox + oy * oy + oz * oz) + fminf(fmaxf(qx, fmaxf(qy, qz)), 0.0f);
}

// the object's SDF at a point: Interpolation::trilinearInterpolate of the nodal field (meshlevelset.cpp:261-263)
__device__ __forceinline__ float object_sdf(const SeedParams &s, const float *__restrict__ sdf, float x, float y, float z) {
    if (sdf) return sample_scalar(sdf, s.I + 1, s.J + 1, s.Kg + 1, s.dx, s.invdx, x, y, z, 0);
    // analytic box: the same trilinear blend of the eight nodal distances around the point
    ScalarSample q;
    q.i = pos2idx(x, s.invdx); q.j = pos2idx(y, s.invdx); q.k = pos2idx(z, s.invdx);
    const float gx = (float)dmul((double)(float)q.i, s.dx), gy = (float)dmul((double)(float)q.j, s.dx), gz = (float)dmul((double)(float)q.k, s.dx);
    q.fx = dmul((double)fsub(x, gx), s.invdx); q.fy = dmul((double)fsub(y, gy), s.invdx); q.fz = dmul((double)fsub(z, gz), s.invdx);
    const float dxf = (float)s.dx;
    // vertex order 000,100,010,001,101,011,110,111
    const int ox[8] = {0, 1, 0, 0, 1, 0, 1, 1}, oy[8] = {0, 0, 1, 0, 0, 1, 1, 1}, oz[8] = {0, 0, 0, 1, 1, 1, 0, 1};
#pragma unroll
    for (int m = 0; m < 8; m++) q.v[m] = box_sdf(s, gx + ox[m] * dxf, gy + oy[m] * dxf, gz + oz[m] * dxf);
    return scalar_value(q);
}

// MeshObject::getCells (meshobject.cpp:101-140): the cells of an object are those with a corner node inside it
__device__ __forceinline__ bool cell_touches_object(const SeedParams &s, const float *__restrict__ sdf, int i, int j, int kg) {
    const float dxf = (float)s.dx;
    const float x0 = (float)dmul((double)(float)i, s.dx), y0 = (float)dmul((double)(float)j, s.dx), z0 = (float)dmul((double)(float)kg, s.dx);
    bool inside = false;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int ni = i + (m & 1), nj = j + ((m >> 1) & 1), nk = kg + (m >> 2);
        const float v = sdf ? __ldg(sdf + (size_t)ni + (size_t)(s.I + 1) * (nj + (size_t)(s.J + 1) * nk))
                            : box_sdf(s, x0 + (m & 1) * dxf, y0 + ((m >> 1) & 1) * dxf, z0 + (m >> 2) * dxf);
        inside |= v <= 0.0f;
    }
    return inside;
}

// The source's level set as the velocity constraint of an inflow reads it (fluidsimulation.cpp:3401, :4138): the grid
// of MeshFluidSource::update, which starts at cell org, sampled at the WORLD position as if it started at the origin;
// nodes outside it contribute 0 (Interpolation::trilinearInterpolate, interpolation.cpp:72-112).
__device__ __forceinline__ float source_sdf_unoffset(const SeedParams &s, const float *__restrict__ sdf, float x, float y, float z) {
    ScalarSample q;
    q.i = pos2idx(x, s.invdx); q.j = pos2idx(y, s.invdx); q.k = pos2idx(z, s.invdx);
    const float gx = (float)dmul((double)(float)q.i, s.dx), gy = (float)dmul((double)(float)q.j, s.dx), gz = (float)dmul((double)(float)q.k, s.dx);
    q.fx = dmul((double)fsub(x, gx), s.invdx); q.fy = dmul((double)fsub(y, gy), s.invdx); q.fz = dmul((double)fsub(z, gz), s.invdx);
    const int ox[8] = {0, 1, 0, 0, 1, 0, 1, 1}, oy[8] = {0, 0, 1, 0, 0, 1, 1, 1}, oz[8] = {0, 0, 0, 1, 1, 1, 0, 1};
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int li = q.i + ox[m], lj = q.j + oy[m], lk = q.k + oz[m];
        float v = 0.0f;
        if (li >= 0 && lj >= 0 && lk >= 0 && li < s.nodes[0] && lj < s.nodes[1] && lk < s.nodes[2]) {
            const int ni = li + s.org[0], nj = lj + s.org[1], nk = lk + s.org[2];       // the node of the domain grid it is
            v = sdf ? __ldg(sdf + (size_t)ni + (size_t)(s.I + 1) * (nj + (size_t)(s.J + 1) * nk))
                    : box_sdf(s, (float)dmul((double)(float)ni, s.dx), (float)dmul((double)(float)nj, s.dx), (float)dmul((double)(float)nk, s.dx));
        }
        q.v[m] = v;
    }
    return scalar_value(q);
}

// One thread per cell of the object's cell range: the eight sub-cell points (_addNewFluidCellsThread :4479-4518, then
// the mask test of _addNewFluidCells :4420-4427).
__global__ void k_seed_emit(SeedParams s, const float *__restrict__ sdf, const float *__restrict__ phiS, unsigned int *__restrict__ mask,
                            ParticleSoA out, int base, int cap, int *__restrict__ count, int *__restrict__ idOut, int idBase) {
    const int ni = s.hi[0] - s.lo[0], nj = s.hi[1] - s.lo[1], nk = s.hi[2] - s.lo[2];
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)ni * nj * nk) return;
    const int i = s.lo[0] + (int)(t % ni), j = s.lo[1] + (int)((t / ni) % nj), kg = s.lo[2] + (int)(t / ((long long)ni * nj));
    const int kl = kg - s.kOff;
    if (kl < s.kOwn0 || kl >= s.kOwn1) return;          // z-slab: every rank seeds its own planes
    // Grid3d::GridIndexToCellCenter (grid3d.h:106-109): (float)i*dx + hw in double, narrowed to float
    const double hw = 0.5 * s.dx;
    const float cx = (float)dadd(dmul((double)(float)i, s.dx), hw), cy = (float)dadd(dmul((double)(float)j, s.dx), hw),
                cz = (float)dadd(dmul((double)(float)kg, s.dx), hw);
    const float q = (float)(0.25 * s.dx);
    const size_t cell = (size_t)i + (size_t)s.I * (j + (size_t)s.J * kl);
    if (!cell_touches_object(s, sdf, i, j, kg)) return;      // only such cells are candidates
    const unsigned int have = (mask[cell >> 2] >> (8 * (cell & 3))) & 0xffu;
    unsigned int added = 0u;
#pragma unroll
    for (int o = 0; o < 8; o++) {
        const float x = fadd(cx, (o & 1) ? q : -q), y = fadd(cy, (o & 2) ? q : -q), z = fadd(cz, (o & 4) ? q : -q);
        int si, sj, sk;
        unsigned int bit;
        subcell_of(s, x, y, z, si, sj, sk, bit);
        bool keep = !((have | added) & bit);                       // maskgrid.isSubCellSet(p)
        if (keep) keep = !(object_sdf(s, sdf, x, y, z) > 0.0f);    // d > 0: outside the object
        // MeshLevelSet::trilinearInterpolate of the solid SDF (:4511): inside the solid -> dropped
        if (keep) keep = sample_scalar(phiS, s.I + 1, s.J + 1, s.K + 1, s.dx, s.invdx, x, y, z, s.kOff) > 0.0f;
        if (keep) {
            const int slot = warp_append_slot(count);
            if (slot < cap) {
                out.px[base + slot] = x; out.py[base + slot] = y; out.pz[base + slot] = z;
                out.vx[base + slot] = s.vel[0]; out.vy[base + slot] = s.vel[1]; out.vz[base + slot] = s.vel[2];
                if (idOut) idOut[base + slot] = idBase + slot;
            }
            added |= bit;
        }
    }
    if (added) atomicOr(&mask[cell >> 2], added << (8 * (cell & 3)));   // maskgrid.addParticle(p): later objects of the queue see it
}

// _updateOutflowMeshFluidSource (:4607-4668, not inversed): particles of the source's cells with object SDF < 0 are
// removed -- marked here by moving them out of the domain, the sort that follows drops them
__global__ void k_seed_outflow(ParticleSoA p, int n, SeedParams s, const float *__restrict__ sdf, int *__restrict__ removed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float x = p.px[t], y = p.py[t], z = p.pz[t];
    const int i = pos2idx(x, s.invdx), j = pos2idx(y, s.invdx), k = pos2idx(z, s.invdx);
    if (i < s.lo[0] || i >= s.hi[0] || j < s.lo[1] || j >= s.hi[1] || k < s.lo[2] || k >= s.hi[2]) return;
    if (!cell_touches_object(s, sdf, i, j, k)) return;      // isOutflowCell
    if (object_sdf(s, sdf, x, y, z) < 0.0f) {
        p.px[t] = -1.0e30f;
        atomicAdd(removed, 1);
    }
}

// Inflow sources with a constrained fluid velocity (MeshFluidSource::_isConstrainedFluidVelocity, default on):
// (1) _getInflowConstrainedVelocityComponents (:3372-3436) + _applyConstantBodyForces: faces that carry a P2G value and
// whose centre has object SDF < 0 get no body force -- here the force has been added to the whole field already and
// the saved (pre-force) value is put back on those faces, which is the same float;
__global__ void k_inflow_restore_faces(SeedParams s, const float *__restrict__ sdf, int axis, float *__restrict__ field,
                                       const float *__restrict__ saved, const unsigned char *__restrict__ valid) {
    const int e0 = axis == 0, e1 = axis == 1, e2 = axis == 2;
    // where the un-offset read of the source's grid can be negative: the object's cells moved by -org
    const int l0 = max(s.lo[0] - s.org[0] - 1, 0), l1 = max(s.lo[1] - s.org[1] - 1, 0), l2 = max(s.lo[2] - s.org[2] - 1, s.kOff);
    const int h0 = min(s.hi[0] - s.org[0] + 1, s.I) + e0, h1 = min(s.hi[1] - s.org[1] + 1, s.J) + e1,
              h2 = min(s.hi[2] - s.org[2] + 1, s.kOff + s.K) + e2;
    const int ni = h0 - l0, nj = h1 - l1, nk = h2 - l2;
    if (ni <= 0 || nj <= 0 || nk <= 0) return;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)ni * nj * nk) return;
    const int i = l0 + (int)(t % ni), j = l1 + (int)((t / ni) % nj), kg = l2 + (int)(t / ((long long)ni * nj));
    const size_t f = (size_t)i + (size_t)(s.I + e0) * (j + (size_t)(s.J + e1) * (kg - s.kOff));
    if (!valid[f]) return;
    // Grid3d::FaceIndexToPositionU/V/W (grid3d.h:111-133)
    const float x = (float)dmul(e0 ? (double)(float)i : dadd((double)(float)i, 0.5), s.dx);
    const float y = (float)dmul(e1 ? (double)(float)j : dadd((double)(float)j, 0.5), s.dx);
    const float z = (float)dmul(e2 ? (double)(float)kg : dadd((double)(float)kg, 0.5), s.dx);
    if (source_sdf_unoffset(s, sdf, x, y, z) < 0.0f) field[f] = saved[f];
}

// (2) _constrainMarkerParticleVelocities (:4113-4159), after the PIC/FLIP update: particles of the source's cells with
// object SDF <= 0 carry the source's velocity.
__global__ void k_inflow_constrain_particles(ParticleSoA p, int n, SeedParams s, const float *__restrict__ sdf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float x = p.px[t], y = p.py[t], z = p.pz[t];
    const int i = pos2idx(x, s.invdx), j = pos2idx(y, s.invdx), k = pos2idx(z, s.invdx);
    if (i < s.lo[0] || i >= s.hi[0] || j < s.lo[1] || j >= s.hi[1] || k < s.lo[2] || k >= s.hi[2]) return;
    if (!cell_touches_object(s, sdf, i, j, k)) return;
    if (source_sdf_unoffset(s, sdf, x, y, z) > 0.0f) return;
    p.vx[t] = s.vel[0]; p.vy[t] = s.vel[1]; p.vz[t] = s.vel[2];
}

static void fill_params(flip_ctx *c, SeedParams &s) {
    const Dims &d = c->d;
    s.I = d.I; s.J = d.J; s.K = d.K; s.Kg = d.Kg; s.kOff = d.kOff; s.kOwn0 = d.kOwn0; s.kOwn1 = d.kOwn1;
    s.dx = d.dx; s.invdx = 1.0 / d.dx; s.invsubdx = 1.0 / (0.5 * d.dx);
}

static long long object_params(flip_ctx *c, flip_ctx::FluidObject &o, SeedParams &s) {
    long long cells = 1;
    for (int a = 0; a < 3; a++) {
        s.lo[a] = o.lo[a]; s.hi[a] = o.hi[a];
        s.boxLo[a] = (float)o.boxLo[a]; s.boxHi[a] = (float)o.boxHi[a]; s.vel[a] = (float)o.vel[a];
        s.org[a] = o.sdfOrigin[a]; s.nodes[a] = o.sdfNodes[a];
        cells *= std::max(0, o.hi[a] - o.lo[a]);
    }
    if (!o.sdf.empty() && !o.dsdf) {          // the nodal field goes to the device once
        FLIP_CUDA_CHECK(cudaMalloc(&o.dsdf, sizeof(float) * o.sdf.size()));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(o.dsdf, o.sdf.data(), sizeof(float) * o.sdf.size(), cudaMemcpyHostToDevice, c->stream));
    }
    return cells;
}

bool has_constrained_inflow(const flip_ctx *c) {
    for (auto &o : c->fluidObjects)
        if (o.kind == 1 && o.enabled && o.constrained) return true;
    return false;
}

// after stage_body_force (the saved field still holds the pre-force values: stage_save runs before it)
void stage_inflow_body_force_exclusion(flip_ctx *c) {
    if (!has_constrained_inflow(c)) return;
    const Dims &d = c->d;
    SeedParams s;
    fill_params(c, s);
    float *field[3] = {c->U, c->V, c->W};
    const float *saved[3] = {c->sU, c->sV, c->sW};
    const unsigned char *valid[3] = {c->validU, c->validV, c->validW};
    for (auto &o : c->fluidObjects) {
        if (o.kind != 1 || !o.enabled || !o.constrained) continue;
        if (object_params(c, o, s) == 0) continue;
        const long long box = (long long)(o.hi[0] - o.lo[0] + 3) * (o.hi[1] - o.lo[1] + 3) * (o.hi[2] - o.lo[2] + 3);
        for (int a = 0; a < 3; a++) {
            if (fabs((float)c->gravity[a]) <= 1e-6f) continue;          // no force was added on this axis
            k_inflow_restore_faces<<<cdiv(box, TPB), TPB, 0, c->stream>>>(s, o.dsdf, a, field[a], saved[a], valid[a]);
            c->launches++;
        }
    }
    FLIP_CUDA_CHECK(cudaGetLastError());
}

// after the G2P stage (the fused G2P + advance pass is not used while such a source is enabled: the CFL speed of the
// next substep reads the constrained velocities)
void stage_inflow_constrain_particles(flip_ctx *c) {
    if (!has_constrained_inflow(c) || c->npStore == 0) return;
    SeedParams s;
    fill_params(c, s);
    for (auto &o : c->fluidObjects) {
        if (o.kind != 1 || !o.enabled || !o.constrained) continue;
        if (object_params(c, o, s) == 0) continue;
        k_inflow_constrain_particles<<<cdiv(c->npStore, TPB), TPB, 0, c->stream>>>(c->P[c->cur_buf], c->npStore, s, o.dsdf);
        c->launches++;
    }
    FLIP_CUDA_CHECK(cudaGetLastError());
}

// _updateFluidObjects (:4794-4806), called at the end of a substep (after the advection and its sort, :5504):
// _updateAddedFluidMeshObjectQueue (the queued objects, once), then _updateMeshFluidSources (inflow sources emit at
// every substep -- substep emissions 1, meshfluidsource.h:119 --, outflow sources remove).
void stage_fluid_objects(flip_ctx *c) {
    bool any = false;
    for (auto &o : c->fluidObjects) any = any || o.kind == 0 || o.enabled;
    if (!any) return;
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    SeedParams s;
    fill_params(c, s);
    // the sub-cell masks live in the extrapolation scratch (idle here), one byte per cell
    unsigned int *mask = reinterpret_cast<unsigned int *>(c->status);
    FLIP_CUDA_CHECK(cudaMemsetAsync(mask, 0, (((size_t)d.nC + 3) / 4) * 4, st));
    // upper bound of what can be added: eight per cell of every emitting object's cell range
    long long bound = 0;
    for (auto &o : c->fluidObjects) {
        if (o.kind == 2 || (o.kind == 1 && !o.enabled)) continue;
        long long cells = 1;
        for (int a = 0; a < 3; a++) cells *= std::max(0, o.hi[a] - o.lo[a]);
        bound += 8 * cells;
    }
    if (c->np + bound >= (1ll << 30)) throw ApiError(FLIP_ERR_UNSUPPORTED, "more than 2^30 particles per GPU");
    particles_alloc(c, (int)(c->np + bound));        // keeps the live particles
    int *count = &c->dS->deferredCount, *removed = &c->dS->frontierCount[0];
    FLIP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
    FLIP_CUDA_CHECK(cudaMemsetAsync(removed, 0, sizeof(int), st));
    ParticleSoA P = c->P[c->cur_buf];
    if (c->np > 0) { k_seed_mask<<<cdiv(c->np, TPB), TPB, 0, st>>>(P, c->np, s, mask); c->launches++; }
    const int cap = c->capacity - c->np;
    auto params_of = [&](flip_ctx::FluidObject &o) { return object_params(c, o, s); };
    // the queue first, then the inflow sources (the mask carries what the earlier ones added), then the outflow sources
    for (int pass = 0; pass < 2; pass++)
        for (auto &o : c->fluidObjects) {
            if (o.kind != pass || (o.kind == 1 && !o.enabled)) continue;
            const long long cells = params_of(o);
            if (cells == 0) continue;
            k_seed_emit<<<cdiv(cells, TPB), TPB, 0, st>>>(s, o.dsdf, c->phiS, mask, P, c->np, cap, count,
                                                          c->trackIds ? c->pid[c->cur_buf] : nullptr, c->nextParticleId);
            c->launches++;
        }
    scalars_to_host(c);
    const int added = std::min(c->hS->deferredCount, cap);
    FLIP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
    c->nextParticleId += added;
    c->np += added;
    c->npStore = c->np;
    bool outflow = false;
    for (auto &o : c->fluidObjects) {
        if (o.kind != 2 || !o.enabled || c->np == 0) continue;
        if (params_of(o) == 0) continue;
        k_seed_outflow<<<cdiv(c->np, TPB), TPB, 0, st>>>(P, c->np, s, o.dsdf, removed);
        c->launches++;
        outflow = true;
    }
    // the queued objects are done; sources stay
    for (size_t q = 0; q < c->fluidObjects.size();) {
        if (c->fluidObjects[q].kind == 0) {
            FLIP_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(c->fluidObjects[q].dsdf);
            c->fluidObjects.erase(c->fluidObjects.begin() + q);
        } else q++;
    }
    int gone = 0;
    if (outflow) { scalars_to_host(c); gone = c->hS->frontierCount[0]; }
    if (added == 0 && gone == 0) return;
    // the store is cell-sorted again (no removal rules: _addMarkerParticle only range-checks, :2637-2642; the particles an
    // outflow source took lie outside the domain now and are dropped)
    particles_sort(c, false, 0.0, 0, c->np, slab_on(c));
}

}  // namespace flip
