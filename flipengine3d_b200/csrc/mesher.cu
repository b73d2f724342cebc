// Surface reconstruction on the device: what FluidSimulation::getIsomesh() returns, i.e.
// _polygonizeOutputSurface -> ParticleMesher::meshParticles (fluidsimulation.cpp:5043-5099, particlemesher.cpp:36-56)
// followed by TriangleMesh::smooth (trianglemesh.cpp:536-583), at the engine's default settings: one compute chunk,
// obstacle meshing offset 0 (the solid SDF of the simulation is used as it is), no minimum polyhedron size, no
// removal near the domain, no inverted contact normals, smoothing 0.5 x 2 iterations.
//
// Reference algorithm restated:
//   * scalar field on the nodes of the grid subdivided `s` times ((I s + 1)(J s + 1)(K s + 1) nodes, spacing dx/s):
//     value = -min(3 r, min over the particles whose index box covers the node of (|node - p| - r)), r = 3 x the
//     marker particle radius (fluidsimulation.cpp:2644-2648, :5083); a particle's box spans the node indices
//     floor((p - 1.5 r)/subdx) .. floor((p + 1.5 r)/subdx) + 1 per axis, in coordinates local to blocks of ten nodes
//     (particlemesher.cpp:582-640) -- evaluated here in the same block-local float arithmetic;
//   * nodes inside the solid are clamped to the threshold 0 (ScalarField::getScalarFieldValue, scalarfield.cpp:439-448;
//     inside = trilinear sample of the solid SDF < 0);
//   * marching cubes over the cells that have an inside (> 0) and an outside node (polygonizer3d.cpp:366-425,643-676),
//     one welded vertex per crossed grid edge, placed by _vertexInterp (:451-487) with its solid-boundary clamp of the
//     interpolation parameter;
//   * two Laplacian smoothing passes: v += 0.5 (mean of the other two vertices of every incident triangle - v).
// Here: a cheap inside/outside pass over all nodes (early exit at the first particle nearer than r; empty space is
// skipped through the 4^3-cell occupancy blocks of the sort), exact values only for the nodes of surface cells (pruned
// nearest-first search over the cell-sorted particles), vertex and triangle counts through CUB scans, and the
// smoothing sums in 64-bit fixed point (integer atomics: deterministic).  The case table is generated at start-up
// (faces traced into loops, fan-triangulated; ambiguous faces separate the inside corners -- the same surface
// topology class as the classic table, triangle counts may differ in the ambiguous cases).
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "device_math.cuh"
#include "flip_internal.h"
#include "mc_tables.h"

namespace flip {

static constexpr int TPB = 256;

// corner c of a cube: bit 0 = +x, bit 1 = +y, bit 2 = +z.  Edge e = axis * 4 + (the two other bits of its lower
// corner, lower axis first); face-traced case table: count and edge triples per sign configuration.
__constant__ unsigned char c_mcCount[256];
__constant__ unsigned char c_mcTris[256][MC_MAX_TRIS * 3];

struct IsoParams {
    int I, J, K, s;                // cells, subdivision
    int ni, nj, nk;                // sub-nodes per axis
    double dx, invdx, subdx, invsubdx, invBlockdx;   // invBlockdx: 1 / (float)(10 subdx)  particlemesher.cpp:490
    float r, sr, maxd;             // particle radius, search radius, 3 r
    int oI, oJ, oK;                // occupancy blocks
};

// block-local geometry of a node along one axis (blocks of ten nodes, particlemesher.cpp:599-601)
struct IsoAxis {
    float B;      // block offset  (float)(block * 10 * subdx)
    float g;      // node position inside the block  (float)(l * subdx)
    int l;
};
__device__ __forceinline__ IsoAxis iso_axis(const IsoParams &P, int idx) {
    IsoAxis a;
    const int b = idx / 10;
    a.l = idx - 10 * b;
    a.B = (float)dmul((double)(float)b, dmul(10.0, P.subdx));      // Grid3d::GridIndexToPosition(block, 10 subdx)
    a.g = (float)dmul((double)(float)a.l, P.subdx);
    return a;
}

// |gpos - p| - r in the reference's arithmetic (vmath::length, float), p local to the node's block
__device__ __forceinline__ float iso_dist(const IsoAxis &ax, const IsoAxis &ay, const IsoAxis &az, float x, float y, float z, float r) {
    const float px = fsub(x, ax.B), py = fsub(y, ay.B), pz = fsub(z, az.B);
    return fsub(length3(fsub(ax.g, px), fsub(ay.g, py), fsub(az.g, pz)), r);
}

// ScalarField::_isVertexSolid for a sub-node: the solid SDF at the node (matching grids) or its trilinear sample
__device__ __forceinline__ bool iso_solid(const IsoParams &P, const float *__restrict__ phiS, int si, int sj, int sk) {
    if (P.s == 1) return __ldg(phiS + (size_t)si + (size_t)(P.I + 1) * (sj + (size_t)(P.J + 1) * sk)) < 0.0f;
    const float x = (float)dmul((double)(float)si, P.subdx), y = (float)dmul((double)(float)sj, P.subdx), z = (float)dmul((double)(float)sk, P.subdx);
    return sample_scalar(phiS, P.I + 1, P.J + 1, P.K + 1, P.dx, P.invdx, x, y, z, 0) < 0.0f;
}

// pass 1: inside (> 0 and not solid) flag of every node
__global__ void k_iso_inside(IsoParams P, ParticleSoA p, const int *__restrict__ cellStart, const unsigned char *__restrict__ occ,
                             const float *__restrict__ phiS, unsigned char *__restrict__ inside) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)P.ni * P.nj * P.nk) return;
    const int si = (int)(t % P.ni), sj = (int)((t / P.ni) % P.nj), sk = (int)(t / ((long long)P.ni * P.nj));
    const IsoAxis ax = iso_axis(P, si), ay = iso_axis(P, sj), az = iso_axis(P, sk);
    const float gx = (float)dmul((double)(float)si, P.subdx), gy = (float)dmul((double)(float)sj, P.subdx), gz = (float)dmul((double)(float)sk, P.subdx);
    // cells that can hold a particle nearer than r (< dx)
    const int ilo = max((int)floorf((gx - P.r) * (float)P.invdx) - 0, 0), ihi = min((int)floorf((gx + P.r) * (float)P.invdx), P.I - 1);
    const int jlo = max((int)floorf((gy - P.r) * (float)P.invdx) - 0, 0), jhi = min((int)floorf((gy + P.r) * (float)P.invdx), P.J - 1);
    const int klo = max((int)floorf((gz - P.r) * (float)P.invdx) - 0, 0), khi = min((int)floorf((gz + P.r) * (float)P.invdx), P.K - 1);
    bool in = false;
    // empty space: the occupancy blocks (4^3 cells) that those cells touch
    bool any = false;
    for (int bk = klo >> 2; bk <= (khi >> 2); bk++)
        for (int bj = jlo >> 2; bj <= (jhi >> 2); bj++)
            for (int bi = ilo >> 2; bi <= (ihi >> 2); bi++) any |= __ldg(occ + bi + P.oI * (bj + P.oJ * bk)) != 0;
    if (any) {
        for (int ck = klo; ck <= khi && !in; ck++)
            for (int cj = jlo; cj <= jhi && !in; cj++) {
                const int rowBase = P.I * (cj + P.J * ck);
                const int qb = __ldg(cellStart + rowBase + ilo), qe = __ldg(cellStart + rowBase + ihi + 1);
                for (int q = qb; q < qe; q++)
                    if (iso_dist(ax, ay, az, __ldg(p.px + q), __ldg(p.py + q), __ldg(p.pz + q), P.r) < 0.0f) { in = true; break; }
            }
    }
    if (in && iso_solid(P, phiS, si, sj, sk)) in = false;
    inside[t] = in ? 1 : 0;
}

// pass 2: sub-cells with an inside and an outside node: triangle count, and their nodes are marked for exact values
__global__ void k_iso_cells(IsoParams P, const unsigned char *__restrict__ inside, unsigned char *__restrict__ need, int *__restrict__ triCount) {
    const int ci = P.ni - 1, cj = P.nj - 1, ck = P.nk - 1;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)ci * cj * ck) return;
    const int i = (int)(t % ci), j = (int)((t / ci) % cj), k = (int)(t / ((long long)ci * cj));
    int cfg = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const size_t n = (size_t)(i + (c & 1)) + (size_t)P.ni * ((j + ((c >> 1) & 1)) + (size_t)P.nj * (k + (c >> 2)));
        cfg |= inside[n] ? (1 << c) : 0;
    }
    int cnt = 0;
    if (cfg != 0 && cfg != 255) {
        cnt = c_mcCount[cfg];
#pragma unroll
        for (int c = 0; c < 8; c++)
            need[(size_t)(i + (c & 1)) + (size_t)P.ni * ((j + ((c >> 1) & 1)) + (size_t)P.nj * (k + (c >> 2)))] = 1;
    }
    triCount[t] = cnt;
}

// pass 3: the exact field value of the marked nodes
__global__ void k_iso_values(IsoParams P, ParticleSoA p, const int *__restrict__ cellStart, const float *__restrict__ phiS,
                             const unsigned char *__restrict__ need, float *__restrict__ value) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)P.ni * P.nj * P.nk) return;
    if (!need[t]) return;
    const int si = (int)(t % P.ni), sj = (int)((t / P.ni) % P.nj), sk = (int)(t / ((long long)P.ni * P.nj));
    const IsoAxis ax = iso_axis(P, si), ay = iso_axis(P, sj), az = iso_axis(P, sk);
    const float gx = (float)dmul((double)(float)si, P.subdx), gy = (float)dmul((double)(float)sj, P.subdx), gz = (float)dmul((double)(float)sk, P.subdx);
    const float dxf = (float)P.dx, reachLo = P.sr + (float)P.subdx + 1.0e-3f * dxf, reachHi = reachLo;
    // cells whose particles can cover the node: floor((p - sr)/subdx) <= l <= floor((p + sr)/subdx) + 1, i.e. p within
    // sr + subdx of the node on either side
    const int ilo = max((int)floorf((gx - reachLo) * (float)P.invdx), 0), ihi = min((int)floorf((gx + reachHi) * (float)P.invdx), P.I - 1);
    const int jlo = max((int)floorf((gy - reachLo) * (float)P.invdx), 0), jhi = min((int)floorf((gy + reachHi) * (float)P.invdx), P.J - 1);
    const int klo = max((int)floorf((gz - reachLo) * (float)P.invdx), 0), khi = min((int)floorf((gz + reachHi) * (float)P.invdx), P.K - 1);
    float best = P.maxd;          // min(3 r, ...)
    const int bx = si / 10, by = sj / 10, bz = sk / 10;       // the node's block
    for (int ck = klo; ck <= khi; ck++) {
        const float dz = fmaxf(fmaxf((float)ck * dxf - gz, gz - (float)(ck + 1) * dxf), 0.0f);
        for (int cj = jlo; cj <= jhi; cj++) {
            const float dy = fmaxf(fmaxf((float)cj * dxf - gy, gy - (float)(cj + 1) * dxf), 0.0f);
            // nothing in this row can be nearer than the best so far (margin against the rounding of the distances)
            if (sqrtf(dy * dy + dz * dz) - P.r > best + 1.0e-4f * dxf) continue;
            const int rowBase = P.I * (cj + P.J * ck);
            const int qb = __ldg(cellStart + rowBase + ilo), qe = __ldg(cellStart + rowBase + ihi + 1);
            for (int q = qb; q < qe; q++) {
                const float x = __ldg(p.px + q), y = __ldg(p.py + q), z = __ldg(p.pz + q);
                // the particle's index box in the node's block (particlemesher.cpp:603-612)
                const float px = fsub(x, ax.B), py = fsub(y, ay.B), pz = fsub(z, az.B);
                // ... of the blocks the particle was sorted into: those its [p - sr, p + sr] overlaps (:490-527)
                if (bx < pos2idx(fsub(x, P.sr), P.invBlockdx) || bx > pos2idx(fadd(x, P.sr), P.invBlockdx) ||
                    by < pos2idx(fsub(y, P.sr), P.invBlockdx) || by > pos2idx(fadd(y, P.sr), P.invBlockdx) ||
                    bz < pos2idx(fsub(z, P.sr), P.invBlockdx) || bz > pos2idx(fadd(z, P.sr), P.invBlockdx)) continue;
                const int ax0 = pos2idx(fsub(px, P.sr), P.invsubdx), ax1 = pos2idx(fadd(px, P.sr), P.invsubdx) + 1;
                const int ay0 = pos2idx(fsub(py, P.sr), P.invsubdx), ay1 = pos2idx(fadd(py, P.sr), P.invsubdx) + 1;
                const int az0 = pos2idx(fsub(pz, P.sr), P.invsubdx), az1 = pos2idx(fadd(pz, P.sr), P.invsubdx) + 1;
                if (ax.l < ax0 || ax.l > ax1 || ay.l < ay0 || ay.l > ay1 || az.l < az0 || az.l > az1) continue;
                best = fminf(best, fsub(length3(fsub(ax.g, px), fsub(ay.g, py), fsub(az.g, pz)), P.r));
            }
        }
    }
    float v = -best;
    if (v > 0.0f && iso_solid(P, phiS, si, sj, sk)) v = 0.0f;
    value[t] = v;
}

// pass 4a: crossed grid edges (node -> node + axis): one vertex each
__global__ void k_iso_edge_count(IsoParams P, const unsigned char *__restrict__ inside, int *__restrict__ edgeFlag) {
    const long long n = (long long)P.ni * P.nj * P.nk;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int a = (int)(t / n);
    const long long node = t - a * n;
    const int si = (int)(node % P.ni), sj = (int)((node / P.ni) % P.nj), sk = (int)(node / ((long long)P.ni * P.nj));
    const int ti = si + (a == 0), tj = sj + (a == 1), tk = sk + (a == 2);
    int f = 0;
    if (ti < P.ni && tj < P.nj && tk < P.nk) f = inside[node] != inside[(size_t)ti + (size_t)P.ni * (tj + (size_t)P.nj * tk)];
    edgeFlag[t] = f;
}

// Polygonizer3d::_vertexInterp (polygonizer3d.cpp:451-487)
__device__ __forceinline__ void iso_vertex(const IsoParams &P, const float *__restrict__ phiS, const float p1[3], const float p2[3],
                                           double v1, double v2, float out[3]) {
    double minmu = 0.0, maxmu = 1.0;
    const double eps = 1e-10;
    const double s1 = (double)sample_scalar(phiS, P.I + 1, P.J + 1, P.K + 1, P.dx, P.invdx, p1[0], p1[1], p1[2], 0);
    const double s2 = (double)sample_scalar(phiS, P.I + 1, P.J + 1, P.K + 1, P.dx, P.invdx, p2[0], p2[1], p2[2], 0);
    if ((s1 < 0.0 && s2 >= 0.0) || (s2 < 0.0 && s1 >= 0.0)) {
        const double diff = s2 - s1;
        if (fabs(diff) > eps) {
            const double su = -s1 / diff;
            if (s1 < 0.0) minmu = su; else maxmu = su;
        } else {
            maxmu = minmu;
        }
    }
    minmu = fmax(minmu, eps);
    maxmu = fmin(maxmu, 1.0 - eps);
    double mu = (0.0 - v1) / (v2 - v1);
    if (mu < minmu) mu = minmu;
    if (mu > maxmu) mu = maxmu;
    const float m = (float)mu;
#pragma unroll
    for (int x = 0; x < 3; x++) out[x] = fadd(p1[x], fmul(m, fsub(p2[x], p1[x])));
}

__global__ void k_iso_vertices(IsoParams P, const float *__restrict__ phiS, const int *__restrict__ edgeFlag, const int *__restrict__ edgeIdx,
                               const float *__restrict__ value, float *__restrict__ verts) {
    const long long n = (long long)P.ni * P.nj * P.nk;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= 3 * n || !edgeFlag[t]) return;
    const int a = (int)(t / n);
    const long long node = t - a * n;
    const int si = (int)(node % P.ni), sj = (int)((node / P.ni) % P.nj), sk = (int)(node / ((long long)P.ni * P.nj));
    const int ti = si + (a == 0), tj = sj + (a == 1), tk = sk + (a == 2);
    // _getVertexPosition: (float)dx * vec3((float)i, (float)j, (float)k)
    const float sdx = (float)P.subdx;
    const float p1[3] = {fmul(sdx, (float)si), fmul(sdx, (float)sj), fmul(sdx, (float)sk)};
    const float p2[3] = {fmul(sdx, (float)ti), fmul(sdx, (float)tj), fmul(sdx, (float)tk)};
    float out[3];
    iso_vertex(P, phiS, p1, p2, (double)value[node], (double)value[(size_t)ti + (size_t)P.ni * (tj + (size_t)P.nj * tk)], out);
    const int v = edgeIdx[t];
    verts[3ll * v] = out[0]; verts[3ll * v + 1] = out[1]; verts[3ll * v + 2] = out[2];
}

// pass 5: triangles of the surface cells
__global__ void k_iso_triangles(IsoParams P, const unsigned char *__restrict__ inside, const int *__restrict__ triCount,
                                const int *__restrict__ triStart, const int *__restrict__ edgeIdx, int *__restrict__ tris) {
    const int ci = P.ni - 1, cj = P.nj - 1, ck = P.nk - 1;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)ci * cj * ck) return;
    const int cnt = triCount[t];
    if (!cnt) return;
    const int i = (int)(t % ci), j = (int)((t / ci) % cj), k = (int)(t / ((long long)ci * cj));
    const long long n = (long long)P.ni * P.nj * P.nk;
    int cfg = 0;
#pragma unroll
    for (int c = 0; c < 8; c++)
        cfg |= inside[(size_t)(i + (c & 1)) + (size_t)P.ni * ((j + ((c >> 1) & 1)) + (size_t)P.nj * (k + (c >> 2)))] ? (1 << c) : 0;
    const int o = triStart[t];
    for (int q = 0; q < cnt * 3; q++) {
        const int e = c_mcTris[cfg][q];
        const int a = e >> 2, b = e & 3;
        // lower corner of edge e: the two other axes take the bits of b, lower axis first
        const int o1 = (a + 1) % 3, o2 = (a + 2) % 3, l1 = min(o1, o2), l2 = max(o1, o2);
        int off[3] = {0, 0, 0};
        off[l1] = b & 1; off[l2] = b >> 1;
        const long long node = (long long)(i + off[0]) + (long long)P.ni * ((j + off[1]) + (long long)P.nj * (k + off[2]));
        tris[3ll * o + q] = edgeIdx[a * n + node];
    }
}

// pass 6: TriangleMesh::_smoothTriangleMesh -- sums of the other two vertices of every incident triangle
__global__ void k_iso_smooth_accumulate(int nt, const int *__restrict__ tris, const float *__restrict__ verts,
                                        long long *__restrict__ acc, int *__restrict__ cnt) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int v[3] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
    long long q[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int x = 0; x < 3; x++) q[a][x] = __double2ll_rn((double)verts[3ll * v[a] + x] * 1073741824.0);    // 2^30 per unit
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int b = (a + 1) % 3, c = (a + 2) % 3;
        if (v[b] == v[a] && v[c] == v[a]) continue;
#pragma unroll
        for (int x = 0; x < 3; x++) {
            long long s = 0;
            if (v[b] != v[a]) s += q[b][x];
            if (v[c] != v[a]) s += q[c][x];
            atomicAdd(reinterpret_cast<unsigned long long *>(&acc[3ll * v[a] + x]), (unsigned long long)s);
        }
        atomicAdd(&cnt[v[a]], (v[b] != v[a]) + (v[c] != v[a]));
    }
}
__global__ void k_iso_smooth_apply(int nv, float value, float *__restrict__ verts, long long *__restrict__ acc, int *__restrict__ cnt) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int n = cnt[v];
    cnt[v] = 0;
#pragma unroll
    for (int x = 0; x < 3; x++) {
        const long long s = acc[3ll * v + x];
        acc[3ll * v + x] = 0;
        if (n > 0) {
            const float avg = __fdiv_rn((float)((double)s * (1.0 / 1073741824.0)), (float)n);
            const float p = verts[3ll * v + x];
            verts[3ll * v + x] = fadd(p, fmul(value, fsub(avg, p)));
        }
    }
}

struct MesherState {
    bool tablesUploaded = false;
    int64_t stamp = -1;           // step counter the cached mesh belongs to
    int nv = 0, nt = 0;
    float *verts = nullptr;
    int *tris = nullptr;
    size_t vertCap = 0, triCap = 0;
    // work arrays of a build, kept between calls while they are small (a viewer asks for the mesh every frame and
    // allocating ~10 arrays per call cost far more than the kernels: 33 ms against 2 ms at 128^3 x 2)
    char *scratch = nullptr;
    size_t scratchBytes = 0;
    char *smooth = nullptr;       // accumulators of the smoothing passes
    size_t smoothBytes = 0;
};
static constexpr size_t MESHER_KEEP_BYTES = 2ull << 30;     // larger work areas are released after the call

static void mesher_build(flip_ctx *c, float *hostValues = nullptr, unsigned char *hostInside = nullptr, unsigned char *hostNeed = nullptr) {
    MesherState *M = (MesherState *)c->mesher;
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    if (slab_on(c)) throw ApiError(FLIP_ERR_UNSUPPORTED, "surface reconstruction is per GPU: gather the particles on one rank");
    if (!M->tablesUploaded) {
        static unsigned char count[256], tris[256][MC_MAX_TRIS * 3];
        build_mc_tables(count, tris);
        FLIP_CUDA_CHECK(cudaMemcpyToSymbol(c_mcCount, count, sizeof(count)));
        FLIP_CUDA_CHECK(cudaMemcpyToSymbol(c_mcTris, tris, sizeof(tris)));
        M->tablesUploaded = true;
    }
    M->nv = M->nt = 0;
    if (c->np == 0) return;
    IsoParams P;
    P.I = d.I; P.J = d.J; P.K = d.K; P.s = std::max(1, c->surfaceSubdivision);
    P.ni = d.I * P.s + 1; P.nj = d.J * P.s + 1; P.nk = d.K * P.s + 1;
    P.dx = d.dx; P.invdx = 1.0 / d.dx; P.subdx = d.dx / (double)P.s; P.invsubdx = 1.0 / P.subdx;
    P.invBlockdx = 1.0 / (double)(float)(10 * P.subdx);
    // _initializeParticleRadii (:2644-2648) and params.radius = _markerParticleRadius * _markerParticleScale (:5083)
    const double volume = d.dx * d.dx * d.dx / 8.0, pi = 3.141592653;
    const double radius = pow(3 * volume / (4 * pi), 1.0 / 3.0) * c->markerParticleScale;     // default scale 3.0
    P.r = (float)radius; P.sr = 1.5f * P.r; P.maxd = (float)(3.0 * (double)P.r);
    P.oI = (d.I + 3) >> 2; P.oJ = (d.J + 3) >> 2; P.oK = (d.K + 3) >> 2;
    const long long nNodes = (long long)P.ni * P.nj * P.nk, nCells = (long long)(P.ni - 1) * (P.nj - 1) * (P.nk - 1);
    if (3 * nNodes >= (1ll << 31)) throw ApiError(FLIP_ERR_UNSUPPORTED, "surface grid too large for 32-bit edge indices");
    unsigned char *inside = nullptr, *need = nullptr;
    float *value = nullptr;
    int *triCount = nullptr, *triStart = nullptr, *edgeFlag = nullptr, *edgeIdx = nullptr;
    void *tmp = nullptr;
    size_t tmpBytes = 0;
    {
        size_t b1 = 0, b2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, b1, (int *)nullptr, (int *)nullptr, (int)(3 * nNodes + 1), st);
        cub::DeviceScan::ExclusiveSum(nullptr, b2, (int *)nullptr, (int *)nullptr, (int)(nCells + 1), st);
        tmpBytes = std::max(b1, b2);
    }
    auto release = [&] {
        if (M->scratchBytes > MESHER_KEEP_BYTES) { cudaFree(M->scratch); M->scratch = nullptr; M->scratchBytes = 0; }
    };
    try {
        // one area, carved into the work arrays (256-byte aligned)
        auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t sizes[8] = {al((size_t)nNodes), al((size_t)nNodes), al(sizeof(float) * nNodes), al(sizeof(int) * (nCells + 1)),
                                 al(sizeof(int) * (nCells + 1)), al(sizeof(int) * (3 * nNodes + 1)), al(sizeof(int) * (3 * nNodes + 1)),
                                 al(tmpBytes)};
        size_t total = 0;
        for (size_t b : sizes) total += b;
        if (total > M->scratchBytes) {
            FLIP_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(M->scratch); M->scratch = nullptr; M->scratchBytes = 0;
            FLIP_CUDA_CHECK(cudaMalloc(&M->scratch, total));
            M->scratchBytes = total;
        }
        char *q = M->scratch;
        inside = (unsigned char *)q; q += sizes[0];
        need = (unsigned char *)q; q += sizes[1];
        value = (float *)q; q += sizes[2];
        triCount = (int *)q; q += sizes[3];
        triStart = (int *)q; q += sizes[4];
        edgeFlag = (int *)q; q += sizes[5];
        edgeIdx = (int *)q; q += sizes[6];
        tmp = q;
        FLIP_CUDA_CHECK(cudaMemsetAsync(need, 0, nNodes, st));
        FLIP_CUDA_CHECK(cudaMemsetAsync(value, 0, sizeof(float) * nNodes, st));
        FLIP_CUDA_CHECK(cudaMemsetAsync(triCount + nCells, 0, sizeof(int), st));
        FLIP_CUDA_CHECK(cudaMemsetAsync(edgeFlag + 3 * nNodes, 0, sizeof(int), st));
        const ParticleSoA p = c->P[c->cur_buf];
        k_iso_inside<<<cdiv(nNodes, TPB), TPB, 0, st>>>(P, p, c->cellStart, c->occ, c->phiS, inside);
        k_iso_cells<<<cdiv(nCells, TPB), TPB, 0, st>>>(P, inside, need, triCount);
        k_iso_values<<<cdiv(nNodes, TPB), TPB, 0, st>>>(P, p, c->cellStart, c->phiS, need, value);
        k_iso_edge_count<<<cdiv(3 * nNodes, TPB), TPB, 0, st>>>(P, inside, edgeFlag);
        size_t bytes = tmpBytes;
        cub::DeviceScan::ExclusiveSum(tmp, bytes, edgeFlag, edgeIdx, (int)(3 * nNodes + 1), st);
        bytes = tmpBytes;
        cub::DeviceScan::ExclusiveSum(tmp, bytes, triCount, triStart, (int)(nCells + 1), st);
        c->launches += 6;
        if (hostValues) FLIP_CUDA_CHECK(cudaMemcpyAsync(hostValues, value, sizeof(float) * nNodes, cudaMemcpyDeviceToHost, st));
        if (hostInside) FLIP_CUDA_CHECK(cudaMemcpyAsync(hostInside, inside, nNodes, cudaMemcpyDeviceToHost, st));
        if (hostNeed) FLIP_CUDA_CHECK(cudaMemcpyAsync(hostNeed, need, nNodes, cudaMemcpyDeviceToHost, st));
        int nv = 0, nt = 0;
        FLIP_CUDA_CHECK(cudaMemcpyAsync(&nv, edgeIdx + 3 * nNodes, sizeof(int), cudaMemcpyDeviceToHost, st));
        FLIP_CUDA_CHECK(cudaMemcpyAsync(&nt, triStart + nCells, sizeof(int), cudaMemcpyDeviceToHost, st));
        FLIP_CUDA_CHECK(cudaStreamSynchronize(st));
        if ((size_t)nv > M->vertCap) {
            cudaFree(M->verts); M->verts = nullptr; M->vertCap = 0;
            FLIP_CUDA_CHECK(cudaMalloc(&M->verts, sizeof(float) * 3 * ((size_t)nv + nv / 8 + 1024)));
            M->vertCap = (size_t)nv + nv / 8 + 1024;
        }
        if ((size_t)nt > M->triCap) {
            cudaFree(M->tris); M->tris = nullptr; M->triCap = 0;
            FLIP_CUDA_CHECK(cudaMalloc(&M->tris, sizeof(int) * 3 * ((size_t)nt + nt / 8 + 1024)));
            M->triCap = (size_t)nt + nt / 8 + 1024;
        }
        if (nv > 0 && nt > 0) {
            k_iso_vertices<<<cdiv(3 * nNodes, TPB), TPB, 0, st>>>(P, c->phiS, edgeFlag, edgeIdx, value, M->verts);
            k_iso_triangles<<<cdiv(nCells, TPB), TPB, 0, st>>>(P, inside, triCount, triStart, edgeIdx, M->tris);
            c->launches += 2;
            // TriangleMesh::smooth(_surfaceReconstructionSmoothingValue = 0.5, iterations = 2)  fluidsimulation.h:1607-1608
            const size_t accBytes = (sizeof(long long) * 3 * (size_t)nv + 255) & ~(size_t)255;
            if (accBytes + sizeof(int) * (size_t)nv > M->smoothBytes) {
                cudaFree(M->smooth); M->smooth = nullptr; M->smoothBytes = 0;
                const size_t want = accBytes + sizeof(int) * ((size_t)nv + nv / 8 + 1024) + (sizeof(long long) * 3 * ((size_t)nv / 8 + 1024));
                FLIP_CUDA_CHECK(cudaMalloc(&M->smooth, want));
                M->smoothBytes = want;
            }
            long long *acc = (long long *)M->smooth;
            int *cnt = (int *)(M->smooth + accBytes);
            FLIP_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(long long) * 3 * (size_t)nv, st));
            FLIP_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)nv, st));
            for (int it = 0; it < c->surfaceSmoothingIterations; it++) {
                k_iso_smooth_accumulate<<<cdiv(nt, TPB), TPB, 0, st>>>(nt, M->tris, M->verts, acc, cnt);
                k_iso_smooth_apply<<<cdiv(nv, TPB), TPB, 0, st>>>(nv, (float)c->surfaceSmoothingValue, M->verts, acc, cnt);
                c->launches += 2;
            }
            FLIP_CUDA_CHECK(cudaStreamSynchronize(st));
        }
        M->nv = nv; M->nt = nt;
        FLIP_CUDA_CHECK(cudaGetLastError());
    } catch (...) {
        release();
        throw;
    }
    release();
}

void mesher_free(flip_ctx *c) {
    MesherState *M = (MesherState *)c->mesher;
    if (!M) return;
    cudaFree(M->verts); cudaFree(M->tris); cudaFree(M->scratch); cudaFree(M->smooth);
    delete M;
    c->mesher = nullptr;
}

// parity seam: the scalar field behind the mesh -- per node of the subdivided grid the inside flag, whether the exact
// value was needed (node of a surface cell) and that value
void mesher_debug_field(flip_ctx *c, float *values, unsigned char *inside, unsigned char *need) {
    if (!c->mesher) c->mesher = new MesherState();
    MesherState *M = (MesherState *)c->mesher;
    mesher_build(c, values, inside, need);
    M->stamp = c->stepCounter;
}

// the mesh of the particles as they stand now; cached until the next step changes them
void mesher_get(flip_ctx *c, int *nv, int *nt, float *verts, int *tris) {
    if (!c->mesher) c->mesher = new MesherState();
    MesherState *M = (MesherState *)c->mesher;
    if (M->stamp != c->stepCounter) {
        mesher_build(c);
        M->stamp = c->stepCounter;
    }
    if (nv) *nv = M->nv;
    if (nt) *nt = M->nt;
    if (verts && M->nv > 0) FLIP_CUDA_CHECK(cudaMemcpy(verts, M->verts, sizeof(float) * 3 * (size_t)M->nv, cudaMemcpyDeviceToHost));
    if (tris && M->nt > 0) FLIP_CUDA_CHECK(cudaMemcpy(tris, M->tris, sizeof(int) * 3 * (size_t)M->nt, cudaMemcpyDeviceToHost));
}

}  // namespace flip
