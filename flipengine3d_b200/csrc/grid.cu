// Grid-side streaming kernels of the FLIP step: layered velocity extrapolation, save, body force,
// solid constraint.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "device_math.cuh"
#include "flip_internal.h"

namespace flip {

static constexpr int TPB = 256;

// ------------------------------------------------------------------------------------------------
// GridUtils::extrapolateGrid  gridutils.cpp:32-228, restated as a frontier sweep over all three MAC
// components at once (blockIdx.y = component).
//
// level grid (1 byte per face): 0 = KNOWN from the start (valid and not on the border),
// 1..L = filled in layer l, 0xFE = border (the reference's DONE-from-the-start: never written, never
// a source of propagation, but counted as a DONE neighbour), 0xFF = UNKNOWN.
// Layer l: every face of the previous frontier (level l-1) looks at its UNKNOWN 6-neighbours; each such face is taken
// by exactly one of the frontier faces around it (the first in its own summation order, see k_ext_layer) and set to
// the mean of its neighbours that were DONE before this layer (level < l, or border) -- faces taken in the same
// layer carry level l or still 0xFF and are therefore not counted, which is the reference's WAITING state -- summed
// in the order +i,-i,+j,-j,+k,-k (gridutils.cpp:190-224).  The result does not depend on thread order:
// bit-reproducible and equal to the reference's.
// ------------------------------------------------------------------------------------------------
static constexpr int EXT_MAX_LAYERS = 31;      // counters per component: the start frontier and one per layer
struct ExtComp {
    float *grid;
    const unsigned char *valid;
    unsigned char *level;
    int *frontier[2];
    int *count;          // [EXT_MAX_LAYERS + 1]: size of the start frontier, then of the faces filled in layer l
    int gi, gj, gk;
};
struct ExtArgs {
    ExtComp c[3];
};

// sixteen byte lanes as two 64-bit words (byte b of the run = byte b & 7 of word b >> 3)
struct Bytes16 {
    unsigned long long lo, hi;
};
__device__ __forceinline__ Bytes16 b16_and(Bytes16 a, Bytes16 b) { return {a.lo & b.lo, a.hi & b.hi}; }
__device__ __forceinline__ Bytes16 b16_or(Bytes16 a, Bytes16 b) { return {a.lo | b.lo, a.hi | b.hi}; }
__device__ __forceinline__ Bytes16 b16_not(Bytes16 a) { return {~a.lo, ~a.hi}; }
// 0xFF in the bytes [0, n), n in [0, 16]
__device__ __forceinline__ Bytes16 b16_below(int n) {
    const unsigned long long lo = n >= 8 ? ~0ull : ((1ull << (8 * n)) - 1ull);
    const unsigned long long hi = n <= 8 ? 0ull : (n >= 16 ? ~0ull : ((1ull << (8 * (n - 8))) - 1ull));
    return {lo, hi};
}
// 0xFF in byte p (nothing when p is outside [0, 16))
__device__ __forceinline__ Bytes16 b16_at(int p) {
    return {(p >= 0 && p < 8) ? 0xFFull << (8 * p) : 0ull, (p >= 8 && p < 16) ? 0xFFull << (8 * (p - 8)) : 0ull};
}
__device__ __forceinline__ Bytes16 b16_if(bool c, Bytes16 a) { return c ? a : Bytes16{0ull, 0ull}; }
// 0xFF per non-zero byte
__device__ __forceinline__ Bytes16 b16_nonzero(uint4 v) {
    const unsigned int x = __vcmpne4(v.x, 0u), y = __vcmpne4(v.y, 0u), z = __vcmpne4(v.z, 0u), w = __vcmpne4(v.w, 0u);
    return {(unsigned long long)x | ((unsigned long long)y << 32), (unsigned long long)z | ((unsigned long long)w << 32)};
}

// sixteen bytes starting at an arbitrary address, as four words (five aligned loads and a funnel shift)
__device__ __forceinline__ uint4 load16_unaligned(const unsigned char *p) {
    const unsigned int *w = reinterpret_cast<const unsigned int *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const unsigned int sh = ((unsigned int)(reinterpret_cast<uintptr_t>(p) & 3)) * 8;
    const unsigned int a0 = __ldg(w), a1 = __ldg(w + 1), a2 = __ldg(w + 2), a3 = __ldg(w + 3), a4 = __ldg(w + 4);
    return make_uint4(__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh), __funnelshift_r(a3, a4, sh));
}

// Start pass, sixteen consecutive faces per thread (linear index; the arrays are padded by 64 entries, so whole-vector
// accesses stay in bounds), all of it byte-mask arithmetic so that no lane of a warp leaves the common path: a run may
// cross ONE row end (rows are at least 17 faces long -- smaller grids take k_ext_init_small), which splits it into two
// segments with their own (j, k); border faces, valid faces and the faces next to the border follow from the segment
// and from the byte position.
//   border   B: Grid3d::isGridIndexOnBorder
//   valid    V
//   a neighbour is "known-like" when it is valid or on the border (the claim test of the layer pass never takes it);
//   a valid, non-border face joins the start frontier when one of its six neighbours is not known-like.
__global__ void __launch_bounds__(TPB) k_ext_init(ExtArgs A) {
    const ExtComp &C = A.c[blockIdx.y];
    const int gi = C.gi, gj = C.gj, gk = C.gk;
    // (face counts fit 31 bits: the frontier lists hold int indices)
    const unsigned int n = (unsigned int)gi * gj * gk;
    const unsigned int t0 = 16u * (blockIdx.x * blockDim.x + threadIdx.x);
    Bytes16 open = {0ull, 0ull};
    if (t0 < n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(C.valid + t0));
        const unsigned int row = t0 / (unsigned int)gi;
        const int i0 = (int)(t0 - row * (unsigned int)gi);
        const int k0 = (int)(row / (unsigned int)gj);
        const int j0 = (int)(row - (unsigned int)k0 * (unsigned int)gj);
        const int sj = gi, sk = gi * gj;
        if (!(v.x | v.y | v.z | v.w) && i0 >= 1 && i0 + 16 < gi && j0 >= 1 && j0 + 1 < gj && k0 >= 1 && k0 + 1 < gk) {
            // nothing valid in a run away from every border: all unknown
            *reinterpret_cast<uint4 *>(C.level + t0) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        } else {
            // segment 1: bytes [0, n1) in row (j0, k0); segment 2: the rest, in the next row
            const int n1 = min(16, gi - i0);
            int j1 = j0 + 1, k1 = k0;
            if (j1 == gj) { j1 = 0; k1 = k0 + 1; }
            const Bytes16 seg1 = b16_below(n1), seg2 = b16_not(seg1);
            auto rows = [&](bool c0, bool c1) { return b16_or(b16_if(c0, seg1), b16_if(c1, seg2)); };
            const Bytes16 B = b16_or(rows(j0 == 0 || j0 == gj - 1 || k0 == 0 || k0 == gk - 1, j1 == 0 || j1 == gj - 1 || k1 == 0 || k1 >= gk - 1),
                                     b16_or(b16_or(b16_at(-i0), b16_at(n1 < 16 ? n1 : -1)), b16_at(gi - 1 - i0)));      // i == 0 | i == gi-1
            const Bytes16 V = b16_nonzero(v);
            const Bytes16 cand = b16_and(V, b16_not(B));
            if (cand.lo | cand.hi) {      // (every face of the run is on the border when a neighbour row would be out of range)
                // known-like neighbours
                const Bytes16 Vjp = b16_nonzero(load16_unaligned(C.valid + t0 + sj)), Vjm = b16_nonzero(load16_unaligned(C.valid + t0 - sj));
                const Bytes16 Vkp = b16_nonzero(load16_unaligned(C.valid + t0 + sk)), Vkm = b16_nonzero(load16_unaligned(C.valid + t0 - sk));
                const unsigned long long left = C.valid[t0 - 1] ? 0xFFull : 0ull, right = C.valid[t0 + 16] ? 0xFFull << 56 : 0ull;
                const Bytes16 Vim = {(V.lo << 8) | left, (V.hi << 8) | (V.lo >> 56)};       // byte b: validity of face b - 1
                const Bytes16 Vip = {(V.lo >> 8) | (V.hi << 56), (V.hi >> 8) | right};      // byte b: validity of face b + 1
                Bytes16 all = b16_or(Vip, b16_at(gi - 2 - i0));                                                   // i + 1 == gi - 1
                all = b16_and(all, b16_or(Vim, b16_or(b16_at(1 - i0), b16_at(n1 < 16 ? n1 + 1 : -1))));          // i - 1 == 0
                all = b16_and(all, b16_or(Vjp, rows(j0 + 1 == gj - 1, j1 + 1 == gj - 1)));
                all = b16_and(all, b16_or(Vjm, rows(j0 - 1 == 0, j1 - 1 == 0)));
                all = b16_and(all, b16_or(Vkp, rows(k0 + 1 == gk - 1, k1 + 1 == gk - 1)));
                all = b16_and(all, b16_or(Vkm, rows(k0 - 1 == 0, k1 - 1 == 0)));
                open = b16_and(cand, b16_not(all));
            }
            // level: 0xFE on the border, 0 where valid, 0xFF (unknown) elsewhere
            const Bytes16 lv = b16_or(b16_and(B, Bytes16{0xFEFEFEFEFEFEFEFEull, 0xFEFEFEFEFEFEFEFEull}), b16_and(b16_not(B), b16_not(V)));
            *reinterpret_cast<uint4 *>(C.level + t0) =
                make_uint4((unsigned int)lv.lo, (unsigned int)(lv.lo >> 32), (unsigned int)lv.hi, (unsigned int)(lv.hi >> 32));
        }
    }
    // the start frontier: the faces next to the free surface, appended with one atomic per block
    const int mine = (__popcll(open.lo) + __popcll(open.hi)) >> 3;
    int slot = block_append_slots<TPB>(&C.count[0], mine);
    const unsigned long long ow[2] = {open.lo, open.hi};
#pragma unroll
    for (int w = 0; w < 2; w++) {
        unsigned long long o = ow[w];
        while (o) {
            const int byte = (__ffsll((long long)o) - 1) >> 3;
            o &= ~(0xFFull << (8 * byte));
            C.frontier[0][slot++] = (int)(t0 + 8 * w + byte);
        }
    }
}

// the same for grids with rows shorter than 17 faces: one face per thread
__global__ void k_ext_init_small(ExtArgs A) {
    const ExtComp &C = A.c[blockIdx.y];
    const int gi = C.gi, gj = C.gj, gk = C.gk;
    const long long n = (long long)gi * gj * gk;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % gi), j = (int)((t / gi) % gj), k = (int)(t / ((long long)gi * gj));
    unsigned char lv = 0xFFu;
    if (i == 0 || j == 0 || k == 0 || i == gi - 1 || j == gj - 1 || k == gk - 1) lv = 0xFEu;
    else if (C.valid[t]) {
        lv = 0u;
        const long long sj = gi, sk = (long long)gi * gj;
        auto unk = [&](long long q, int qi, int qj, int qk) {
            const bool b = (qi == 0 || qj == 0 || qk == 0 || qi == gi - 1 || qj == gj - 1 || qk == gk - 1);
            return !b && !C.valid[q];
        };
        const bool open = unk(t + 1, i + 1, j, k) || unk(t - 1, i - 1, j, k) || unk(t + sj, i, j + 1, k) ||
                          unk(t - sj, i, j - 1, k) || unk(t + sk, i, j, k + 1) || unk(t - sk, i, j, k - 1);
        if (open) C.frontier[0][warp_append_slot(&C.count[0])] = (int)t;
    }
    C.level[t] = lv;
}

// One layer: _findExtrapolationCells (gridutils.cpp:123-177) and _extrapolateCellsThread (:179-228) in one pass.
// Every face of the previous frontier looks at its UNKNOWN neighbours q; q is taken by exactly one of them -- the first
// of q's own neighbours, in the order +i,-i,+j,-j,+k,-k, that belongs to the previous frontier (all faces of level
// layer-1 next to an unknown face are in the list) -- so no atomic is needed to list q once.  The owner fills q with the
// mean of q's neighbours that were DONE before this layer and marks it.  Faces taken concurrently show either 0xFF or
// `layer` to other threads; neither counts as DONE nor as a previous-frontier face, so the outcome does not depend on
// the interleaving, and the values those means read were written by earlier launches.
__global__ void __launch_bounds__(TPB) k_ext_layer(ExtArgs A, int layer) {
    const ExtComp &C = A.c[blockIdx.y];
    const int *__restrict__ frontierIn = C.frontier[(layer - 1) & 1];
    int *__restrict__ frontierOut = C.frontier[layer & 1];
    unsigned char *level = C.level;     // (a stale 0xFF for a face taken meanwhile reads like an earlier look: see above)
    float *grid = C.grid;
    const int n = C.count[layer - 1];
    const int sj = C.gi, sk = C.gi * C.gj;
    const unsigned int prev = (unsigned int)(layer - 1);
    // one thread per (frontier face, direction): six short dependent chains side by side instead of one long one;
    // block-uniform grid-stride loop (the frontier size lives on the device), one counter atomic per block and trip
    const int offs[6] = {1, -1, sj, -sj, sk, -sk};
    for (int w0 = blockIdx.x * blockDim.x; w0 < 6 * n; w0 += gridDim.x * blockDim.x) {
        const int w = w0 + threadIdx.x;
        int q = -1;
        if (w < 6 * n) {
            const int t = w / 6, m = w - 6 * t;
            const int cand = frontierIn[t] + offs[m];
            if (level[cand] == 0xFFu) {
                const int qn[6] = {cand + 1, cand - 1, cand + sj, cand - sj, cand + sk, cand - sk};
                unsigned int lq[6];
#pragma unroll
                for (int r = 0; r < 6; r++) lq[r] = level[qn[r]];
                int owner = 6;
#pragma unroll
                for (int r = 5; r >= 0; r--) if (lq[r] == prev) owner = r;
                if (owner == (m ^ 1)) {                     // the frontier face is the neighbour in the opposite direction
                    float sum = 0.0f;
                    int cnt = 0;
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        const bool done = (lq[r] == 0xFEu) || (lq[r] < (unsigned int)layer);
                        if (done) { sum = fadd(sum, grid[qn[r]]); cnt++; }
                    }
                    q = cand;
                    grid[q] = __fdiv_rn(sum, (float)cnt);
                    level[q] = (unsigned char)layer;
                }
            }
        }
        const int slot = block_append_slots<TPB>(&C.count[layer], q >= 0 ? 1 : 0);
        if (q >= 0) frontierOut[slot] = q;
    }
}

// MACVelocityField::extrapolateVelocityField  macvelocityfield.cpp:671-677
static void extrapolate_fields(flip_ctx *c, float *const grids[3], const unsigned char *const valids[3], int layers, bool timed);
void stage_extrapolate(flip_ctx *c) {
    float *grids[3] = {c->U, c->V, c->W};
    const unsigned char *valids[3] = {c->validU, c->validV, c->validW};
    extrapolate_fields(c, grids, valids, c->extrapolationLayers, true);
}

// The velocity data of the solid SDF after all solids were merged (MeshLevelSet::normalizeVelocityGrid, meshlevelset.cpp:
// 640-702, _normalizeVelocityGridThread :1738-1756): value / weight where the summed weight exceeds 1e-6 (those faces are
// valid), 0 elsewhere, then extrapolated over `layers` (5, meshlevelset.h:364) layers with the same routine as the fluid.
__global__ void k_solid_normalize(float *__restrict__ field, const float *__restrict__ weight, unsigned char *__restrict__ valid, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float w = weight[t];
    const bool ok = w > 1e-6f;
    field[t] = ok ? __fdiv_rn(field[t], w) : 0.0f;
    valid[t] = ok ? 1 : 0;
}
void solid_velocity_normalize_extrapolate(flip_ctx *c, int layers) {
    const Dims &d = c->d;
    float *grids[3] = {c->solU, c->solV, c->solW};
    const int n[3] = {d.nU, d.nV, d.nW};
    for (int m = 0; m < 3; m++) {
        k_solid_normalize<<<cdiv(n[m], TPB), TPB, 0, c->stream>>>(grids[m], c->solidWeightSum[m], c->solidValid[m], n[m]);
        c->launches++;
    }
    const unsigned char *valids[3] = {c->solidValid[0], c->solidValid[1], c->solidValid[2]};
    extrapolate_fields(c, grids, valids, layers, false);
}

static void extrapolate_fields(flip_ctx *c, float *const grids[3], const unsigned char *const valids[3], int layers, bool timed) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    size_t kt = timed ? kt_begin(c) : 0;
    const size_t nmax = ext_stride(d);   // per-component stride of the scratch arrays
    ExtArgs A;
    const int gi[3] = {d.I + 1, d.I, d.I}, gj[3] = {d.J, d.J + 1, d.J}, gk[3] = {d.K, d.K, d.K + 1};
    long long nbig = 0;
    for (int m = 0; m < 3; m++) {
        A.c[m].grid = grids[m]; A.c[m].valid = valids[m];
        A.c[m].level = c->status + m * nmax;
        A.c[m].frontier[0] = c->frontier[0] + m * nmax;
        A.c[m].frontier[1] = c->frontier[1] + m * nmax;
        A.c[m].count = &c->dS->extCount[(EXT_MAX_LAYERS + 1) * m];
        A.c[m].gi = gi[m]; A.c[m].gj = gj[m]; A.c[m].gk = gk[m];
        nbig = std::max(nbig, (long long)gi[m] * gj[m] * gk[m]);
    }
    static const bool trace = getenv("FLIP_EXT_TRACE") != nullptr;      // developer knob: frontier sizes of the previous call
    if (trace) {
        for (int m = 0; m < 3; m++) {
            fprintf(stderr, "ext comp %d:", m);
            for (int l = 0; l <= layers; l++) fprintf(stderr, " %d", c->hS->extCount[(EXT_MAX_LAYERS + 1) * m + l]);
            fprintf(stderr, "\n");
        }
    }
    if (layers > EXT_MAX_LAYERS) throw ApiError(FLIP_ERR_UNSUPPORTED, "more than 31 extrapolation layers (CFL > 29)");
    FLIP_CUDA_CHECK(cudaMemsetAsync(c->dS->extCount, 0, sizeof(c->dS->extCount), st));
    if (d.I >= 17) k_ext_init<<<dim3(cdiv(cdiv(nbig, 16), TPB), 3), TPB, 0, st>>>(A);
    else k_ext_init_small<<<dim3(cdiv(nbig, TPB), 3), TPB, 0, st>>>(A);
    c->launches++;
    // frontier sizes live on the device; launches use a fixed grid with a grid-stride loop
    const dim3 blocks(148 * 8, 3);
    for (int layer = 1; layer <= layers; layer++) {
        k_ext_layer<<<blocks, TPB, 0, st>>>(A, layer);
        c->launches++;
    }
    FLIP_CUDA_CHECK(cudaGetLastError());
    if (timed) kt_end(c, FLIP_KERNEL_EXTRAPOLATE, kt);
}

// ------------------------------------------------------------------------------------------------
// _saveVelocityField (fluidsimulation.cpp:3287) and _applyConstantBodyForces (:3450-3489)
// ------------------------------------------------------------------------------------------------
void stage_save(flip_ctx *c) {
    const Dims &d = c->d;
    FLIP_CUDA_CHECK(cudaMemcpyAsync(c->sU, c->U, sizeof(float) * d.nU, cudaMemcpyDeviceToDevice, c->stream));
    FLIP_CUDA_CHECK(cudaMemcpyAsync(c->sV, c->V, sizeof(float) * d.nV, cudaMemcpyDeviceToDevice, c->stream));
    FLIP_CUDA_CHECK(cudaMemcpyAsync(c->sW, c->W, sizeof(float) * d.nW, cudaMemcpyDeviceToDevice, c->stream));
}

// MACVelocityField::addU(i,j,k, bodyForce.x * dt): the addend is float*double -> double, narrowed to
// float by the `double num` -> `_u.add(i,j,k,num)` call (macvelocityfield.cpp:255-261), then a float add.
// four faces per thread (the grids are 256-byte aligned)
__global__ void k_add_scalar(float *__restrict__ g, int n, float a) {
    int t4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * t4 >= n) return;
    if (4 * t4 + 3 < n) {
        float4 v = reinterpret_cast<float4 *>(g)[t4];
        v.x = fadd(v.x, a); v.y = fadd(v.y, a); v.z = fadd(v.z, a); v.w = fadd(v.w, a);
        reinterpret_cast<float4 *>(g)[t4] = v;
    } else {
        for (int t = 4 * t4; t < n; t++) g[t] = fadd(g[t], a);
    }
}

void stage_body_force(flip_ctx *c, double dt) {
    const Dims &d = c->d;
    const float eps = 1e-6f;
    float bf[3] = {(float)c->gravity[0], (float)c->gravity[1], (float)c->gravity[2]};   // vmath::vec3 components
    float *g[3] = {c->U, c->V, c->W};
    int n[3] = {d.nU, d.nV, d.nW};
    for (int a = 0; a < 3; a++) {
        if (fabs(bf[a]) > eps) {
            float add = (float)((double)bf[a] * dt);
            k_add_scalar<<<cdiv(cdiv(n[a], 4), TPB), TPB, 0, c->stream>>>(g[a], n[a], add);
            c->launches++;
        }
    }
    FLIP_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// _constrainVelocityFields (fluidsimulation.cpp:3884-3946) for static solids with zero friction:
// weight == 0 -> solid face velocity (0); 0 < weight < 1 -> f*uface + (1-f)*umac with f = 0, which is
// umac itself (0*0 + 1*umac), so only the first rule writes.  Applied to the saved field and to the
// new field.
// ------------------------------------------------------------------------------------------------
__global__ void k_constrain(float *__restrict__ a, float *__restrict__ b, const float *__restrict__ w, int n) {
    int t4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * t4 >= n) return;
    const float4 w4 = __ldg(reinterpret_cast<const float4 *>(w) + t4);      // the weights: one 16-byte load per four faces
    const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int t = 4 * t4 + m;
        if (t < n && ws[m] == 0.0f) { a[t] = 0.0f; b[t] = 0.0f; }
    }
}

// ... with moving solids: such a face takes the solid's face velocity (fluidsimulation.cpp:3893-3894); the friction of
// partly open faces is 0 as before.
__global__ void k_constrain_moving(float *__restrict__ a, float *__restrict__ b, const float *__restrict__ w,
                                   const float *__restrict__ solid, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && w[t] == 0.0f) { const float v = solid[t]; a[t] = v; b[t] = v; }
}

// ... with friction: a partly open face is pulled towards the solid's velocity, f u_solid + (1 - f) u (:3895-3900), f the
// face friction of build_face_friction; solid: null while every solid is at rest.
__global__ void k_constrain_friction(float *__restrict__ a, float *__restrict__ b, const float *__restrict__ w,
                                     const float *__restrict__ solid, const float *__restrict__ friction, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float wt = w[t];
    const float us = solid ? solid[t] : 0.0f;
    if (wt == 0.0f) { a[t] = us; b[t] = us; }
    else if (wt < 1.0f) {
        const float f = friction[t], g = fsub(1.0f, f);
        a[t] = fadd(fmul(f, us), fmul(g, a[t]));
        b[t] = fadd(fmul(f, us), fmul(g, b[t]));
    }
}

void stage_constrain(flip_ctx *c) {
    const Dims &d = c->d;
    if (c->fricU) {
        k_constrain_friction<<<cdiv(d.nU, TPB), TPB, 0, c->stream>>>(c->sU, c->U, c->wU, c->solU, c->fricU, d.nU);
        k_constrain_friction<<<cdiv(d.nV, TPB), TPB, 0, c->stream>>>(c->sV, c->V, c->wV, c->solV, c->fricV, d.nV);
        k_constrain_friction<<<cdiv(d.nW, TPB), TPB, 0, c->stream>>>(c->sW, c->W, c->wW, c->solW, c->fricW, d.nW);
        c->launches += 3;
        FLIP_CUDA_CHECK(cudaGetLastError());
        return;
    }
    if (c->solU) {
        k_constrain_moving<<<cdiv(d.nU, TPB), TPB, 0, c->stream>>>(c->sU, c->U, c->wU, c->solU, d.nU);
        k_constrain_moving<<<cdiv(d.nV, TPB), TPB, 0, c->stream>>>(c->sV, c->V, c->wV, c->solV, d.nV);
        k_constrain_moving<<<cdiv(d.nW, TPB), TPB, 0, c->stream>>>(c->sW, c->W, c->wW, c->solW, d.nW);
        c->launches += 3;
        FLIP_CUDA_CHECK(cudaGetLastError());
        return;
    }
    k_constrain<<<cdiv(cdiv(d.nU, 4), TPB), TPB, 0, c->stream>>>(c->sU, c->U, c->wU, d.nU);
    k_constrain<<<cdiv(cdiv(d.nV, 4), TPB), TPB, 0, c->stream>>>(c->sV, c->V, c->wV, d.nV);
    k_constrain<<<cdiv(cdiv(d.nW, 4), TPB), TPB, 0, c->stream>>>(c->sW, c->W, c->wW, d.nW);
    c->launches += 3;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flip
