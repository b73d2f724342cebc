// Grid-side streaming kernels of the FLIP step: layered velocity extrapolation, save, body force,
// solid constraint.
#include <algorithm>
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
// Layer l: every face of the previous frontier (level l-1) claims its UNKNOWN 6-neighbours by
// setting their level to l (byte-wide atomicCAS, so each face is listed once); each claimed face then
// takes the mean of its neighbours that were DONE before this layer (level < l, or border) -- faces
// claimed in the same layer carry level l and are therefore not counted, which is the reference's
// WAITING state -- summed in the order +i,-i,+j,-j,+k,-k (gridutils.cpp:190-224).  Claiming and filling
// are separate launches, as in the reference, so the result does not depend on thread order:
// bit-reproducible and equal to the reference's.
// ------------------------------------------------------------------------------------------------
struct ExtComp {
    float *grid;
    const unsigned char *valid;
    unsigned char *level;
    int *frontier[2];
    int *count;          // [2]
    int gi, gj, gk;
};
struct ExtArgs {
    ExtComp c[3];
};

// four consecutive faces per thread (the arrays are padded by 64 entries, so whole-word accesses stay in bounds)
__global__ void k_ext_init(ExtArgs A) {
    const ExtComp &C = A.c[blockIdx.y];
    const int gi = C.gi, gj = C.gj, gk = C.gk;
    const long long n = (long long)gi * gj * gk;
    const long long t0 = 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
    if (t0 >= n) return;
    const unsigned int v4 = *reinterpret_cast<const unsigned int *>(C.valid + t0);
    int i = (int)(t0 % gi);
    int j = (int)((t0 / gi) % gj);
    int k = (int)(t0 / ((long long)gi * gj));
    const long long sj = gi, sk = (long long)gi * gj;
    unsigned int l4 = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const long long t = t0 + m;
        unsigned int lv = 0xFFu;
        if (t < n) {
            // Grid3d::isGridIndexOnBorder
            const bool border = (i == 0 || j == 0 || k == 0 || i == gi - 1 || j == gj - 1 || k == gk - 1);
            const bool val = ((v4 >> (8 * m)) & 0xFFu) != 0;
            if (border) lv = 0xFEu;
            else if (val) {
                lv = 0u;
                // only faces with at least one non-valid, non-border neighbour can claim anything
                // (neighbours are in range because the face is not on the border)
                auto unk = [&](long long q, int qi, int qj, int qk) {
                    bool b = (qi == 0 || qj == 0 || qk == 0 || qi == gi - 1 || qj == gj - 1 || qk == gk - 1);
                    return !b && !C.valid[q];
                };
                const bool open = unk(t + 1, i + 1, j, k) || unk(t - 1, i - 1, j, k) || unk(t + sj, i, j + 1, k) ||
                                  unk(t - sj, i, j - 1, k) || unk(t + sk, i, j, k + 1) || unk(t - sk, i, j, k - 1);
                if (open) C.frontier[0][warp_append_slot(&C.count[0])] = (int)t;
            }
        }
        l4 |= lv << (8 * m);
        if (++i == gi) { i = 0; if (++j == gj) { j = 0; k++; } }
    }
    *reinterpret_cast<unsigned int *>(C.level + t0) = l4;
}

// _findExtrapolationCells (gridutils.cpp:123-177): frontier faces claim UNKNOWN neighbours.
__global__ void k_ext_claim(ExtArgs A, int layer) {
    const ExtComp &C = A.c[blockIdx.y];
    const int in = (layer - 1) & 1, out = layer & 1;
    const int *__restrict__ frontierIn = C.frontier[in];
    int *__restrict__ frontierOut = C.frontier[out];
    unsigned char *level = C.level;
    const int n = C.count[in];
    const int sj = C.gi, sk = C.gi * C.gj;
    // grid-stride: the launch is sized for the worst case known on the host
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int f = frontierIn[t];
        const int nb[6] = {f + 1, f - 1, f + sj, f - sj, f + sk, f - sk};
#pragma unroll
        for (int m = 0; m < 6; m++) {
            const int q = nb[m];
            if (level[q] != 0xFFu) continue;
            // byte-wide compare-and-swap through the containing 32-bit word
            unsigned int *word = (unsigned int *)(level + (q & ~3));
            const int shift = (q & 3) * 8;
            unsigned int old = *word;
            while (((old >> shift) & 0xFFu) == 0xFFu) {
                unsigned int nw = (old & ~(0xFFu << shift)) | ((unsigned int)layer << shift);
                unsigned int prev = atomicCAS(word, old, nw);
                if (prev == old) {
                    frontierOut[warp_append_slot(&C.count[out])] = q;
                    break;
                }
                old = prev;
            }
        }
    }
}

// _extrapolateCellsThread (gridutils.cpp:179-228).  Also clears the counter the next layer appends to.
__global__ void k_ext_fill(ExtArgs A, int layer) {
    const ExtComp &C = A.c[blockIdx.y];
    const int in = (layer - 1) & 1, out = layer & 1;
    const int *__restrict__ frontier = C.frontier[out];
    const unsigned char *__restrict__ level = C.level;
    float *grid = C.grid;
    const int n = C.count[out];
    const int sj = C.gi, sk = C.gi * C.gj;
    if (blockIdx.x == 0 && threadIdx.x == 0) C.count[in] = 0;    // consumed by this layer's claim launch
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int f = frontier[t];
        const int nb[6] = {f + 1, f - 1, f + sj, f - sj, f + sk, f - sk};
        float sum = 0.0f;
        int cnt = 0;
#pragma unroll
        for (int m = 0; m < 6; m++) {
            unsigned char lv = level[nb[m]];
            bool done = (lv == 0xFE) || (lv < (unsigned char)layer);
            if (done) { sum = fadd(sum, grid[nb[m]]); cnt++; }
        }
        grid[f] = __fdiv_rn(sum, (float)cnt);
    }
}

// MACVelocityField::extrapolateVelocityField  macvelocityfield.cpp:671-677
void stage_extrapolate(flip_ctx *c) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    size_t kt = kt_begin(c);
    const size_t nmax = ext_stride(d);   // per-component stride of the scratch arrays
    ExtArgs A;
    float *grids[3] = {c->U, c->V, c->W};
    const unsigned char *valids[3] = {c->validU, c->validV, c->validW};
    const int gi[3] = {d.I + 1, d.I, d.I}, gj[3] = {d.J, d.J + 1, d.J}, gk[3] = {d.K, d.K, d.K + 1};
    long long nbig = 0;
    for (int m = 0; m < 3; m++) {
        A.c[m].grid = grids[m]; A.c[m].valid = valids[m];
        A.c[m].level = c->status + m * nmax;
        A.c[m].frontier[0] = c->frontier[0] + m * nmax;
        A.c[m].frontier[1] = c->frontier[1] + m * nmax;
        A.c[m].count = &c->dS->extCount[2 * m];
        A.c[m].gi = gi[m]; A.c[m].gj = gj[m]; A.c[m].gk = gk[m];
        nbig = std::max(nbig, (long long)gi[m] * gj[m] * gk[m]);
    }
    FLIP_CUDA_CHECK(cudaMemsetAsync(c->dS->extCount, 0, 6 * sizeof(int), st));
    k_ext_init<<<dim3(cdiv(cdiv(nbig, 4), TPB), 3), TPB, 0, st>>>(A);
    c->launches++;
    // frontier sizes live on the device; launches use a fixed grid with a grid-stride loop
    const dim3 blocks(148 * 3, 3);
    for (int layer = 1; layer <= c->extrapolationLayers; layer++) {
        k_ext_claim<<<blocks, TPB, 0, st>>>(A, layer);
        k_ext_fill<<<blocks, TPB, 0, st>>>(A, layer);
        c->launches += 2;
    }
    FLIP_CUDA_CHECK(cudaGetLastError());
    kt_end(c, FLIP_KERNEL_EXTRAPOLATE, kt);
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

void stage_constrain(flip_ctx *c) {
    const Dims &d = c->d;
    k_constrain<<<cdiv(cdiv(d.nU, 4), TPB), TPB, 0, c->stream>>>(c->sU, c->U, c->wU, d.nU);
    k_constrain<<<cdiv(cdiv(d.nV, 4), TPB), TPB, 0, c->stream>>>(c->sV, c->V, c->wV, d.nV);
    k_constrain<<<cdiv(cdiv(d.nW, 4), TPB), TPB, 0, c->stream>>>(c->sW, c->W, c->wW, d.nW);
    c->launches += 3;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flip
