// Grid-side streaming kernels of the FLIP step: layered velocity extrapolation, save, body force,
// solid constraint.
#include "device_math.cuh"
#include "flip_internal.h"

namespace flip {

static constexpr int TPB = 256;

// ------------------------------------------------------------------------------------------------
// GridUtils::extrapolateGrid  gridutils.cpp:32-228, restated as a frontier sweep.
//
// level grid (1 byte per face): 0 = KNOWN from the start (valid and not on the border),
// 1..L = filled in layer l, 0xFE = border (the reference's DONE-from-the-start: never written, never
// a source of propagation, but counted as a DONE neighbour), 0xFF = UNKNOWN.
// Layer l: every face of the previous frontier (level l-1) claims its UNKNOWN 6-neighbours
// (atomicCAS, so each face is listed once); each claimed face then takes the mean of its neighbours
// that were DONE at that time (level < l, or border), summed in the order +i,-i,+j,-j,+k,-k
// (gridutils.cpp:190-224).  Claiming and filling are separate launches, as in the reference, so the
// result does not depend on thread order: bit-reproducible and equal to the reference's.
// ------------------------------------------------------------------------------------------------
__global__ void k_ext_init(const unsigned char *__restrict__ valid, unsigned char *__restrict__ level, int gi, int gj,
                           int gk, int *__restrict__ frontier, int *__restrict__ count) {
    long long n = (long long)gi * gj * gk;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    int i = (int)(t % gi);
    int j = (int)((t / gi) % gj);
    int k = (int)(t / ((long long)gi * gj));
    // Grid3d::isGridIndexOnBorder
    bool border = (i == 0 || j == 0 || k == 0 || i == gi - 1 || j == gj - 1 || k == gk - 1);
    unsigned char lv = 0xFF;
    if (border) lv = 0xFE;
    else if (valid[t]) lv = 0;
    level[t] = lv;
    if (lv == 0) {
        // only faces with at least one non-valid, non-border neighbour can claim anything
        long long sj = gi, sk = (long long)gi * gj;
        bool open = false;
        // neighbours are in range because the face is not on the border
        auto unk = [&](long long q, int qi, int qj, int qk) {
            bool b = (qi == 0 || qj == 0 || qk == 0 || qi == gi - 1 || qj == gj - 1 || qk == gk - 1);
            return !b && !valid[q];
        };
        open = unk(t + 1, i + 1, j, k) || unk(t - 1, i - 1, j, k) || unk(t + sj, i, j + 1, k) ||
               unk(t - sj, i, j - 1, k) || unk(t + sk, i, j, k + 1) || unk(t - sk, i, j, k - 1);
        if (open) {
            int slot = atomicAdd(count, 1);
            frontier[slot] = (int)t;
        }
    }
}

// _findExtrapolationCells (gridutils.cpp:123-177): frontier faces claim UNKNOWN neighbours.
__global__ void k_ext_claim(unsigned char *__restrict__ level, int gi, int gj, const int *__restrict__ frontierIn,
                            const int *__restrict__ countIn, int *__restrict__ frontierOut, int *__restrict__ countOut,
                            int layer) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int n = *countIn;
    // grid-stride: the launch is sized for the worst case known on the host
    for (; t < n; t += gridDim.x * blockDim.x) {
        int f = frontierIn[t];
        int sj = gi, sk = gi * gj;
        int nb[6] = {f + 1, f - 1, f + sj, f - sj, f + sk, f - sk};
#pragma unroll
        for (int m = 0; m < 6; m++) {
            int q = nb[m];
            // byte-wide compare-and-swap through the containing 32-bit word
            unsigned int *word = (unsigned int *)(level + (q & ~3));
            int shift = (q & 3) * 8;
            unsigned int old = *word;
            while (((old >> shift) & 0xFFu) == 0xFFu) {
                unsigned int nw = (old & ~(0xFFu << shift)) | ((unsigned int)(0x80 | layer) << shift);
                unsigned int prev = atomicCAS(word, old, nw);
                if (prev == old) {
                    int slot = atomicAdd(countOut, 1);
                    frontierOut[slot] = q;
                    break;
                }
                old = prev;
            }
        }
    }
}

// _extrapolateCellsThread (gridutils.cpp:179-228). Claimed faces carry level 0x80|layer ("WAITING")
// until this kernel ends, so they are never counted as DONE neighbours of each other.
__global__ void k_ext_fill(float *__restrict__ grid, unsigned char *__restrict__ level, int gi, int gj,
                           const int *__restrict__ frontier, const int *__restrict__ count, int layer) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int n = *count;
    for (; t < n; t += gridDim.x * blockDim.x) {
        int f = frontier[t];
        int sj = gi, sk = gi * gj;
        int nb[6] = {f + 1, f - 1, f + sj, f - sj, f + sk, f - sk};
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

// WAITING -> KNOWN for the next layer (gridutils.cpp:96-98)
__global__ void k_ext_commit(unsigned char *__restrict__ level, const int *__restrict__ frontier,
                             const int *__restrict__ count, int layer, int *__restrict__ nextCountToZero) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int n = *count;
    if (t == 0 && nextCountToZero) *nextCountToZero = 0;
    for (; t < n; t += gridDim.x * blockDim.x) level[frontier[t]] = (unsigned char)layer;
}

static void extrapolate_component(flip_ctx *c, float *grid, const unsigned char *valid, int gi, int gj, int gk) {
    cudaStream_t st = c->stream;
    long long n = (long long)gi * gj * gk;
    int *cnt0 = &c->dS->frontierCount[0];
    int *cnt1 = &c->dS->frontierCount[1];
    FLIP_CUDA_CHECK(cudaMemsetAsync(cnt0, 0, 2 * sizeof(int), st));
    k_ext_init<<<cdiv(n, TPB), TPB, 0, st>>>(valid, c->status, gi, gj, gk, c->frontier[0], cnt0);
    c->launches++;
    // frontier sizes live on the device; launches use a fixed grid with a grid-stride loop
    int blocks = 148 * 8;
    int *cnt[2] = {cnt0, cnt1};
    for (int layer = 1; layer <= c->extrapolationLayers; layer++) {
        int in = (layer - 1) & 1, out = layer & 1;
        k_ext_claim<<<blocks, TPB, 0, st>>>(c->status, gi, gj, c->frontier[in], cnt[in], c->frontier[out], cnt[out], layer);
        k_ext_fill<<<blocks, TPB, 0, st>>>(grid, c->status, gi, gj, c->frontier[out], cnt[out], layer);
        k_ext_commit<<<blocks, TPB, 0, st>>>(c->status, c->frontier[out], cnt[out], layer, cnt[in]);
        c->launches += 3;
    }
    FLIP_CUDA_CHECK(cudaGetLastError());
}

// MACVelocityField::extrapolateVelocityField  macvelocityfield.cpp:671-677
void stage_extrapolate(flip_ctx *c) {
    const Dims &d = c->d;
    size_t kt = kt_begin(c);
    extrapolate_component(c, c->U, c->validU, d.I + 1, d.J, d.K);
    extrapolate_component(c, c->V, c->validV, d.I, d.J + 1, d.K);
    extrapolate_component(c, c->W, c->validW, d.I, d.J, d.K + 1);
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
__global__ void k_add_scalar(float *__restrict__ g, int n, float a) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) g[t] = fadd(g[t], a);
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
            k_add_scalar<<<cdiv(n[a], TPB), TPB, 0, c->stream>>>(g[a], n[a], add);
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
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (w[t] == 0.0f) { a[t] = 0.0f; b[t] = 0.0f; }
}

void stage_constrain(flip_ctx *c) {
    const Dims &d = c->d;
    k_constrain<<<cdiv(d.nU, TPB), TPB, 0, c->stream>>>(c->sU, c->U, c->wU, d.nU);
    k_constrain<<<cdiv(d.nV, TPB), TPB, 0, c->stream>>>(c->sV, c->V, c->wV, d.nV);
    k_constrain<<<cdiv(d.nW, TPB), TPB, 0, c->stream>>>(c->sW, c->W, c->wW, d.nW);
    c->launches += 3;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace flip
