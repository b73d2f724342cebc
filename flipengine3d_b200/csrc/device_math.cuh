// Device helpers that restate the reference's scalar arithmetic with explicit round-to-nearest
// intrinsics, so that nvcc cannot contract mul+add into FMA: the parity oracle is built with
// -ffp-contract=off.
#pragma once
#include <cuda_runtime.h>

namespace flip {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// vmath::dot / lengthsq: x*x + y*y + z*z, left to right in float (vmath.h:76-88)
__device__ __forceinline__ float lengthsq3(float x, float y, float z) {
    return fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z));
}
// vmath::length (vmath.h:90): sqrt of the float sum; (float)sqrt((double)s) == sqrtf(s) (53 >= 2*24+2)
__device__ __forceinline__ float length3(float x, float y, float z) { return __fsqrt_rn(lengthsq3(x, y, z)); }

// Grid3d::positionToGridIndex (grid3d.h:56-61): (int)floor(p * (1.0/dx)) in double
__device__ __forceinline__ int pos2idx(float p, double invdx) { return (int)floor(dmul((double)p, invdx)); }
__device__ __forceinline__ int pos2idx_d(double p, double invdx) { return (int)floor(dmul(p, invdx)); }

// Interpolation::trilinearInterpolate(double p[8], x, y, z)  interpolation.cpp:61-70
// vertex order {000,100,010,001,101,011,110,111}; products and sums left to right, no FMA.
__device__ __forceinline__ double trilerp8(const double p[8], double x, double y, double z) {
    double mx = dsub(1.0, x), my = dsub(1.0, y), mz = dsub(1.0, z);
    double s = dmul(dmul(dmul(p[0], mx), my), mz);
    s = dadd(s, dmul(dmul(dmul(p[1], x), my), mz));
    s = dadd(s, dmul(dmul(dmul(p[2], mx), y), mz));
    s = dadd(s, dmul(dmul(dmul(p[3], mx), my), z));
    s = dadd(s, dmul(dmul(dmul(p[4], x), my), z));
    s = dadd(s, dmul(dmul(dmul(p[5], mx), y), z));
    s = dadd(s, dmul(dmul(dmul(p[6], x), y), mz));
    s = dadd(s, dmul(dmul(dmul(p[7], x), y), z));
    return s;
}

// One MAC component sampled as MACVelocityField::_interpolateLinearU/V/W do
// (macvelocityfield.cpp:519-613): shift the two transverse axes by dx/2 in double, floor, 8
// samples with out-of-range -> 0, weights in double.  (sx,sy,sz) = 1 where the half-cell shift applies.
// gi,gj,gk = dimensions of this component's array.
template <int SX, int SY, int SZ>
__device__ __forceinline__ double sample_mac(const float *__restrict__ g, int gi, int gj, int gk, double x, double y,
                                             double z, double dx, double invdx, double hdx, int kOff) {
    if (SX) x = dsub(x, hdx);
    if (SY) y = dsub(y, hdx);
    if (SZ) z = dsub(z, hdx);
    int i = pos2idx_d(x, invdx), j = pos2idx_d(y, invdx), k = pos2idx_d(z, invdx);
    double fx = dmul(dsub(x, dmul((double)i, dx)), invdx);
    double fy = dmul(dsub(y, dmul((double)j, dx)), invdx);
    double fz = dmul(dsub(z, dmul((double)k, dx)), invdx);
    double p[8];
    bool i0 = (i >= 0 && i < gi), i1 = (i + 1 >= 0 && i + 1 < gi);
    bool j0 = (j >= 0 && j < gj), j1 = (j + 1 >= 0 && j + 1 < gj);
    // arrays hold the local planes of a z-slab: plane k (global) is stored at k - kOff
    int kl = k - kOff;
    bool k0 = (kl >= 0 && kl < gk), k1 = (kl + 1 >= 0 && kl + 1 < gk);
    long long base = (long long)i + (long long)gi * ((long long)j + (long long)gj * (long long)kl);
    long long sj = gi, sk = (long long)gi * gj;
    p[0] = (i0 && j0 && k0) ? (double)__ldg(g + base) : 0.0;
    p[1] = (i1 && j0 && k0) ? (double)__ldg(g + base + 1) : 0.0;
    p[2] = (i0 && j1 && k0) ? (double)__ldg(g + base + sj) : 0.0;
    p[3] = (i0 && j0 && k1) ? (double)__ldg(g + base + sk) : 0.0;
    p[4] = (i1 && j0 && k1) ? (double)__ldg(g + base + sk + 1) : 0.0;
    p[5] = (i0 && j1 && k1) ? (double)__ldg(g + base + sk + sj) : 0.0;
    p[6] = (i1 && j1 && k0) ? (double)__ldg(g + base + sj + 1) : 0.0;
    p[7] = (i1 && j1 && k1) ? (double)__ldg(g + base + sk + sj + 1) : 0.0;
    return trilerp8(p, fx, fy, fz);
}

struct MacField {
    const float *U, *V, *W;
};

// grid geometry seen by the particle kernels: I,J,K = local array extents, Kg = global K, kOff = global
// index of local plane 0 (single GPU: K == Kg, kOff == 0)
struct GridGeom {
    int I, J, K, Kg, kOff;
};

// MACVelocityField::evaluateVelocityAtPositionLinear(vec3)  macvelocityfield.cpp:635-645
__device__ __forceinline__ void sample_velocity(const MacField &f, const GridGeom &G, double dx, double invdx,
                                                double hdx, float px, float py, float pz, float &ox, float &oy,
                                                float &oz) {
    double x = px, y = py, z = pz;
    // Grid3d::isPositionInGrid (grid3d.h:135): x < dx*i in double
    if (!(x >= 0 && y >= 0 && z >= 0 && x < dmul(dx, (double)G.I) && y < dmul(dx, (double)G.J) && z < dmul(dx, (double)G.Kg))) {
        ox = oy = oz = 0.0f;
        return;
    }
    ox = (float)sample_mac<0, 1, 1>(f.U, G.I + 1, G.J, G.K, x, y, z, dx, invdx, hdx, G.kOff);
    oy = (float)sample_mac<1, 0, 1>(f.V, G.I, G.J + 1, G.K, x, y, z, dx, invdx, hdx, G.kOff);
    oz = (float)sample_mac<1, 1, 0>(f.W, G.I, G.J, G.K + 1, x, y, z, dx, invdx, hdx, G.kOff);
}

// ---- single-precision sampling (FLIP_SAMPLING_FAST) ---------------------------------------------
// Valid when dx is a power of two and the position lies at least one cell inside the local arrays on
// every axis.  Then p/dx, the half-cell shift, floor() and the fraction are all EXACT in float (the
// shifted coordinate is a multiple of ulp(p) no larger than p; scaling by 1/dx only changes the
// exponent), so the cell indices and the interpolation weights are bit-identical to the reference's
// double-precision ones (macvelocityfield.cpp:519-613) and all 8 samples are in range.  Only the
// trilinear blend itself is evaluated in float (7 lerps, FMA) instead of double: <= a few ulp of the
// sampled values, far inside the 1e-4 rel-L2 tolerance of the parity contract.
struct FastGeom {
    float invdx, hdx;                 // 1/dx and dx/2 (exact: powers of two)
    float lox, loy, loz, hix, hiy, hiz;   // interior box [dx, (N-1)dx) of the LOCAL arrays, world coordinates
    int I, J, kOff;                   // cells per row / rows per plane of the local grid, global k of local plane 0
};

__device__ __forceinline__ bool fast_interior(const FastGeom &F, float x, float y, float z) {
    return x >= F.lox && x < F.hix && y >= F.loy && y < F.hiy && z >= F.loz && z < F.hiz;
}

__device__ __forceinline__ float trilerp_fast(const float *__restrict__ g, int base, int sj, int sk, float fx, float fy,
                                              float fz) {
    const float p000 = __ldg(g + base), p100 = __ldg(g + base + 1);
    const float p010 = __ldg(g + base + sj), p110 = __ldg(g + base + sj + 1);
    const float p001 = __ldg(g + base + sk), p101 = __ldg(g + base + sk + 1);
    const float p011 = __ldg(g + base + sk + sj), p111 = __ldg(g + base + sk + sj + 1);
    const float a = fmaf(fx, p100 - p000, p000), b = fmaf(fx, p110 - p010, p010);
    const float c = fmaf(fx, p101 - p001, p001), d = fmaf(fx, p111 - p011, p011);
    const float e = fmaf(fy, b - a, a), f = fmaf(fy, d - c, c);
    return fmaf(fz, f - e, e);
}

// cell index and fraction of the unshifted / half-cell-shifted coordinate along one axis
struct FastAxis {
    int i0, i1;       // floor(p/dx), floor((p - dx/2)/dx)
    float f0, f1;     // fractions
};
__device__ __forceinline__ FastAxis fast_axis(float p, float invdx, float hdx) {
    FastAxis a;
    const float s0 = p * invdx, s1 = (p - hdx) * invdx;
    const float q0 = floorf(s0), q1 = floorf(s1);
    a.i0 = (int)q0; a.i1 = (int)q1;
    a.f0 = s0 - q0; a.f1 = s1 - q1;
    return a;
}

struct FastStencil {
    int bu, bv, bw;            // base index of the 2x2x2 stencil in U, V, W
    float ux, uy, uz, vx, vy, vz, wx, wy, wz;   // fractions per component
};
__device__ __forceinline__ FastStencil fast_stencil(const FastGeom &F, float x, float y, float z) {
    const FastAxis ax = fast_axis(x, F.invdx, F.hdx), ay = fast_axis(y, F.invdx, F.hdx), az = fast_axis(z, F.invdx, F.hdx);
    const int k0 = az.i0 - F.kOff, k1 = az.i1 - F.kOff;
    FastStencil s;
    s.bu = ax.i0 + (F.I + 1) * (ay.i1 + F.J * k1);        // U: (x, y - h, z - h), array (I+1, J, K)
    s.bv = ax.i1 + F.I * (ay.i0 + (F.J + 1) * k1);        // V: (x - h, y, z - h), array (I, J+1, K)
    s.bw = ax.i1 + F.I * (ay.i1 + F.J * k0);              // W: (x - h, y - h, z), array (I, J, K+1)
    s.ux = ax.f0; s.uy = ay.f1; s.uz = az.f1;
    s.vx = ax.f1; s.vy = ay.f0; s.vz = az.f1;
    s.wx = ax.f1; s.wy = ay.f1; s.wz = az.f0;
    return s;
}
__device__ __forceinline__ void fast_sample(const MacField &f, const FastGeom &F, const FastStencil &s, float &ox, float &oy,
                                            float &oz) {
    ox = trilerp_fast(f.U, s.bu, F.I + 1, (F.I + 1) * F.J, s.ux, s.uy, s.uz);
    oy = trilerp_fast(f.V, s.bv, F.I, F.I * (F.J + 1), s.vx, s.vy, s.vz);
    oz = trilerp_fast(f.W, s.bw, F.I, F.I * F.J, s.wx, s.wy, s.wz);
}

// Interpolation::trilinearInterpolate(vec3 p, double dx, Array3d<float>&)  interpolation.cpp:72-110
// (node position narrowed to float, fraction = (float - float) * inv_dx in double)
struct ScalarSample {
    int i, j, k;
    double fx, fy, fz;
    float v[8];  // order 000,100,010,001,101,011,110,111
};

__device__ __forceinline__ void fetch_scalar(const float *__restrict__ g, int gi, int gj, int gk, double dx,
                                             double invdx, float px, float py, float pz, ScalarSample &s, int kOff = 0) {
    s.i = pos2idx(px, invdx);
    s.j = pos2idx(py, invdx);
    s.k = pos2idx(pz, invdx);
    float gx = (float)dmul((double)(float)s.i, dx);
    float gy = (float)dmul((double)(float)s.j, dx);
    float gz = (float)dmul((double)(float)s.k, dx);
    s.fx = dmul((double)fsub(px, gx), invdx);
    s.fy = dmul((double)fsub(py, gy), invdx);
    s.fz = dmul((double)fsub(pz, gz), invdx);
    int i = s.i, j = s.j, k = s.k - kOff;   // local plane of a z-slab
    bool i0 = (i >= 0 && i < gi), i1 = (i + 1 >= 0 && i + 1 < gi);
    bool j0 = (j >= 0 && j < gj), j1 = (j + 1 >= 0 && j + 1 < gj);
    bool k0 = (k >= 0 && k < gk), k1 = (k + 1 >= 0 && k + 1 < gk);
    long long base = (long long)i + (long long)gi * ((long long)j + (long long)gj * (long long)k);
    long long sj = gi, sk = (long long)gi * gj;
    s.v[0] = (i0 && j0 && k0) ? __ldg(g + base) : 0.0f;
    s.v[1] = (i1 && j0 && k0) ? __ldg(g + base + 1) : 0.0f;
    s.v[2] = (i0 && j1 && k0) ? __ldg(g + base + sj) : 0.0f;
    s.v[3] = (i0 && j0 && k1) ? __ldg(g + base + sk) : 0.0f;
    s.v[4] = (i1 && j0 && k1) ? __ldg(g + base + sk + 1) : 0.0f;
    s.v[5] = (i0 && j1 && k1) ? __ldg(g + base + sk + sj) : 0.0f;
    s.v[6] = (i1 && j1 && k0) ? __ldg(g + base + sj + 1) : 0.0f;
    s.v[7] = (i1 && j1 && k1) ? __ldg(g + base + sk + sj + 1) : 0.0f;
}

__device__ __forceinline__ float scalar_value(const ScalarSample &s) {
    double p[8];
#pragma unroll
    for (int m = 0; m < 8; m++) p[m] = (double)s.v[m];
    return (float)trilerp8(p, s.fx, s.fy, s.fz);
}

__device__ __forceinline__ float sample_scalar(const float *__restrict__ g, int gi, int gj, int gk, double dx,
                                               double invdx, float px, float py, float pz, int kOff = 0) {
    ScalarSample s;
    fetch_scalar(g, gi, gj, gk, dx, invdx, px, py, pz, s, kOff);
    return scalar_value(s);
}

// Interpolation::bilinearInterpolate  interpolation.cpp:188-194
__device__ __forceinline__ double bilerp(double v00, double v10, double v01, double v11, double ix, double iy) {
    double l1 = dadd(dmul(dsub(1.0, ix), v00), dmul(ix, v10));
    double l2 = dadd(dmul(dsub(1.0, ix), v01), dmul(ix, v11));
    return dadd(dmul(dsub(1.0, iy), l1), dmul(iy, l2));
}

// Interpolation::trilinearInterpolateGradient  interpolation.cpp:196-262
__device__ __forceinline__ void scalar_gradient(const ScalarSample &s, float &gx, float &gy, float &gz) {
    float v000 = s.v[0], v100 = s.v[1], v010 = s.v[2], v001 = s.v[3];
    float v101 = s.v[4], v011 = s.v[5], v110 = s.v[6], v111 = s.v[7];
    gx = (float)bilerp(fsub(v100, v000), fsub(v110, v010), fsub(v101, v001), fsub(v111, v011), s.fy, s.fz);
    gy = (float)bilerp(fsub(v010, v000), fsub(v110, v100), fsub(v011, v001), fsub(v111, v101), s.fx, s.fz);
    gz = (float)bilerp(fsub(v001, v000), fsub(v101, v100), fsub(v011, v010), fsub(v111, v110), s.fx, s.fy);
}

// Warp-aggregated append: the lanes that reach this call together take consecutive slots with ONE atomic
// (queues fed by hundreds of thousands of threads would otherwise serialise on a single address in L2).
__device__ __forceinline__ int warp_append_slot(int *counter) {
    const unsigned int m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// Block-aggregated append: EVERY thread of the block calls it together with the number of items it holds; the return
// value is the slot of the thread's first item.  One global atomic per block -- appends of a few lanes at a time from
// a whole grid (warp_append_slot in divergent code) serialise on the counter's address in L2.
template <int THREADS>
__device__ __forceinline__ int block_append_slots(int *counter, int mine) {
    __shared__ int warpTot[THREADS / 32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warpTot[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) { const int v = warpTot[w]; warpTot[w] = tot; tot += v; }
        base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    const int slot = base + warpTot[wid] + incl - mine;
    __syncthreads();        // the shared words are free for the next call
    return slot;
}

// warp reductions
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_maxd(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace flip
