// Variational pressure projection (PressureSolver::solve, pressuresolver.cpp:44-77) as a
// matrix-free 7-point PCG on the GPU.
//
// Layout.  All solver vectors are dense over the cells of the grid and indexed by the flat cell id
// c = i + I*(j + J*k), so the six neighbours of a row are at c+-1, c+-I, c+-I*J and no column
// indices are stored.  Work is driven by the list of ACTIVE SEGMENTS: aligned runs of 32
// consecutive cell ids that contain at least one pressure row, each with a 32-bit lane mask; one
// warp processes one segment.  The matrix is kept as the diagonal (fp64) plus, per cell, the three
// float face weights towards +i,+j,+k masked to zero where the neighbour is not a row; the
// off-diagonal is -(double)w * dt/dx^2, formed on the fly exactly as the reference forms it
// (pressuresolver.cpp:723-808).  Algorithmic traffic of one operator application: 36 B/row.
//
// The reference's preconditioner is a serial MIC(0) (pcgsolver.h:69-221); here it is replaced by a
// GPU-parallel one (Jacobi, or an aggregation multigrid V-cycle in fp32).  The Krylov vectors,
// the residual test (||r||_inf <= tol*||b||_inf, pcgsolver.h:262-288), the iteration cap and the
// "acceptable" fallback (pressuresolver.cpp:810-840) are the reference's, in fp64.
#include <cooperative_groups.h>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include "device_math.cuh"
#include "flip_internal.h"

namespace cg = cooperative_groups;

namespace flip {

static constexpr int TPB = 256;
static constexpr int WPB = TPB / 32;

struct PGrid {
    int I, J, K;  // local grid
    int sj, sk;   // strides I, I*J
    int kOff, Kg; // z-slab: global index of local plane 0, global K (single GPU: 0, K)
    int kOwn0, kOwn1;   // owned local planes
    int cut;            // z-slabs whose boundaries are not aligned to the multigrid aggregates: the V-cycle is
                        // block-local per slab and rows beyond the cut take no part in it
};

// a pressure row of the GLOBAL system (pressuresolver.cpp:101-110): liquid cell in [1,N-2]^3
__device__ __forceinline__ bool is_row(const float *__restrict__ phi, const PGrid &g, int c) {
    int i = c % g.I;
    int j = (c / g.I) % g.J;
    int k = c / g.sk + g.kOff;
    return i >= 1 && j >= 1 && k >= 1 && i < g.I - 1 && j < g.J - 1 && k < g.Kg - 1 && __ldg(phi + c) < 0.0f;
}
// ... that this slab owns
__device__ __forceinline__ bool is_owned(const PGrid &g, int c) {
    int k = c / g.sk;
    return k >= g.kOwn0 && k < g.kOwn1;
}

// ---- 1. rows -> segments (pressuresolver.cpp:101-114: cells with phi<0 in [1,N-2]^3, (k,j,i) order)
__global__ void k_seg_flag(const float *__restrict__ phi, PGrid g, int nC, int nSeg, unsigned int *__restrict__ mask,
                           int *__restrict__ flag, DeviceScalars *S) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= nSeg) return;
    int c = warp * 32 + lane;
    bool row = (c < nC) && is_owned(g, c) && is_row(phi, g, c);
    unsigned int m = __ballot_sync(0xffffffffu, row);
    if (lane == 0) {
        mask[warp] = m;
        flag[warp] = m ? 1 : 0;
        if (m) atomicAdd(&S->numRows, __popc(m));
    }
}

__global__ void k_seg_compact(int nSeg, const unsigned int *__restrict__ mask, const int *__restrict__ flag,
                              const int *__restrict__ pos, int *__restrict__ segCell, unsigned int *__restrict__ segMask,
                              DeviceScalars *S) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nSeg) return;
    if (flag[s]) {
        int o = pos[s];
        segCell[o] = s * 32;
        segMask[o] = mask[s];
    }
    if (s == nSeg - 1) S->numSegments = pos[s] + flag[s];
}

// ---- 2. right-hand side and matrix
// ---- moving solids: _conditionSolidVelocityField (pressuresolver.cpp:124-244).  Liquid cells of [1,N-2]^3 are joined
// through faces of weight >= 1e-6; a joined region of more than one cell none of whose cells has an air neighbour
// across such a face (_computeBordersAirGridThread :246-268) is an enclosed pocket, and the solid velocities on the six
// faces of each of its cells are set to zero (in the stored arrays, as there).  The reference flood-fills region by
// region on one thread; here "reaches air" spreads from the air-bordering cells, one cell per pass and direction, until
// a pass changes nothing, and the cells it never reached are the pockets.
// flag: 0 not a liquid interior cell, 1 liquid and not (yet) reached, 2 reached.
#define POCKET_EPS 1e-6f
__device__ __forceinline__ bool pocket_interior(const PGrid &g, int i, int j, int k) {
    return i >= 1 && j >= 1 && k >= 1 && i < g.I - 1 && j < g.J - 1 && k < g.K - 1;
}
__global__ void k_pocket_init(PGrid g, int nC, const float *__restrict__ phi, const float *__restrict__ wU,
                              const float *__restrict__ wV, const float *__restrict__ wW, unsigned char *__restrict__ flag) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int i = c % g.I, j = (c / g.I) % g.J, k = c / g.sk;
    unsigned char f = 0;
    if (pocket_interior(g, i, j, k) && phi[c] < 0.0f) {
        const int fu = i + (g.I + 1) * (j + g.J * k), fv = i + g.I * (j + (g.J + 1) * k);
        const bool air = (wU[fu] >= POCKET_EPS && phi[c - 1] >= 0.0f) || (wU[fu + 1] >= POCKET_EPS && phi[c + 1] >= 0.0f) ||
                         (wV[fv] >= POCKET_EPS && phi[c - g.sj] >= 0.0f) || (wV[fv + g.I] >= POCKET_EPS && phi[c + g.sj] >= 0.0f) ||
                         (wW[c] >= POCKET_EPS && phi[c - g.sk] >= 0.0f) || (wW[c + g.sk] >= POCKET_EPS && phi[c + g.sk] >= 0.0f);
        f = air ? 2 : 1;
    }
    flag[c] = f;
}
__global__ void k_pocket_spread(PGrid g, int nC, const float *__restrict__ wU, const float *__restrict__ wV,
                                const float *__restrict__ wW, volatile unsigned char *flag, int *changed) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC || flag[c] != 1) return;
    const int i = c % g.I, j = (c / g.I) % g.J, k = c / g.sk;
    const int fu = i + (g.I + 1) * (j + g.J * k), fv = i + g.I * (j + (g.J + 1) * k);
    // (a neighbour with flag 2 is a liquid interior cell)
    const bool reached = (flag[c - 1] == 2 && wU[fu] >= POCKET_EPS) || (flag[c + 1] == 2 && wU[fu + 1] >= POCKET_EPS) ||
                         (flag[c - g.sj] == 2 && wV[fv] >= POCKET_EPS) || (flag[c + g.sj] == 2 && wV[fv + g.I] >= POCKET_EPS) ||
                         (flag[c - g.sk] == 2 && wW[c] >= POCKET_EPS) || (flag[c + g.sk] == 2 && wW[c + g.sk] >= POCKET_EPS);
    if (reached) { flag[c] = 2; *changed = 1; }
}
__global__ void k_pocket_zero(PGrid g, int nC, const float *__restrict__ wU, const float *__restrict__ wV,
                              const float *__restrict__ wW, const unsigned char *__restrict__ flag,
                              float *__restrict__ solU, float *__restrict__ solV, float *__restrict__ solW) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC || flag[c] != 1) return;
    const int i = c % g.I, j = (c / g.I) % g.J, k = c / g.sk;
    const int fu = i + (g.I + 1) * (j + g.J * k), fv = i + g.I * (j + (g.J + 1) * k);
    // joined to another cell of the pocket?  (a neighbour with a nonzero flag is a liquid interior cell; next to an
    // unreached cell it is unreached too)
    const bool joined = (flag[c - 1] && wU[fu] >= POCKET_EPS) || (flag[c + 1] && wU[fu + 1] >= POCKET_EPS) ||
                        (flag[c - g.sj] && wV[fv] >= POCKET_EPS) || (flag[c + g.sj] && wV[fv + g.I] >= POCKET_EPS) ||
                        (flag[c - g.sk] && wW[c] >= POCKET_EPS) || (flag[c + g.sk] && wW[c + g.sk] >= POCKET_EPS);
    if (!joined) return;        // a region of one cell is left alone (:219)
    solU[fu] = 0.0f; solU[fu + 1] = 0.0f;
    solV[fv] = 0.0f; solV[fv + g.I] = 0.0f;
    solW[c] = 0.0f; solW[c + g.sk] = 0.0f;
}

static void condition_solid_velocities(flip_ctx *c, const PGrid &g) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    const int blocks = cdiv(d.nC, TPB);
    int *changed = &c->dS->pocketChanged;
    k_pocket_init<<<blocks, TPB, 0, st>>>(g, d.nC, c->phiL, c->wU, c->wV, c->wW, c->pocketFlag);
    c->launches++;
    for (;;) {
        FLIP_CUDA_CHECK(cudaMemsetAsync(changed, 0, sizeof(int), st));
        for (int pass = 0; pass < 16; pass++) k_pocket_spread<<<blocks, TPB, 0, st>>>(g, d.nC, c->wU, c->wV, c->wW, c->pocketFlag, changed);
        c->launches += 16;
        int h = 0;
        FLIP_CUDA_CHECK(cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, st));
        FLIP_CUDA_CHECK(cudaStreamSynchronize(st));
        if (!h) break;
    }
    k_pocket_zero<<<blocks, TPB, 0, st>>>(g, d.nC, c->wU, c->wV, c->wW, c->pocketFlag, c->solU, c->solV, c->solW);
    c->launches++;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

struct BuildParams {
    PGrid g;
    double invdx;     // 1.0/_dx                    pressuresolver.cpp:582
    double factor;    // _deltaTime/(_dx*_dx)       pressuresolver.cpp:725
};

__global__ void k_build_system(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                               BuildParams bp, const float *__restrict__ phi,
                               const float *__restrict__ U, const float *__restrict__ V, const float *__restrict__ W,
                               const float *__restrict__ wU, const float *__restrict__ wV, const float *__restrict__ wW,
                               double *__restrict__ Adiag, float *__restrict__ AoffU, float *__restrict__ AoffV,
                               float *__restrict__ AoffW, double *__restrict__ b, double *__restrict__ x,
                               DeviceScalars *S, const unsigned int *__restrict__ prevRowBits,
                               const float *__restrict__ solU, const float *__restrict__ solV,
                               const float *__restrict__ solW, const float *__restrict__ wC) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= S->numSegments) return;
    const PGrid &g = bp.g;
    int c = segCell[warp] + lane;
    bool row = (segMask[warp] >> lane) & 1u;
    double babs = 0.0;
    if (row) {
        int i = c % g.I, j = (c / g.I) % g.J, k = c / g.sk;
        long long fu = (long long)i + (long long)(g.I + 1) * (j + (long long)g.J * k);
        long long fv = (long long)i + (long long)g.I * (j + (long long)(g.J + 1) * k);
        long long fw = c;
        double volRight = wU[fu + 1], volLeft = wU[fu];
        double volTop = wV[fv + g.I], volBottom = wV[fv];
        double volFront = wW[fw + g.sk], volBack = wW[fw];
        // _calculateNegativeDivergenceVectorThread  pressuresolver.cpp:580-613 (solids at rest: the
        // solid-velocity terms are +-0 and are skipped)
        double f = bp.invdx;
        double div = 0.0;
        div = dadd(div, dmul(dmul(-f, volRight), (double)U[fu + 1]));
        div = dadd(div, dmul(dmul(f, volLeft), (double)U[fu]));
        div = dadd(div, dmul(dmul(-f, volTop), (double)V[fv + g.I]));
        div = dadd(div, dmul(dmul(f, volBottom), (double)V[fv]));
        div = dadd(div, dmul(dmul(-f, volFront), (double)W[fw + g.sk]));
        div = dadd(div, dmul(dmul(f, volBack), (double)W[fw]));
        if (solU) {
            // moving solids (pressuresolver.cpp:595, :608-613): +-factor * (w_face - w_centre) * u_solid, in this order
            const double volCenter = wC[c];
            div = dadd(div, dmul(dmul(f, dsub(volRight, volCenter)), (double)solU[fu + 1]));
            div = dadd(div, dmul(dmul(-f, dsub(volLeft, volCenter)), (double)solU[fu]));
            div = dadd(div, dmul(dmul(f, dsub(volTop, volCenter)), (double)solV[fv + g.I]));
            div = dadd(div, dmul(dmul(-f, dsub(volBottom, volCenter)), (double)solV[fv]));
            div = dadd(div, dmul(dmul(f, dsub(volFront, volCenter)), (double)solW[fw + g.sk]));
            div = dadd(div, dmul(dmul(-f, dsub(volBack, volCenter)), (double)solW[fw]));
        }
        b[c] = div;
        // initial guess: 0 as in the reference (pcgsolver.h:258), or -- warm start -- the pressure this cell had in
        // the previous solve if it was a row then (segments are aligned: bit `lane` of word c/32)
        if (!(prevRowBits && ((prevRowBits[c >> 5] >> lane) & 1u))) x[c] = 0.0;
        babs = fabs(div);

        // _calculateMatrixCoefficientsThread  pressuresolver.cpp:723-808
        const double eps = 1e-9, maxtheta = 25.0;
        double factor = bp.factor;
        double phiC = phi[c];
        double diag = dmul(dadd(dadd(dadd(dadd(dadd(volRight, volLeft), volTop), volBottom), volFront), volBack), factor);
        auto side = [&](int nb, double vol, bool &nbRow) {
            double phin = phi[nb];
            nbRow = false;
            if (phin < 0.0) {
                nbRow = is_row(phi, g, nb);
            } else {
                double theta = phin / dadd(phiC, eps);
                theta = fmax(-maxtheta, fmin(theta, maxtheta));
                diag = dsub(diag, dmul(dmul(vol, factor), theta));
            }
        };
        bool rR, rL, rT, rB, rF, rK;
        side(c + 1, volRight, rR);
        side(c - 1, volLeft, rL);
        side(c + g.sj, volTop, rT);
        side(c - g.sj, volBottom, rB);
        side(c + g.sk, volFront, rF);
        side(c - g.sk, volBack, rK);
        diag = fmax(diag, 0.0);
        Adiag[c] = diag;
        AoffU[c] = rR ? (float)volRight : 0.0f;
        AoffV[c] = rT ? (float)volTop : 0.0f;
        AoffW[c] = rF ? (float)volFront : 0.0f;
        // entries this row reads that no row of this slab owns
        if (!rL) AoffU[c - 1] = 0.0f;
        if (!rB) AoffV[c - g.sj] = 0.0f;
        if (!rK) AoffW[c - g.sk] = 0.0f;
        else if (!is_owned(g, c - g.sk)) AoffW[c - g.sk] = (float)volBack;   // row of the lower neighbouring slab
    }
    babs = warp_maxd(babs);
    if (lane == 0 && babs > 0.0) atomicMax(&S->rhsMaxBits, (unsigned long long)__double_as_longlong(babs));
}

// y = A v at a row (off-diagonals only where the stored weight is nonzero: non-row neighbours may
// hold stale data)
__device__ __forceinline__ double apply_row(const PGrid &g, int c, double factor, const double *__restrict__ Adiag,
                                            const float *__restrict__ AoffU, const float *__restrict__ AoffV,
                                            const float *__restrict__ AoffW, const double *__restrict__ v) {
    // all 14 loads are issued before any is consumed (rows are interior cells, so every index is in
    // range); a stale vector entry behind a zero weight is discarded by the select, never multiplied
    float a0 = AoffW[c - g.sk], a1 = AoffV[c - g.sj], a2 = AoffU[c - 1], a3 = AoffU[c], a4 = AoffV[c], a5 = AoffW[c];
    double v0 = v[c - g.sk], v1 = v[c - g.sj], v2 = v[c - 1], v3 = v[c + 1], v4 = v[c + g.sj], v5 = v[c + g.sk];
    double acc = Adiag[c] * v[c];
    double off = 0.0;
    off += (a0 != 0.0f) ? (double)a0 * v0 : 0.0;
    off += (a1 != 0.0f) ? (double)a1 * v1 : 0.0;
    off += (a2 != 0.0f) ? (double)a2 * v2 : 0.0;
    off += (a3 != 0.0f) ? (double)a3 * v3 : 0.0;
    off += (a4 != 0.0f) ? (double)a4 * v4 : 0.0;
    off += (a5 != 0.0f) ? (double)a5 * v5 : 0.0;
    return acc - factor * off;
}

struct PcgParams {
    PGrid g;
    double factor;
    double tolFactor;
};

__device__ __forceinline__ void block_add(double v, double *target) {
    __shared__ double sh[WPB];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < WPB) ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0 && t != 0.0) atomicAdd(target, t);
    }
    __syncthreads();
}
__device__ __forceinline__ void block_max(double v, unsigned long long *target) {
    __shared__ double shm[WPB];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_maxd(v);
    if (lane == 0) shm[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < WPB) ? shm[lane] : 0.0;
        t = warp_maxd(t);
        if (lane == 0 && t > 0.0) atomicMax(target, (unsigned long long)__double_as_longlong(t));
    }
    __syncthreads();
}

// r = b [- A x0 with a warm start]; z = M^-1 r (Jacobi); s = z; rho = z.r     (pcgsolver.h:258-276)
__global__ void k_pcg_init(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask, PcgParams pp,
                           const double *__restrict__ b, const double *__restrict__ Adiag, double *__restrict__ r,
                           double *__restrict__ z, double *__restrict__ s, DeviceScalars *S, int jacobi,
                           const float *__restrict__ AoffU, const float *__restrict__ AoffV, const float *__restrict__ AoffW,
                           const double *__restrict__ x0) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    double part = 0.0;
    if (warp < S->numSegments) {
        int c = segCell[warp] + lane;
        if ((segMask[warp] >> lane) & 1u) {
            double rv = b[c];
            if (x0) rv -= apply_row(pp.g, c, pp.factor, Adiag, AoffU, AoffV, AoffW, x0);
            r[c] = rv;
            if (jacobi) {
                double d = Adiag[c];
                double zv = (d != 0.0) ? rv / d : 0.0;
                z[c] = zv;
                s[c] = zv;
                part = zv * rv;
            }
        }
    }
    if (jacobi) block_add(part, &S->rho[0]);
}

__global__ void k_pcg_scalars_init(DeviceScalars *S, double tolFactor) {
    double bmax = __longlong_as_double((long long)S->rhsMaxBits);
    S->pcgTol = tolFactor * bmax;
    S->pcgError = bmax;
    S->pcgIterations = 0;
    S->pcgDone = 0;
    for (int q = 0; q < 3; q++) { S->dotSZ[q] = 0.0; S->rho[q] = 0.0; S->rMaxBits[q] = 0ull; }
}

// The PCG passes below run grid-stride over the active segments with a grid capped at one resident wave
// (seg_blocks): per-thread partial sums span several segments, so a pass issues ~1 k instead of ~9 k
// same-address atomics for its dot product / norm.

// z = A s ; dotSZ += s.z
// (two segments per loop trip: the loads of the second do not wait for the first to retire)
__global__ void __launch_bounds__(TPB, 6) k_pcg_spmv(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                                                     PcgParams pp, const double *__restrict__ Adiag,
                                                     const float *__restrict__ AoffU, const float *__restrict__ AoffV,
                                                     const float *__restrict__ AoffW, const double *__restrict__ s,
                                                     double *__restrict__ z, DeviceScalars *S, int it) {
    if (S->pcgDone) return;
    const int lane = threadIdx.x & 31;
    const int nseg = S->numSegments, nw = (gridDim.x * blockDim.x) >> 5;
    double part = 0.0;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nseg; w += 2 * nw) {
        const int w2 = w + nw;
        const bool has2 = w2 < nseg;
        const int cA = segCell[w] + lane;
        const unsigned int mA = segMask[w];
        const int cB = has2 ? segCell[w2] + lane : cA;
        const unsigned int mB = has2 ? segMask[w2] : 0u;
        const bool rA = (mA >> lane) & 1u, rB = (mB >> lane) & 1u;
        double zA = 0.0, zB = 0.0, sA = 0.0, sB = 0.0;
        if (rA) { zA = apply_row(pp.g, cA, pp.factor, Adiag, AoffU, AoffV, AoffW, s); sA = s[cA]; }
        if (rB) { zB = apply_row(pp.g, cB, pp.factor, Adiag, AoffU, AoffV, AoffW, s); sB = s[cB]; }
        if (rA) { z[cA] = zA; part += sA * zA; }
        if (rB) { z[cB] = zB; part += sB * zB; }
    }
    block_add(part, &S->dotSZ[it % 3]);
}

// alpha = rho/(s.z); x += alpha s; r -= alpha z; rmax = ||r||_inf; [Jacobi: z = r/diag; rhoNew += z.r]
__global__ void __launch_bounds__(TPB, 5) k_pcg_update(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                                                       const double *__restrict__ Adiag, const double *__restrict__ s,
                                                       double *__restrict__ z, double *__restrict__ x, double *__restrict__ r,
                                                       DeviceScalars *S, int it, int jacobi) {
    if (S->pcgDone) return;
    const int lane = threadIdx.x & 31;
    const int nseg = S->numSegments, nw = (gridDim.x * blockDim.x) >> 5;
    const double alpha = S->rho[it % 3] / S->dotSZ[it % 3];
    double part = 0.0, rabs = 0.0;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nseg; w += 2 * nw) {
        const int w2 = w + nw;
        const bool has2 = w2 < nseg;
        const int cA = segCell[w] + lane;
        const unsigned int mA = segMask[w];
        const int cB = has2 ? segCell[w2] + lane : cA;
        const unsigned int mB = has2 ? segMask[w2] : 0u;
        const bool rA = (mA >> lane) & 1u, rB = (mB >> lane) & 1u;
        double xA = 0.0, sA = 0.0, qA = 0.0, zA = 0.0, xB = 0.0, sB = 0.0, qB = 0.0, zB = 0.0, dA = 1.0, dB = 1.0;
        if (rA) { xA = x[cA]; sA = s[cA]; qA = r[cA]; zA = z[cA]; if (jacobi) dA = Adiag[cA]; }
        if (rB) { xB = x[cB]; sB = s[cB]; qB = r[cB]; zB = z[cB]; if (jacobi) dB = Adiag[cB]; }
        if (rA) {
            x[cA] = xA + alpha * sA;
            double rv = qA - alpha * zA;
            r[cA] = rv;
            rabs = fmax(rabs, fabs(rv));
            if (jacobi) {
                double zv = (dA != 0.0) ? rv / dA : 0.0;
                z[cA] = zv;
                part += zv * rv;
            }
        }
        if (rB) {
            x[cB] = xB + alpha * sB;
            double rv = qB - alpha * zB;
            r[cB] = rv;
            rabs = fmax(rabs, fabs(rv));
            if (jacobi) {
                double zv = (dB != 0.0) ? rv / dB : 0.0;
                z[cB] = zv;
                part += zv * rv;
            }
        }
    }
    block_max(rabs, &S->rMaxBits[it % 3]);
    if (jacobi) block_add(part, &S->rho[(it + 1) % 3]);
}

// rhoNew = z.r for preconditioners that produce z in a separate pass
__global__ void k_pcg_dot_zr(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                             const double *__restrict__ z, const double *__restrict__ r, DeviceScalars *S, int slot) {
    if (S->pcgDone) return;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    double part = 0.0;
    if (warp < S->numSegments) {
        int c = segCell[warp] + lane;
        if ((segMask[warp] >> lane) & 1u) part = z[c] * r[c];
    }
    block_add(part, &S->rho[slot]);
}

// convergence test, beta = rhoNew/rho, s = z + beta s   (pcgsolver.h:286-297)
__global__ void k_pcg_direction(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                                const double *__restrict__ z, double *__restrict__ s, DeviceScalars *S, int it) {
    if (S->pcgDone) return;
    double rmax = __longlong_as_double((long long)S->rMaxBits[it % 3]);
    double rho = S->rho[it % 3], rhoNew = S->rho[(it + 1) % 3];
    bool converged = rmax <= S->pcgTol;
    bool breakdown = !converged && (rhoNew == 0.0 || rhoNew != rhoNew);
    const int lane = threadIdx.x & 31;
    if (!converged && !breakdown) {
        const double beta = rhoNew / rho;
        const int nseg = S->numSegments, nw = (gridDim.x * blockDim.x) >> 5;
        for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nseg; w += 2 * nw) {
            const int w2 = w + nw;
            const bool has2 = w2 < nseg;
            const int cA = segCell[w] + lane;
            const unsigned int mA = segMask[w];
            const int cB = has2 ? segCell[w2] + lane : cA;
            const unsigned int mB = has2 ? segMask[w2] : 0u;
            const bool rA = (mA >> lane) & 1u, rB = (mB >> lane) & 1u;
            double zA = 0.0, sA = 0.0, zB = 0.0, sB = 0.0;
            if (rA) { zA = z[cA]; sA = s[cA]; }
            if (rB) { zB = z[cB]; sB = s[cB]; }
            if (rA) s[cA] = zA + beta * sA;
            if (rB) s[cB] = zB + beta * sB;
        }
    }
    // the last block to finish publishes the scalars (no other block reads them after this point in
    // this launch: every block sampled them above, before any block can reach the ticket below only
    // if all blocks have passed -> use a ticket counter)
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int ticket = atomicAdd(&S->pad[0], 1);
        last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        S->pad[0] = 0;
        S->pcgIterations = it + 1;
        S->pcgError = rmax;
        if (converged) S->pcgDone = 1;
        else if (breakdown) S->pcgDone = 3;
        S->dotSZ[(it + 1) % 3] = 0.0;
        S->rMaxBits[(it + 1) % 3] = 0ull;
        S->rho[(it + 2) % 3] = 0.0;
    }
}


// ------------------------------------------------------------------------------------------------
// Multigrid preconditioner: z = M^-1 r as one symmetric V-cycle of an aggregation multigrid.
//
// Level 0 is the pressure system itself (rows = liquid cells, driven by the active-segment list).
// Level l+1 aggregates 2x2x2 cells of level l; its operator is the Galerkin product P^T A P with the
// piecewise-constant P, which for a 7-point operator is again a 7-point operator: the coarse
// off-diagonal towards +i is the sum of the fine off-diagonals crossing that aggregate face, the
// coarse diagonal is the sum of the children's diagonals minus twice the off-diagonals interior to
// the aggregate.  Free surface (ghost-fluid diagonal terms) and solid walls (zero face weights) are
// carried to every level by construction, so no geometric special-casing is needed.  Smoother:
// damped Jacobi, nu sweeps before and after the coarse correction (first sweep from a zero guess,
// first post sweep fused with the prolongation); the coarse correction is over-weighted by a
// constant (piecewise-constant aggregation under-estimates smooth corrections by ~2x).  Everything
// is a fixed symmetric linear operator, as CG requires.  Levels hold fp32; the Krylov vectors, the
// residual and the stopping test stay fp64 and are the reference's (pcgsolver.h:248-302).
// Levels with <= MG_SMALL cells are run by one CTA in a single launch.
// ------------------------------------------------------------------------------------------------
static constexpr int MG_MAX_LEVELS = 12;
// zero padding (floats) in front of and behind every array of a level >= 1: one plane + one cell, 256-byte aligned
inline int mg_pad(int sk) { return ((sk + 1 + 63) / 64) * 64; }
static constexpr int MG_SMALL = 4096;

struct MgLevel {
    int I, J, K, sj, sk, n;
    int cut;   // see PGrid::cut
    float *diag, *invD, *oU, *oV, *oW, *x, *x2, *b;
};

static constexpr int MG_MAX_SWEEPS = 8;
struct MgParams {
    float om[MG_MAX_SWEEPS];   // damping of pre-sweep s; post-sweep s uses om[nu-1-s] (the V-cycle stays symmetric)
    float omegaCoarse;         // damping of the sweeps on the coarsest level
    float scale;      // coarse-correction weight
    int nu;           // pre = post sweeps
    int coarseSweeps;
};

// sum_nb off(c,nb) * x(nb) on a dense level (bounds-checked: coarse cells can sit on the grid border)
__device__ __forceinline__ float mg_offsum(const MgLevel &L, const float *x, int c, int i, int j, int k) {
    float s = 0.0f, a;
    if (k > 0) { a = L.oW[c - L.sk]; if (a != 0.0f) s += a * x[c - L.sk]; }
    if (j > 0) { a = L.oV[c - L.sj]; if (a != 0.0f) s += a * x[c - L.sj]; }
    if (i > 0) { a = L.oU[c - 1];    if (a != 0.0f) s += a * x[c - 1]; }
    a = L.oU[c]; if (a != 0.0f) s += a * x[c + 1];
    a = L.oV[c]; if (a != 0.0f) s += a * x[c + L.sj];
    a = L.oW[c]; if (a != 0.0f) s += a * x[c + L.sk];
    return s;
}

// x + scale * (P e)(c) at a cell and at its six neighbours is what the fused prolongation+sweep reads
__device__ __forceinline__ float mg_corrected(const MgLevel &L, const MgLevel &C, const float *x, const float *e,
                                              float scale, int c, int i, int j, int k) {
    // an inactive cell (no row below it: outside the liquid, or beyond a z-slab cut) stays zero
    if (L.cut && L.invD[c] == 0.0f) return 0.0f;
    return x[c] + scale * e[(i >> 1) + C.sj * (j >> 1) + C.sk * (k >> 1)];
}

__device__ __forceinline__ float mg_offsum_corrected(const MgLevel &L, const MgLevel &C, const float *x, const float *e,
                                                     float scale, int c, int i, int j, int k) {
    float s = 0.0f, a;
    if (k > 0) { a = L.oW[c - L.sk]; if (a != 0.0f) s += a * mg_corrected(L, C, x, e, scale, c - L.sk, i, j, k - 1); }
    if (j > 0) { a = L.oV[c - L.sj]; if (a != 0.0f) s += a * mg_corrected(L, C, x, e, scale, c - L.sj, i, j - 1, k); }
    if (i > 0) { a = L.oU[c - 1];    if (a != 0.0f) s += a * mg_corrected(L, C, x, e, scale, c - 1, i - 1, j, k); }
    a = L.oU[c]; if (a != 0.0f) s += a * mg_corrected(L, C, x, e, scale, c + 1, i + 1, j, k);
    a = L.oV[c]; if (a != 0.0f) s += a * mg_corrected(L, C, x, e, scale, c + L.sj, i, j + 1, k);
    a = L.oW[c]; if (a != 0.0f) s += a * mg_corrected(L, C, x, e, scale, c + L.sk, i, j, k + 1);
    return s;
}

// the iterate after one sweep from a zero guess, at any cell q: what mode 3 reads instead of a stored array
__device__ __forceinline__ float mg_first(const MgLevel &L, int q, float omega) {
    float inv = L.invD[q];
    return (inv == 0.0f) ? 0.0f : omega * inv * L.b[q];
}
__device__ __forceinline__ float mg_offsum_first(const MgLevel &L, int c, int i, int j, int k, float omega) {
    float s = 0.0f, a;
    if (k > 0) { a = L.oW[c - L.sk]; if (a != 0.0f) s += a * mg_first(L, c - L.sk, omega); }
    if (j > 0) { a = L.oV[c - L.sj]; if (a != 0.0f) s += a * mg_first(L, c - L.sj, omega); }
    if (i > 0) { a = L.oU[c - 1];    if (a != 0.0f) s += a * mg_first(L, c - 1, omega); }
    a = L.oU[c]; if (a != 0.0f) s += a * mg_first(L, c + 1, omega);
    a = L.oV[c]; if (a != 0.0f) s += a * mg_first(L, c + L.sj, omega);
    a = L.oW[c]; if (a != 0.0f) s += a * mg_first(L, c + L.sk, omega);
    return s;
}

// one damped-Jacobi sweep at cell c of a dense level.  mode 0: from a zero guess; 1: regular;
// 2: regular on (xin + scale * P e); 3: the first TWO sweeps from a zero guess in one pass (the first one is
// pointwise, so its result at the six neighbours is recomputed instead of being stored and re-read)
__device__ __forceinline__ float mg_sweep_cell(const MgLevel &L, const MgLevel &C, const float *xin, const float *e,
                                               float omega, float scale, int mode, int c, int i, int j, int k,
                                               float omega0 = 0.0f) {
    float inv = L.invD[c];
    if (inv == 0.0f) return 0.0f;
    float b = L.b[c];
    if (mode == 0) return omega * inv * b;
    float xc, ns;
    if (mode == 3) { xc = omega0 * inv * b; ns = mg_offsum_first(L, c, i, j, k, omega0); }
    else if (mode == 1) { xc = xin[c]; ns = mg_offsum(L, xin, c, i, j, k); }
    else { xc = mg_corrected(L, C, xin, e, scale, c, i, j, k); ns = mg_offsum_corrected(L, C, xin, e, scale, c, i, j, k); }
    return (1.0f - omega) * xc + omega * inv * (b + ns);
}

// residual b - A x at a cell of a dense level
__device__ __forceinline__ float mg_residual_cell(const MgLevel &L, const float *x, int c, int i, int j, int k) {
    if (L.invD[c] == 0.0f) return 0.0f;
    return L.b[c] - (L.diag[c] * x[c] - mg_offsum(L, x, c, i, j, k));
}

// b_coarse(C) = sum over the 8 children of the fine residual
__device__ __forceinline__ float mg_restrict_cell(const MgLevel &F, const float *x, int ci, int cj, int ck) {
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int i = 2 * ci + (q & 1), j = 2 * cj + ((q >> 1) & 1), k = 2 * ck + (q >> 2);
        if (i < F.I && j < F.J && k < F.K) s += mg_residual_cell(F, x, i + F.sj * j + F.sk * k, i, j, k);
    }
    return s;
}

__global__ void k_mg_sweep(MgLevel L, MgLevel C, const float *xin, const float *e, float *xout, float omega,
                           float scale, int mode, const DeviceScalars *S) {
    if (S->pcgDone) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L.n) return;
    int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
    xout[c] = mg_sweep_cell(L, C, xin, e, omega, scale, mode, c, i, j, k);
}

__global__ void k_mg_restrict(MgLevel F, MgLevel C, const float *x, const DeviceScalars *S) {
    if (S->pcgDone) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C.n) return;
    if (C.invD[c] == 0.0f) { C.b[c] = 0.0f; return; }
    int i = c % C.I, j = (c / C.I) % C.J, k = c / C.sk;
    C.b[c] = mg_restrict_cell(F, x, i, j, k);
}

// ---- the same passes driven by the list of ACTIVE 32-cell segments of a level (rebuilt every solve): most of
// the box is air, and a dense launch over a coarse level spends its time finding that out
__global__ void k_mg_build_list(MgLevel L, int *__restrict__ list, int *__restrict__ count) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int c = warp * 32 + lane;
    bool act = (c < L.n) && L.invD[c] != 0.0f;
    unsigned int m = __ballot_sync(0xffffffffu, act);
    if (lane == 0 && m) list[atomicAdd(count, 1)] = warp * 32;
}

__global__ void k_mg_sweep_list(MgLevel L, MgLevel C, const float *xin, const float *e, float *xout, float omega,
                                float scale, int mode, const int *__restrict__ list, const int *__restrict__ count,
                                const DeviceScalars *S, float omega0) {
    if (S->pcgDone) return;
    const int nseg = *count;
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nseg; s += nw) {
        int c = list[s] + lane;
        if (c < L.n) {
            int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
            xout[c] = mg_sweep_cell(L, C, xin, e, omega, scale, mode, c, i, j, k, omega0);
        }
    }
}

// b_C = P^T (b_F - A_F x_F) with EIGHT threads per coarse cell (one per child, summed by a fixed shuffle tree:
// deterministic), coarse cells taken from the active list of level C; one 256-thread block per coarse segment
__global__ void __launch_bounds__(256) k_mg_restrict_list(MgLevel F, MgLevel C, const float *x,
                                                          const int *__restrict__ list, const int *__restrict__ count,
                                                          const DeviceScalars *S) {
    if (S->pcgDone) return;
    const int nseg = *count;
    const int t = threadIdx.x, q = t & 7;
    for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
        const int cc = list[s] + (t >> 3);
        float v = 0.0f;
        const bool act = (cc < C.n) && C.invD[cc] != 0.0f;
        if (act) {
            int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
            int i = 2 * ci + (q & 1), j = 2 * cj + ((q >> 1) & 1), k = 2 * ck + (q >> 2);
            if (i < F.I && j < F.J && k < F.K) v = mg_residual_cell(F, x, i + F.sj * j + F.sk * k, i, j, k);
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (q == 0 && cc < C.n) C.b[cc] = act ? v : 0.0f;
    }
}

struct MgSmallArgs {
    MgLevel lv[MG_MAX_LEVELS];
    int first, last;   // levels first..last are run by this launch (b of `first` is already set)
    MgParams p;
};

// the small levels of the V-cycle in one CTA: down, coarsest sweeps, up.  Leaves the result in lv[first].x.
__device__ __forceinline__ void mg_small_body(const MgLevel *lv, int first, int last, const MgParams &p) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const float scale = p.scale;
    for (int l = first; l <= last; l++) {
        const MgLevel &L = lv[l];
        int sweeps = (l == last) ? p.coarseSweeps : p.nu;
        float *xa = L.x, *xb = L.x2;
        for (int s = 0; s < sweeps; s++) {
            const float omega = (l == last) ? p.omegaCoarse : p.om[s];
            for (int c = tid; c < L.n; c += nt) {
                int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
                xb[c] = mg_sweep_cell(L, L, xa, nullptr, omega, scale, s == 0 ? 0 : 1, c, i, j, k);
            }
            __syncthreads();
            float *t = xa; xa = xb; xb = t;
        }
        // after an odd number of sweeps the result sits in x2: copy so that x always holds it
        if (sweeps & 1) {
            for (int c = tid; c < L.n; c += nt) L.x[c] = L.x2[c];
            __syncthreads();
        }
        if (l < last) {
            const MgLevel &C = lv[l + 1];
            for (int c = tid; c < C.n; c += nt) {
                int i = c % C.I, j = (c / C.I) % C.J, k = c / C.sk;
                C.b[c] = (C.invD[c] == 0.0f) ? 0.0f : mg_restrict_cell(L, L.x, i, j, k);
            }
            __syncthreads();
        }
    }
    for (int l = last - 1; l >= first; l--) {
        const MgLevel &L = lv[l];
        const MgLevel &C = lv[l + 1];
        float *xa = L.x, *xb = L.x2;
        for (int s = 0; s < p.nu; s++) {
            const float omega = p.om[p.nu - 1 - s];
            for (int c = tid; c < L.n; c += nt) {
                int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
                xb[c] = mg_sweep_cell(L, C, xa, C.x, omega, scale, s == 0 ? 2 : 1, c, i, j, k);
            }
            __syncthreads();
            float *t = xa; xa = xb; xb = t;
        }
        if (p.nu & 1) {
            for (int c = tid; c < L.n; c += nt) L.x[c] = L.x2[c];
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(1024) k_mg_small(MgSmallArgs A, const DeviceScalars *S) {
    if (S->pcgDone) return;
    mg_small_body(A.lv, A.first, A.last, A.p);
}

// ------------------------------------------------------------------------------------------------
// All coarse levels (1 .. coarsest) of the V-cycle in ONE cooperative kernel.
//
// A level >= 1 holds at most 1/8 of the rows of the level below, so its passes are pure latency: as separate
// launches the ~17 coarse passes of a cycle cost more than the level-0 passes that move all the data.  Here one
// CTA per SM walks them with grid-wide barriers in between (list-driven, same device functions as the
// per-pass kernels); the levels of <= MG_SMALL cells are run by block 0 alone out of SHARED memory, with the
// arrays zero-padded by one plane on both sides so that the 7-point sums need neither bounds checks nor
// a != 0 tests (couplings that would wrap around a row or plane are zero by construction: border cells are
// never pressure rows) -- twelve independent loads per cell instead of six dependent pairs.
// ------------------------------------------------------------------------------------------------
struct SmLevel {
    int I, J, K, sj, sk, n;
    float *diag, *invD, *b;             // [n]
    float *oU, *oV, *oW, *x, *x2;       // [n], zero-padded by `sm_pad` entries on both sides
};
__host__ __device__ inline int sm_pad(int sk) { return (sk + 1 + 3) & ~3; }
__host__ __device__ inline size_t sm_level_floats(int n, int sk) { return 3 * (size_t)n + 5 * ((size_t)n + 2 * sm_pad(sk)); }

__device__ __forceinline__ float sm_offsum(const SmLevel &L, const float *x, int c) {
    float s = 0.0f;
    s += L.oW[c - L.sk] * x[c - L.sk];
    s += L.oV[c - L.sj] * x[c - L.sj];
    s += L.oU[c - 1] * x[c - 1];
    s += L.oU[c] * x[c + 1];
    s += L.oV[c] * x[c + L.sj];
    s += L.oW[c] * x[c + L.sk];
    return s;
}
__device__ __forceinline__ float sm_residual(const SmLevel &L, const float *x, int c) {
    return (L.invD[c] == 0.0f) ? 0.0f : L.b[c] - (L.diag[c] * x[c] - sm_offsum(L, x, c));
}

// ---- the shared-memory levels.  The passes are written once for a generic thread group (tid, nt, sync): the whole
// CTA with __syncthreads for the levels of more than SM_WARP_CELLS cells, ONE WARP with __syncwarp below that
// (a 1024-thread barrier per pass costs more than the pass itself on a level of 4^3 cells or fewer; measured on
// the B200: with the 8^3 level in one warp as well the levels take 38 us instead of 19 -- a single warp has nothing
// to hide the shared-memory latency of its sixteen cells per lane behind).
static constexpr int SM_WARP_CELLS = 64;

// nu pre-sweeps from a zero guess (or `sweeps` of them on the coarsest level); leaves the result in L.x
// (the descriptors live in shared memory: each pass copies what it uses into registers first -- through the reference
// every array pointer would be re-read from shared memory after every store, a chain of two dependent loads per access)
template <class Sync>
__device__ __forceinline__ void sm_presweeps(const SmLevel &Lref, int sweeps, const MgParams &p, bool coarsest, int tid, int nt, Sync sync) {
    const SmLevel L = Lref;
    float *xa = L.x, *xb = L.x2;
    for (int s = 0; s < sweeps; s++) {
        const float omega = coarsest ? p.omegaCoarse : p.om[s];
        for (int c = tid; c < L.n; c += nt) {
            const float inv = L.invD[c];
            float v = 0.0f;
            if (inv != 0.0f) {
                if (s == 0) v = omega * inv * L.b[c];
                else v = (1.0f - omega) * xa[c] + omega * inv * (L.b[c] + sm_offsum(L, xa, c));
            }
            xb[c] = v;
        }
        sync();
        float *t = xa; xa = xb; xb = t;
    }
    if (sweeps & 1) {      // after an odd number of sweeps the result sits in x2: copy so that x always holds it
        for (int c = tid; c < L.n; c += nt) L.x[c] = L.x2[c];
        sync();
    }
}
// b_C = P^T (b_L - A_L x_L)
template <class Sync>
__device__ __forceinline__ void sm_restrict(const SmLevel &Lref, const SmLevel &Cref, int tid, int nt, Sync sync) {
    const SmLevel L = Lref, C = Cref;
    for (int cc = tid; cc < C.n; cc += nt) {
        float acc = 0.0f;
        if (C.invD[cc] != 0.0f) {
            const int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int i = 2 * ci + (q & 1), j = 2 * cj + ((q >> 1) & 1), k = 2 * ck + (q >> 2);
                if (i < L.I && j < L.J && k < L.K) acc += sm_residual(L, L.x, i + L.sj * j + L.sk * k);
            }
        }
        C.b[cc] = acc;
    }
    sync();
}
// nu post-sweeps on (x + scale * P e), e = the solution of the next coarser level; leaves the result in L.x
template <class Sync>
__device__ __forceinline__ void sm_postsweeps(const SmLevel &Lref, const SmLevel &Cref, const MgParams &p, int tid, int nt, Sync sync) {
    const SmLevel L = Lref, C = Cref;
    const float scale = p.scale;
    const float *e = C.x;
    float *xa = L.x, *xb = L.x2;
    for (int s = 0; s < p.nu; s++) {
        const float omega = p.om[p.nu - 1 - s];
        for (int c = tid; c < L.n; c += nt) {
            const float inv = L.invD[c];
            float v = 0.0f;
            if (inv != 0.0f) {
                if (s == 0) {
                    // sweep on (x + scale * P e): parents of the cell and of its six neighbours
                    const int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
                    const int pi = i >> 1, pj = C.sj * (j >> 1), pk = C.sk * (k >> 1);
                    const float xc = xa[c] + scale * e[pi + pj + pk];
                    float ns = 0.0f;
                    ns += L.oW[c - L.sk] * (xa[c - L.sk] + scale * e[pi + pj + C.sk * ((k - 1) >> 1)]);
                    ns += L.oV[c - L.sj] * (xa[c - L.sj] + scale * e[pi + C.sj * ((j - 1) >> 1) + pk]);
                    ns += L.oU[c - 1] * (xa[c - 1] + scale * e[((i - 1) >> 1) + pj + pk]);
                    ns += L.oU[c] * (xa[c + 1] + scale * e[((i + 1) >> 1) + pj + pk]);
                    ns += L.oV[c] * (xa[c + L.sj] + scale * e[pi + C.sj * ((j + 1) >> 1) + pk]);
                    ns += L.oW[c] * (xa[c + L.sk] + scale * e[pi + pj + C.sk * ((k + 1) >> 1)]);
                    v = (1.0f - omega) * xc + omega * inv * (L.b[c] + ns);
                } else {
                    v = (1.0f - omega) * xa[c] + omega * inv * (L.b[c] + sm_offsum(L, xa, c));
                }
            }
            xb[c] = v;
        }
        sync();
        float *t = xa; xa = xb; xb = t;
    }
    if (p.nu & 1) {
        for (int c = tid; c < L.n; c += nt) L.x[c] = L.x2[c];
        sync();
    }
}
// down, coarsest sweeps, up over the levels first..last with one thread group; leaves the result in sl[first].x
template <class Sync>
__device__ __forceinline__ void sm_cycle(const SmLevel *sl, int first, int last, const MgParams &p, int tid, int nt, Sync sync) {
    for (int l = first; l <= last; l++) {
        sm_presweeps(sl[l], (l == last) ? p.coarseSweeps : p.nu, p, l == last, tid, nt, sync);
        if (l < last) sm_restrict(sl[l], sl[l + 1], tid, nt, sync);
    }
    for (int l = last - 1; l >= first; l--) sm_postsweeps(sl[l], sl[l + 1], p, tid, nt, sync);
}

// ---- the levels of at most SM_WARP_CELLS (64) cells: a direct solve.  Walking them with sweeps in one warp costs
// ~12 us per V-cycle (thirteen dependent passes over <= 64 cells, measured with FLIP_MG_TRACE); their operator is fixed
// for the whole solve, so its dense inverse is computed once per solve (Gauss-Jordan in shared memory, no pivoting: the
// Galerkin operators are symmetric positive definite M-matrices unless the liquid has no free surface at all, in
// which case the flag stays 0 and the sweeps are used) and a V-cycle applies it as a 64 x 64 product.  An exact
// coarse solve keeps the preconditioner symmetric positive definite.
__global__ void __launch_bounds__(1024) k_mg_invert_small(MgLevel L, float *__restrict__ inv, int *__restrict__ ok) {
    constexpr int N = SM_WARP_CELLS;              // the system is padded to N unknowns (identity rows): shifts, no divisions
    __shared__ float A[N][2 * N + 1];
    __shared__ float col[N];
    __shared__ float piv;
    __shared__ int bad;
    const int n = L.n, tid = threadIdx.x, nt = blockDim.x;
    if (n > N) { if (tid == 0) *ok = 0; return; }
    for (int q = tid; q < N * 2 * N; q += nt) { const int r = q >> 7, c = q & (2 * N - 1); A[r][c] = (c == r + N || c == r) ? 1.0f : 0.0f; }
    if (tid == 0) bad = 0;
    __syncthreads();
    float dloc = 0.0f;
    for (int c = tid; c < n; c += nt) {
        const int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
        if (L.invD[c] == 0.0f) continue;          // an inactive cell keeps x = 0 (its right-hand side is 0)
        A[c][c] = L.diag[c];
        // couplings are stored with the lower cell of the pair; a coupling to an inactive cell is 0 by construction
        if (i + 1 < L.I) { const float o = L.oU[c]; A[c][c + 1] = -o; A[c + 1][c] = -o; }
        if (j + 1 < L.J) { const float o = L.oV[c]; A[c][c + L.sj] = -o; A[c + L.sj][c] = -o; }
        if (k + 1 < L.K) { const float o = L.oW[c]; A[c][c + L.sk] = -o; A[c + L.sk][c] = -o; }
        dloc = fmaxf(dloc, fabsf(L.diag[c]));
    }
    __syncthreads();
    float dmax = 1.0f;
    for (int c = 0; c < n; c++) dmax = fmaxf(dmax, fabsf(A[c][c]));
    (void)dloc;
    for (int p = 0; p < N; p++) {
        if (tid == 0) { piv = A[p][p]; if (!(piv > 1.0e-6f * dmax)) bad = 1; }
        if (tid < N) col[tid] = A[tid][p];          // column p before this step touches it
        __syncthreads();
        if (bad) break;
        const float ip = 1.0f / piv;
        // rows r != p: A[r] -= col[r] * (A[p] / piv); row p: A[p] /= piv   (row p is read unscaled, every element once)
        for (int q = tid; q < N * 2 * N; q += nt) {
            const int r = q >> 7, c = q & (2 * N - 1);
            const float ap = A[p][c] * ip;
            if (r != p) A[r][c] -= col[r] * ap;
        }
        __syncthreads();
        for (int c = tid; c < 2 * N; c += nt) A[p][c] *= ip;
        __syncthreads();
    }
    __syncthreads();
    if (bad) { if (tid == 0) *ok = 0; return; }
    for (int q = tid; q < N * N; q += nt) {
        const int r = q >> 6, c = q & (N - 1);
        const bool act = r < n && c < n && L.invD[r] != 0.0f && L.invD[c] != 0.0f;
        inv[r * N + c] = act ? A[r][N + c] : 0.0f;
    }
    if (tid == 0) *ok = 1;
}

// x = inv * b on the level, by the whole CTA (sixteen threads per row, four columns each; rows >= n idle)
__device__ __forceinline__ void sm_direct_solve(const SmLevel &Lref, const float *__restrict__ inv, int tid) {
    const SmLevel L = Lref;
    const int row = tid >> 4, part = tid & 15;
    float acc = 0.0f;
    if (row < L.n) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(inv + row * SM_WARP_CELLS + part * 4));
        const int c = part * 4;
        acc = a.x * (c < L.n ? L.b[c] : 0.0f) + a.y * (c + 1 < L.n ? L.b[c + 1] : 0.0f) + a.z * (c + 2 < L.n ? L.b[c + 2] : 0.0f) +
              a.w * (c + 3 < L.n ? L.b[c + 3] : 0.0f);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (row < L.n && part == 0) L.x[row] = acc;
}

// the shared-memory levels first..last of the V-cycle, called by every thread of ONE CTA
// (tr: developer probe, a globaltimer stamp after every CTA-wide pass)
__device__ __forceinline__ void sm_small_body(const SmLevel *sl, int first, int last, const MgParams &p,
                                              unsigned long long *tr = nullptr, int *ntr = nullptr,
                                              const float *smallInv = nullptr) {
    const int tid = threadIdx.x, nt = blockDim.x;
    auto bsync = [&] {
        __syncthreads();
        if (tr && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            tr[(*ntr)++] = t;
        }
    };
    auto wsync = [] { __syncwarp(); };
    int lw = first;     // first level small enough for one warp
    while (lw <= last && sl[lw].n > SM_WARP_CELLS) lw++;
    if (lw > last) { sm_cycle(sl, first, last, p, tid, nt, bsync); return; }
    for (int l = first; l < lw; l++) {
        sm_presweeps(sl[l], p.nu, p, false, tid, nt, bsync);
        sm_restrict(sl[l], sl[l + 1], tid, nt, bsync);
    }
    if (smallInv) sm_direct_solve(sl[lw], smallInv, tid);       // (block-uniform)
    else if (tid < 32) sm_cycle(sl, lw, last, p, tid, 32, wsync);
    __syncthreads();
    for (int l = lw - 1; l >= first; l--) sm_postsweeps(sl[l], sl[l + 1], p, tid, nt, bsync);
}

// ---- branch-free passes on the zero-padded global arrays of the levels >= 1 (same arithmetic as mg_sweep_cell /
// mg_residual_cell; a zero coupling multiplies a finite neighbour value instead of skipping the load)
__device__ __forceinline__ float mgp_offsum(const MgLevel &L, const float *x, int c) {
    const float a0 = L.oW[c - L.sk], a1 = L.oV[c - L.sj], a2 = L.oU[c - 1], a3 = L.oU[c], a4 = L.oV[c], a5 = L.oW[c];
    const float x0 = x[c - L.sk], x1 = x[c - L.sj], x2 = x[c - 1], x3 = x[c + 1], x4 = x[c + L.sj], x5 = x[c + L.sk];
    float s = 0.0f;
    s += a0 * x0; s += a1 * x1; s += a2 * x2; s += a3 * x3; s += a4 * x4; s += a5 * x5;
    return s;
}
__device__ __forceinline__ float mgp_residual(const MgLevel &L, const float *x, int c) {
    const float inv = L.invD[c], b = L.b[c], d = L.diag[c], xc = x[c];
    const float ns = mgp_offsum(L, x, c);
    return (inv == 0.0f) ? 0.0f : b - (d * xc - ns);
}
__device__ __forceinline__ float mgp_sweep(const MgLevel &L, const MgLevel &C, const float *xin, const float *e, float omega,
                                           float scale, int mode, int c, float omega0) {
    const float inv = L.invD[c], b = L.b[c];
    float xc, ns;
    if (mode == 0) {
        return (inv == 0.0f) ? 0.0f : omega * inv * b;
    } else if (mode == 3) {
        const int sj = L.sj, sk = L.sk;
        const float a0 = L.oW[c - sk], a1 = L.oV[c - sj], a2 = L.oU[c - 1], a3 = L.oU[c], a4 = L.oV[c], a5 = L.oW[c];
        const float d0 = L.invD[c - sk], d1 = L.invD[c - sj], d2 = L.invD[c - 1], d3 = L.invD[c + 1], d4 = L.invD[c + sj], d5 = L.invD[c + sk];
        const float b0 = L.b[c - sk], b1 = L.b[c - sj], b2 = L.b[c - 1], b3 = L.b[c + 1], b4 = L.b[c + sj], b5 = L.b[c + sk];
        xc = omega0 * inv * b;
        ns = 0.0f;
        ns += a0 * (omega0 * d0 * b0); ns += a1 * (omega0 * d1 * b1); ns += a2 * (omega0 * d2 * b2);
        ns += a3 * (omega0 * d3 * b3); ns += a4 * (omega0 * d4 * b4); ns += a5 * (omega0 * d5 * b5);
    } else if (mode == 1) {
        xc = xin[c];
        ns = mgp_offsum(L, xin, c);
    } else {
        const int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
        const int pi = i >> 1, pj = C.sj * (j >> 1), pk = C.sk * (k >> 1);
        const float a0 = L.oW[c - L.sk], a1 = L.oV[c - L.sj], a2 = L.oU[c - 1], a3 = L.oU[c], a4 = L.oV[c], a5 = L.oW[c];
        const float x0 = xin[c - L.sk], x1 = xin[c - L.sj], x2 = xin[c - 1], x3 = xin[c + 1], x4 = xin[c + L.sj], x5 = xin[c + L.sk];
        const float e0 = e[pi + pj + C.sk * ((k - 1) >> 1)], e1 = e[pi + C.sj * ((j - 1) >> 1) + pk], e2 = e[((i - 1) >> 1) + pj + pk];
        const float e3 = e[((i + 1) >> 1) + pj + pk], e4 = e[pi + C.sj * ((j + 1) >> 1) + pk], e5 = e[pi + pj + C.sk * ((k + 1) >> 1)];
        xc = xin[c] + scale * e[pi + pj + pk];
        ns = 0.0f;
        ns += a0 * (x0 + scale * e0); ns += a1 * (x1 + scale * e1); ns += a2 * (x2 + scale * e2);
        ns += a3 * (x3 + scale * e3); ns += a4 * (x4 + scale * e4); ns += a5 * (x5 + scale * e5);
    }
    return (inv == 0.0f) ? 0.0f : (1.0f - omega) * xc + omega * inv * (b + ns);
}

struct MgCoarseArgs {
    MgLevel lv[MG_MAX_LEVELS];
    const int *seg[MG_MAX_LEVELS];   // active-segment lists of levels 1..fs
    const int *segCount;             // [MG_MAX_LEVELS]
    int fs, last;                    // levels fs..last: block 0, shared memory
    MgParams p;
    int smemFloats;
    int group;                       // CTAs that run the levels 2..fs-1 among themselves (group barrier)
    unsigned int *groupBar;          // arrival counter of that barrier (zero between launches)
    unsigned long long *trace;       // developer probe (FLIP_MG_TRACE): globaltimer at every phase boundary, block 0
    const float *smallInv;           // dense inverse of the first level of at most SM_WARP_CELLS cells (k_mg_invert_small)
    const int *smallInvOk;           // ... usable (the level's operator was not singular)
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Level 1 holds ~1/8 of the rows and keeps all 148 CTAs busy: its passes are separated by grid-wide barriers.
// From level 2 on a pass is a few thousand cells: a grid-wide barrier (148 CTAs, ~3 us) costs far more than the
// pass, so those levels are run by the first `group` CTAs only, separated by a barrier among just those CTAs (one
// atomic arrival + an acquire spin per CTA, < 1 us); the other CTAs go straight to the grid barrier in front of the
// up-sweeps of level 1.  Every CTA executes the same number of grid barriers.
__global__ void __launch_bounds__(1024) k_mg_coarse(MgCoarseArgs A, const DeviceScalars *S) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float smf[];
    __shared__ SmLevel sl[MG_MAX_LEVELS];
    if (S->pcgDone) return;      // grid-uniform: written by the previous launch only
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int wpb = nt >> 5;
    const int nu = A.p.nu;
    const float scale = A.p.scale;
    const int G = min(max(A.group, 1), (int)gridDim.x);
    const bool inGroup = (int)blockIdx.x < G;
    unsigned int gen = 0;        // group barriers passed so far (uniform over the group)
    auto group_sync = [&]() {
        gen++;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(A.groupBar, 1u);
            const unsigned int target = gen * (unsigned int)G;
            while (ld_acquire_gpu_u32(A.groupBar) < target) {}
            __threadfence();
        }
        __syncthreads();
    };
    int ntrace = 0;
    auto stamp = [&]() {
        if (A.trace && blockIdx.x == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            A.trace[ntrace++] = t;
        }
    };
    stamp();

    // block 0 stages the operators of its levels while the other blocks start on level 1
    if (blockIdx.x == 0) {
        if (tid == 0) {
            float *q = smf;
            for (int l = A.fs; l <= A.last; l++) {
                const MgLevel &Gl = A.lv[l];
                SmLevel L;
                L.I = Gl.I; L.J = Gl.J; L.K = Gl.K; L.sj = Gl.sj; L.sk = Gl.sk; L.n = Gl.n;
                const int pad = sm_pad(Gl.sk);
                L.diag = q; q += Gl.n; L.invD = q; q += Gl.n; L.b = q; q += Gl.n;
                L.oU = q + pad; q += Gl.n + 2 * pad; L.oV = q + pad; q += Gl.n + 2 * pad; L.oW = q + pad; q += Gl.n + 2 * pad;
                L.x = q + pad; q += Gl.n + 2 * pad; L.x2 = q + pad; q += Gl.n + 2 * pad;
                sl[l] = L;
            }
        }
        for (int q = tid; q < A.smemFloats; q += nt) smf[q] = 0.0f;
        __syncthreads();
    }
    stamp();
    // second half of the staging (the operator arrays): block 0 runs it during the down passes of level 1,
    // in which it takes no part
    auto stage_operators = [&]() {
        for (int l = A.fs; l <= A.last; l++) {
            const MgLevel &Gl = A.lv[l];
            const SmLevel &L = sl[l];
            // all loads of four strides first (read-only path): through generic pointers the compiler must
            // otherwise order every shared store before the next global load, ~20 dependent L2 latencies
            for (int c0 = tid; c0 < Gl.n; c0 += 4 * nt) {
                float v[4][5];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int c = c0 + u * nt;
                    if (c < Gl.n) {
                        v[u][0] = __ldg(Gl.diag + c); v[u][1] = __ldg(Gl.invD + c); v[u][2] = __ldg(Gl.oU + c);
                        v[u][3] = __ldg(Gl.oV + c); v[u][4] = __ldg(Gl.oW + c);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int c = c0 + u * nt;
                    if (c < Gl.n) { L.diag[c] = v[u][0]; L.invD[c] = v[u][1]; L.oU[c] = v[u][2]; L.oV[c] = v[u][3]; L.oW[c] = v[u][4]; }
                }
            }
        }
        __syncthreads();
    };
    bool staged = false;

    // one pre-smoothing pass + the restriction of level l over the warps gw, gw + nw, ... / the CTAs rb, rb + rnb, ...
    // (an index of 0x3fffffff: this CTA takes no part); `sync` separates the passes
    const float *lx[MG_MAX_LEVELS];
    auto down_level = [&](int l, int gw, int nw, int rb, int rnb, auto &&sync) {
        const MgLevel &L = A.lv[l];
        const int *__restrict__ list = A.seg[l];
        const int nseg = A.segCount[l];
        float *xa = L.x, *xb = L.x2;
        int sw = 0;
        if (nu >= 2) {
            for (int s = gw; s < nseg; s += nw) {
                const int c = list[s] + lane;
                if (c < L.n) xb[c] = mgp_sweep(L, L, xa, nullptr, A.p.om[1], scale, 3, c, A.p.om[0]);
            }
            sync();
            float *t = xa; xa = xb; xb = t;
            sw = 2;
        }
        for (; sw < nu; sw++) {
            for (int s = gw; s < nseg; s += nw) {
                const int c = list[s] + lane;
                if (c < L.n) xb[c] = mgp_sweep(L, L, xa, nullptr, A.p.om[sw], scale, sw == 0 ? 0 : 1, c, 0.0f);
            }
            sync();
            float *t = xa; xa = xb; xb = t;
        }
        lx[l] = xa;
        if (blockIdx.x == 0 && !staged && l == 1) { stage_operators(); staged = true; }
        // restriction into level l+1: eight threads per coarse cell, four coarse segments per CTA pass
        {
            const MgLevel &C = A.lv[l + 1];
            const int *__restrict__ listC = A.seg[l + 1];
            const int nsegC = A.segCount[l + 1];
            const int t = tid & 255, q = t & 7;
            for (long long s = (rb >= 0x3fffffff ? (long long)nsegC : (long long)rb * (nt >> 8) + (tid >> 8)); s < nsegC;
                 s += (long long)rnb * (nt >> 8)) {
                const int cc = listC[s] + (t >> 3);
                float v = 0.0f;
                const bool act = (cc < C.n) && C.invD[cc] != 0.0f;
                if (act) {
                    const int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
                    const int i = 2 * ci + (q & 1), j = 2 * cj + ((q >> 1) & 1), k = 2 * ck + (q >> 2);
                    if (i < L.I && j < L.J && k < L.K) v = mgp_residual(L, xa, i + L.sj * j + L.sk * k);
                }
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (q == 0 && cc < C.n) C.b[cc] = act ? v : 0.0f;
            }
        }
        sync();
    };
    // the post-smoothing passes of level l (the first one on x + scale * P e); lastSync: also after the last pass
    auto up_level = [&](int l, int gw, int nw, bool lastSync, auto &&sync) {
        const MgLevel &L = A.lv[l];
        const MgLevel &C = A.lv[l + 1];
        const int *__restrict__ list = A.seg[l];
        const int nseg = A.segCount[l];
        float *xa = const_cast<float *>(lx[l]);
        float *xb = (xa == L.x) ? L.x2 : L.x;
        const float *e = lx[l + 1];
        for (int sw = 0; sw < nu; sw++) {
            const float omega = A.p.om[nu - 1 - sw];
            for (int s = gw; s < nseg; s += nw) {
                const int c = list[s] + lane;
                if (c < L.n) xb[c] = mgp_sweep(L, C, xa, e, omega, scale, sw == 0 ? 2 : 1, c, 0.0f);
            }
            if (sw + 1 < nu || lastSync) sync();
            float *t = xa; xa = xb; xb = t;
        }
        lx[l] = xa;
    };
    auto gsync = [&]() { grid.sync(); stamp(); };
    auto psync = [&]() { group_sync(); stamp(); };

    // ---- level 1 down (all CTAs; block 0 has just spent ~10 us staging and stages the operators meanwhile: the
    // other blocks share the passes among themselves, so that the staging is off the critical path)
    if (A.fs > 1) {
        const bool skip0 = gridDim.x > 1;
        const int gw = skip0 ? (blockIdx.x == 0 ? 0x3fffffff : ((blockIdx.x - 1) * nt + tid) >> 5) : ((blockIdx.x * nt + tid) >> 5);
        const int nw = skip0 ? (gridDim.x - 1) * wpb : gridDim.x * wpb;
        const int rb = skip0 ? (blockIdx.x == 0 ? 0x3fffffff : (int)blockIdx.x - 1) : (int)blockIdx.x;
        const int rnb = skip0 ? (int)gridDim.x - 1 : (int)gridDim.x;
        down_level(1, gw, nw, rb, rnb, gsync);
    }
    // ---- levels 2 .. coarsest and back up to level 2: the group
    if (inGroup) {
        const int gwG = (blockIdx.x * nt + tid) >> 5, nwG = G * wpb;
        for (int l = 2; l < A.fs; l++) down_level(l, gwG, nwG, (int)blockIdx.x, G, psync);
        if (blockIdx.x == 0) {
            if (!staged) { stage_operators(); staged = true; }
            const MgLevel &Gl = A.lv[A.fs];
            const SmLevel &L = sl[A.fs];
            for (int c = tid; c < Gl.n; c += nt) L.b[c] = Gl.b[c];
            __syncthreads();
            stamp();
            sm_small_body(sl, A.fs, A.last, A.p, A.trace, &ntrace, (A.smallInv && *A.smallInvOk) ? A.smallInv : nullptr);
            stamp();
            for (int c = tid; c < Gl.n; c += nt) Gl.x[c] = L.x[c];
        }
        lx[A.fs] = A.lv[A.fs].x;
        if (A.fs > 2) {
            psync();      // the solution of level fs is visible to the group
            for (int l = A.fs - 1; l >= 2; l--) up_level(l, gwG, nwG, l > 2, psync);
        }
    }
    lx[A.fs] = A.lv[A.fs].x;
    // where the group left the results of the levels it ran (parity of its ping-pong, the same on every CTA)
    {
        const int swaps = (nu >= 2 ? nu - 1 : nu) + nu;
        for (int l = 2; l < A.fs; l++) lx[l] = (swaps & 1) ? A.lv[l].x2 : A.lv[l].x;
    }
    if (A.fs > 1) {
        gsync();          // join: what level 1 prolongates from is complete
        if (blockIdx.x == 0 && tid == 0) *A.groupBar = 0u;     // every group barrier of this launch has been passed
        up_level(1, (blockIdx.x * nt + tid) >> 5, gridDim.x * wpb, false, gsync);
    }
}

// (Tried: a contiguous chunk of segments per block for L1 reuse of the j neighbours -- no faster, less balanced.)
// Walk of a warp over the active segments w0, w0 + nw, w0 + 2 nw, ... with the descriptors (first cell, row mask) of
// the next 32 trips fetched at once, one per lane, and handed out by shuffle.  Body variables: segc, segm.
// (`continue` inside the body goes on to the next trip; every lane reaches the shuffles of every trip.)
#define SEG_WALK_BEGIN(segCell, segMask, nseg, nw, lane)                                                     \
    for (int segBase_ = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; segBase_ < (nseg); segBase_ += 32 * (nw)) { \
        const long long segMine_ = (long long)segBase_ + (long long)(lane) * (nw);                            \
        int segCellMine_ = 0;                                                                                 \
        unsigned int segMaskMine_ = 0u;                                                                       \
        if (segMine_ < (nseg)) { segCellMine_ = (segCell)[segMine_]; segMaskMine_ = (segMask)[segMine_]; }    \
        const int segTrips_ = min(32, ((nseg) - segBase_ + (nw) - 1) / (nw));                                 \
        for (int segTrip_ = 0; segTrip_ < segTrips_; segTrip_++) {                                            \
            const int segc = __shfl_sync(0xffffffffu, segCellMine_, segTrip_);                                \
            const unsigned int segm = __shfl_sync(0xffffffffu, segMaskMine_, segTrip_);
#define SEG_WALK_END \
        }            \
    }

// ---- level 0: segment-driven, right-hand side = the fp64 PCG residual
struct Mg0 {
    PGrid g;
    float fac;                      // (float)(dt/dx^2): level-0 off-diagonal magnitude = fac * weight
    const double *Adiag;
    const float *oU, *oV, *oW;      // masked face weights (AoffU/V/W)
    const float *invD;
    const unsigned int *rowBits;    // 1 bit per cell: is a pressure row
};

__device__ __forceinline__ float mg0_offsum(const Mg0 &M, const float *x, int c) {
    float a0 = M.oW[c - M.g.sk], a1 = M.oV[c - M.g.sj], a2 = M.oU[c - 1], a3 = M.oU[c], a4 = M.oV[c], a5 = M.oW[c];
    float x0 = x[c - M.g.sk], x1 = x[c - M.g.sj], x2 = x[c - 1], x3 = x[c + 1], x4 = x[c + M.g.sj], x5 = x[c + M.g.sk];
    float s = 0.0f;
    s += (a0 != 0.0f) ? a0 * x0 : 0.0f;
    s += (a1 != 0.0f) ? a1 * x1 : 0.0f;
    s += (a2 != 0.0f) ? a2 * x2 : 0.0f;
    s += (a3 != 0.0f) ? a3 * x3 : 0.0f;
    s += (a4 != 0.0f) ? a4 * x4 : 0.0f;
    s += (a5 != 0.0f) ? a5 * x5 : 0.0f;
    return M.fac * s;
}

__device__ __forceinline__ float mg0_corr(const Mg0 &M, const MgLevel &C, const float *x, const float *e, float scale, int c) {
    int i = c % M.g.I, j = (c / M.g.I) % M.g.J, k = c / M.g.sk;
    return x[c] + scale * e[(i >> 1) + C.sj * (j >> 1) + C.sk * (k >> 1)];
}

__device__ __forceinline__ float mg0_offsum_corr(const Mg0 &M, const MgLevel &C, const float *x, const float *e,
                                                 float scale, int c) {
    int i = c % M.g.I, j = (c / M.g.I) % M.g.J, k = c / M.g.sk;
    float a0 = M.oW[c - M.g.sk], a1 = M.oV[c - M.g.sj], a2 = M.oU[c - 1], a3 = M.oU[c], a4 = M.oV[c], a5 = M.oW[c];
    float x0 = x[c - M.g.sk], x1 = x[c - M.g.sj], x2 = x[c - 1], x3 = x[c + 1], x4 = x[c + M.g.sj], x5 = x[c + M.g.sk];
    // parents of the six neighbours (rows are interior: i,j,k >= 1 and +1 stays inside the grid)
    int pi = i >> 1, pj = C.sj * (j >> 1), pk = C.sk * (k >> 1);
    float e0 = e[pi + pj + C.sk * ((k - 1) >> 1)], e1 = e[pi + C.sj * ((j - 1) >> 1) + pk], e2 = e[((i - 1) >> 1) + pj + pk];
    float e3 = e[((i + 1) >> 1) + pj + pk], e4 = e[pi + C.sj * ((j + 1) >> 1) + pk], e5 = e[pi + pj + C.sk * ((k + 1) >> 1)];
    // rows of a neighbouring z-slab (beyond the cut) take no part in the block-local correction
    const bool own0 = !M.g.cut || (k - 1 >= M.g.kOwn0), own5 = !M.g.cut || (k + 1 < M.g.kOwn1);
    float s = 0.0f;
    s += (a0 != 0.0f && own0) ? a0 * (x0 + scale * e0) : 0.0f;
    s += (a1 != 0.0f) ? a1 * (x1 + scale * e1) : 0.0f;
    s += (a2 != 0.0f) ? a2 * (x2 + scale * e2) : 0.0f;
    s += (a3 != 0.0f) ? a3 * (x3 + scale * e3) : 0.0f;
    s += (a4 != 0.0f) ? a4 * (x4 + scale * e4) : 0.0f;
    s += (a5 != 0.0f && own5) ? a5 * (x5 + scale * e5) : 0.0f;
    return M.fac * s;
}

// level-0 analogue of mg_offsum_first: the first sweep from a zero guess is omega * invD * (float)r, pointwise
__device__ __forceinline__ float mg0_offsum_first(const Mg0 &M, const double *__restrict__ r, int c, float omega) {
    const int sj = M.g.sj, sk = M.g.sk;
    float a0 = M.oW[c - sk], a1 = M.oV[c - sj], a2 = M.oU[c - 1], a3 = M.oU[c], a4 = M.oV[c], a5 = M.oW[c];
    double r0 = r[c - sk], r1 = r[c - sj], r2 = r[c - 1], r3 = r[c + 1], r4 = r[c + sj], r5 = r[c + sk];
    float d0 = M.invD[c - sk], d1 = M.invD[c - sj], d2 = M.invD[c - 1], d3 = M.invD[c + 1], d4 = M.invD[c + sj], d5 = M.invD[c + sk];
    float s = 0.0f;
    s += (a0 != 0.0f && d0 != 0.0f) ? a0 * (omega * d0 * (float)r0) : 0.0f;
    s += (a1 != 0.0f && d1 != 0.0f) ? a1 * (omega * d1 * (float)r1) : 0.0f;
    s += (a2 != 0.0f && d2 != 0.0f) ? a2 * (omega * d2 * (float)r2) : 0.0f;
    s += (a3 != 0.0f && d3 != 0.0f) ? a3 * (omega * d3 * (float)r3) : 0.0f;
    s += (a4 != 0.0f && d4 != 0.0f) ? a4 * (omega * d4 * (float)r4) : 0.0f;
    s += (a5 != 0.0f && d5 != 0.0f) ? a5 * (omega * d5 * (float)r5) : 0.0f;
    return M.fac * s;
}

// mode as in mg_sweep_cell.  last != 0: also write z (fp64) and accumulate rho = z.r into slot `rhoSlot`.
template <int MODE>
__device__ __forceinline__ float mg0_sweep_row(const Mg0 &M, const MgLevel &C, const double *__restrict__ r, const float *xin,
                                               const float *e, float omega, float omega0, float scale, int c, double &rd) {
    const float inv = M.invD[c];
    rd = r[c];
    const float b = (float)rd;
    if (inv == 0.0f) return 0.0f;
    if (MODE == 0) return omega * inv * b;
    if (MODE == 3) return (1.0f - omega) * (omega0 * inv * b) + omega * inv * (b + mg0_offsum_first(M, r, c, omega0));
    if (MODE == 1) return (1.0f - omega) * xin[c] + omega * inv * (b + mg0_offsum(M, xin, c));
    return (1.0f - omega) * mg0_corr(M, C, xin, e, scale, c) + omega * inv * (b + mg0_offsum_corr(M, C, xin, e, scale, c));
}

template <int MODE>
__global__ void __launch_bounds__(TPB, 6) k_mg0_sweep(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                                                      Mg0 M, MgLevel C, const double *__restrict__ r, const float *xin,
                                                      const float *e, float *xout, float omega, float scale, int last,
                                                      double *zout, DeviceScalars *S, int rhoSlot, float omega0) {
    if (S->pcgDone) return;
    const int lane = threadIdx.x & 31;
    const int nseg = S->numSegments, nw = (gridDim.x * blockDim.x) >> 5;
    double part = 0.0;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nseg; w += 2 * nw) {
        const int w2 = w + nw;
        const bool has2 = w2 < nseg;
        const int cA = segCell[w] + lane;
        const unsigned int mA = segMask[w];
        const int cB = has2 ? segCell[w2] + lane : cA;
        const unsigned int mB = has2 ? segMask[w2] : 0u;
        const bool rA = (mA >> lane) & 1u, rB = (mB >> lane) & 1u;
        float xA = 0.0f, xB = 0.0f;
        double rdA = 0.0, rdB = 0.0;
        if (rA) xA = mg0_sweep_row<MODE>(M, C, r, xin, e, omega, omega0, scale, cA, rdA);
        if (rB) xB = mg0_sweep_row<MODE>(M, C, r, xin, e, omega, omega0, scale, cB, rdB);
        if (rA) {
            xout[cA] = xA;
            if (last) { zout[cA] = (double)xA; part += (double)xA * rdA; }
        }
        if (rB) {
            xout[cB] = xB;
            if (last) { zout[cB] = (double)xB; part += (double)xB * rdB; }
        }
    }
    if (last) block_add(part, &S->rho[rhoSlot]);
}

// ---- fused passes of the single-GPU multigrid PCG (two level-0 passes fewer per iteration) ----------------------
// (1) k_pcg_update + the fused double pre-sweep (mode 3) in one pass.  The pre-sweep needs the NEW residual at the
// six neighbours; it is recomputed there as r - alpha*q instead of being read back, and written to the other buffer
// of a ping-pong pair (rOut) because neighbouring threads still read the old one.
//     alpha = rho/(s.q);  x += alpha s;  r' = r - alpha q;  ||r'||_inf;  xout = two damped-Jacobi sweeps on r' from zero
__global__ void __launch_bounds__(TPB, 5) k_pcg_update_presweep(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                                                                Mg0 M, const double *__restrict__ s, const double *__restrict__ q,
                                                                double *__restrict__ x, const double *__restrict__ r,
                                                                double *__restrict__ rOut, float *__restrict__ xout, float omega,
                                                                float omega0, DeviceScalars *S, int it) {
    if (S->pcgDone) return;
    const int lane = threadIdx.x & 31;
    const int nseg = S->numSegments, nw = (gridDim.x * blockDim.x) >> 5;
    const double alpha = S->rho[it % 3] / S->dotSZ[it % 3];
    const int sj = M.g.sj, sk = M.g.sk;
    double rabs = 0.0;
    SEG_WALK_BEGIN(segCell, segMask, nseg, nw, lane)
        const int c = segc + lane;
        if (!((segm >> lane) & 1u)) continue;
        // everything is loaded before anything is consumed (rows are interior cells: every index is in range)
        const float a0 = M.oW[c - sk], a1 = M.oV[c - sj], a2 = M.oU[c - 1], a3 = M.oU[c], a4 = M.oV[c], a5 = M.oW[c];
        const float d0 = M.invD[c - sk], d1 = M.invD[c - sj], d2 = M.invD[c - 1], d3 = M.invD[c + 1], d4 = M.invD[c + sj], d5 = M.invD[c + sk];
        const double r0 = r[c - sk], r1 = r[c - sj], r2 = r[c - 1], r3 = r[c + 1], r4 = r[c + sj], r5 = r[c + sk];
        const double q0 = q[c - sk], q1 = q[c - sj], q2 = q[c - 1], q3 = q[c + 1], q4 = q[c + sj], q5 = q[c + sk];
        const double rc = r[c], qc = q[c], xc = x[c], sc = s[c];
        const float inv = M.invD[c];
        const double rv = rc - alpha * qc;
        x[c] = xc + alpha * sc;
        rOut[c] = rv;
        rabs = fmax(rabs, fabs(rv));
        const float b = (float)rv;
        float ns = 0.0f;
        // a neighbour that is not a row has a zero weight (and possibly stale vector entries): select, never multiply
        ns += (a0 != 0.0f && d0 != 0.0f) ? a0 * (omega0 * d0 * (float)(r0 - alpha * q0)) : 0.0f;
        ns += (a1 != 0.0f && d1 != 0.0f) ? a1 * (omega0 * d1 * (float)(r1 - alpha * q1)) : 0.0f;
        ns += (a2 != 0.0f && d2 != 0.0f) ? a2 * (omega0 * d2 * (float)(r2 - alpha * q2)) : 0.0f;
        ns += (a3 != 0.0f && d3 != 0.0f) ? a3 * (omega0 * d3 * (float)(r3 - alpha * q3)) : 0.0f;
        ns += (a4 != 0.0f && d4 != 0.0f) ? a4 * (omega0 * d4 * (float)(r4 - alpha * q4)) : 0.0f;
        ns += (a5 != 0.0f && d5 != 0.0f) ? a5 * (omega0 * d5 * (float)(r5 - alpha * q5)) : 0.0f;
        xout[c] = (inv == 0.0f) ? 0.0f : (1.0f - omega) * (omega0 * inv * b) + omega * inv * (b + M.fac * ns);
    SEG_WALK_END
    block_max(rabs, &S->rMaxBits[it % 3]);
}

// (2) k_pcg_direction of iteration it-1 + k_pcg_spmv of iteration it in one pass: the new search vector is formed at
// the row AND at its six neighbours (s' = z + beta s), written to the other buffer of a ping-pong pair, and q = A s'
// follows at once.  it == 0: s' = z.  The convergence test of iteration it-1 (||r||_inf <= tol, pcgsolver.h:286-288)
// is taken here by every block; the last block to finish publishes it.
__global__ void __launch_bounds__(TPB, 5) k_pcg_dir_spmv(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                                                         PcgParams pp, const double *__restrict__ Adiag,
                                                         const float *__restrict__ AoffU, const float *__restrict__ AoffV,
                                                         const float *__restrict__ AoffW, const double *__restrict__ z,
                                                         const double *__restrict__ sOld, double *__restrict__ sNew,
                                                         double *__restrict__ q, DeviceScalars *S, int it) {
    if (S->pcgDone) return;
    double beta = 0.0, rmax = 0.0;
    bool converged = false, breakdown = false;
    if (it > 0) {
        rmax = __longlong_as_double((long long)S->rMaxBits[(it - 1) % 3]);
        const double rho = S->rho[(it - 1) % 3], rhoNew = S->rho[it % 3];
        converged = rmax <= S->pcgTol;
        breakdown = !converged && (rhoNew == 0.0 || rhoNew != rhoNew);
        beta = rhoNew / rho;
    }
    const int lane = threadIdx.x & 31;
    double part = 0.0;
    if (!converged && !breakdown) {
        const PGrid &g = pp.g;
        const int nseg = S->numSegments, nw = (gridDim.x * blockDim.x) >> 5;
        // A warp walks the segments w0, w0 + nw, ...: lane l fetches the descriptor of the l-th of the next 32 trips in
        // ONE coalesced-by-stride load, and every trip takes its descriptor by shuffle -- otherwise each trip is two
        // dependent round trips (descriptor, then rows), and with ~15 trips per warp that chain, not bandwidth, is
        // the kernel's time.
        SEG_WALK_BEGIN(segCell, segMask, nseg, nw, lane)
            const int c = segc + lane;
            if (!((segm >> lane) & 1u)) continue;
            const float a0 = AoffW[c - g.sk], a1 = AoffV[c - g.sj], a2 = AoffU[c - 1], a3 = AoffU[c], a4 = AoffV[c], a5 = AoffW[c];
            const double z0 = z[c - g.sk], z1 = z[c - g.sj], z2 = z[c - 1], z3 = z[c + 1], z4 = z[c + g.sj], z5 = z[c + g.sk];
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, sc = 0;
            if (it > 0) {
                s0 = sOld[c - g.sk]; s1 = sOld[c - g.sj]; s2 = sOld[c - 1]; s3 = sOld[c + 1]; s4 = sOld[c + g.sj]; s5 = sOld[c + g.sk];
                sc = sOld[c];
            }
            const double zc = z[c], dg = Adiag[c];
            const double vc = zc + beta * sc;
            double off = 0.0;
            off += (a0 != 0.0f) ? (double)a0 * (z0 + beta * s0) : 0.0;
            off += (a1 != 0.0f) ? (double)a1 * (z1 + beta * s1) : 0.0;
            off += (a2 != 0.0f) ? (double)a2 * (z2 + beta * s2) : 0.0;
            off += (a3 != 0.0f) ? (double)a3 * (z3 + beta * s3) : 0.0;
            off += (a4 != 0.0f) ? (double)a4 * (z4 + beta * s4) : 0.0;
            off += (a5 != 0.0f) ? (double)a5 * (z5 + beta * s5) : 0.0;
            const double qv = dg * vc - pp.factor * off;
            sNew[c] = vc;
            q[c] = qv;
            part += vc * qv;
        SEG_WALK_END
    }
    block_add(part, &S->dotSZ[it % 3]);
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int ticket = atomicAdd(&S->pad[0], 1);
        last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        S->pad[0] = 0;
        if (it > 0) {
            S->pcgIterations = it;
            S->pcgError = rmax;
            if (converged) S->pcgDone = 1;
            else if (breakdown) S->pcgDone = 3;
        }
        // slots of the rotation that nobody reads or accumulates during this launch
        S->dotSZ[(it + 1) % 3] = 0.0;
        S->rMaxBits[it % 3] = 0ull;
        S->rho[(it + 1) % 3] = 0.0;
    }
}

// host-side dispatch on the sweep mode
static void launch_mg0_sweep(int blocks, cudaStream_t st, const int *segCell, const unsigned int *segMask, const Mg0 &M,
                             const MgLevel &C, const double *r, const float *xin, const float *e, float *xout, float omega,
                             float scale, int mode, int last, double *zout, DeviceScalars *S, int rhoSlot, float omega0) {
    switch (mode) {
        case 0: k_mg0_sweep<0><<<blocks, TPB, 0, st>>>(segCell, segMask, M, C, r, xin, e, xout, omega, scale, last, zout, S, rhoSlot, omega0); break;
        case 1: k_mg0_sweep<1><<<blocks, TPB, 0, st>>>(segCell, segMask, M, C, r, xin, e, xout, omega, scale, last, zout, S, rhoSlot, omega0); break;
        case 2: k_mg0_sweep<2><<<blocks, TPB, 0, st>>>(segCell, segMask, M, C, r, xin, e, xout, omega, scale, last, zout, S, rhoSlot, omega0); break;
        default: k_mg0_sweep<3><<<blocks, TPB, 0, st>>>(segCell, segMask, M, C, r, xin, e, xout, omega, scale, last, zout, S, rhoSlot, omega0); break;
    }
}

// b_1(C) = sum over the row children of (r - A0 x0)
__global__ void k_mg0_restrict(Mg0 M, MgLevel C, const double *__restrict__ r, const float *x, const DeviceScalars *S) {
    if (S->pcgDone) return;
    int cc = blockIdx.x * blockDim.x + threadIdx.x;
    if (cc >= C.n) return;
    if (C.invD[cc] == 0.0f) { C.b[cc] = 0.0f; return; }
    int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int i = 2 * ci + (q & 1), j = 2 * cj + ((q >> 1) & 1), k = 2 * ck + (q >> 2);
        if (i >= M.g.I || j >= M.g.J || k >= M.g.K) continue;
        int c = i + M.g.sj * j + M.g.sk * k;
        if (!((M.rowBits[c >> 5] >> (c & 31)) & 1u)) continue;
        s += (float)r[c] - ((float)M.Adiag[c] * x[c] - mg0_offsum(M, x, c));
    }
    C.b[cc] = s;
}

// the same with eight threads per coarse cell over the active segments of level 1 (see k_mg_restrict_list)
__global__ void __launch_bounds__(256) k_mg0_restrict_list(Mg0 M, MgLevel C, const double *__restrict__ r, const float *x,
                                                           const int *__restrict__ list, const int *__restrict__ count,
                                                           const DeviceScalars *S) {
    if (S->pcgDone) return;
    const int nseg = *count;
    const int t = threadIdx.x, q = t & 7;
    for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
        const int cc = list[s] + (t >> 3);
        float v = 0.0f;
        const bool act = (cc < C.n) && C.invD[cc] != 0.0f;
        if (act) {
            int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
            int i = 2 * ci + (q & 1), j = 2 * cj + ((q >> 1) & 1), k = 2 * ck + (q >> 2);
            if (i < M.g.I && j < M.g.J && k < M.g.K) {
                int c = i + M.g.sj * j + M.g.sk * k;
                if ((M.rowBits[c >> 5] >> (c & 31)) & 1u) v = (float)r[c] - ((float)M.Adiag[c] * x[c] - mg0_offsum(M, x, c));
            }
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (q == 0 && cc < C.n) C.b[cc] = act ? v : 0.0f;
    }
}

// ---- hierarchy construction (every solve: the liquid region changes every substep)
__global__ void k_mg0_build(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                            const double *__restrict__ Adiag, float *__restrict__ invD, const DeviceScalars *S) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= S->numSegments) return;
    if ((segMask[warp] >> lane) & 1u) {
        int c = segCell[warp] + lane;
        double d = Adiag[c];
        invD[c] = (d > 0.0) ? (float)(1.0 / d) : 0.0f;
    }
}

// level 1 from level 0
__global__ void k_mg_coarsen0(Mg0 M, MgLevel C) {
    int cc = blockIdx.x * blockDim.x + threadIdx.x;
    if (cc >= C.n) return;
    int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
    float d = 0.0f, u = 0.0f, v = 0.0f, w = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int qi = q & 1, qj = (q >> 1) & 1, qk = q >> 2;
        int i = 2 * ci + qi, j = 2 * cj + qj, k = 2 * ck + qk;
        if (i >= M.g.I || j >= M.g.J || k >= M.g.K) continue;
        int c = i + M.g.sj * j + M.g.sk * k;
        if (!((M.rowBits[c >> 5] >> (c & 31)) & 1u)) continue;
        d += (float)M.Adiag[c];
        float au = M.fac * M.oU[c], av = M.fac * M.oV[c], aw = M.fac * M.oW[c];
        if (M.g.cut && qk == 0 && k + 1 >= M.g.kOwn1) aw = 0.0f;   // partner inside the aggregate belongs to another z-slab
        if (qi == 0) d -= 2.0f * au; else u += au;
        if (qj == 0) d -= 2.0f * av; else v += av;
        if (qk == 0) d -= 2.0f * aw; else w += aw;
    }
    C.diag[cc] = d;
    C.invD[cc] = (d > 0.0f) ? 1.0f / d : 0.0f;
    C.oU[cc] = u; C.oV[cc] = v; C.oW[cc] = w;
}

// level l+1 from level l (l >= 1)
__global__ void k_mg_coarsen(MgLevel F, MgLevel C) {
    int cc = blockIdx.x * blockDim.x + threadIdx.x;
    if (cc >= C.n) return;
    int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
    float d = 0.0f, u = 0.0f, v = 0.0f, w = 0.0f;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int qi = q & 1, qj = (q >> 1) & 1, qk = q >> 2;
        int i = 2 * ci + qi, j = 2 * cj + qj, k = 2 * ck + qk;
        if (i >= F.I || j >= F.J || k >= F.K) continue;
        int c = i + F.sj * j + F.sk * k;
        d += F.diag[c];
        float au = F.oU[c], av = F.oV[c], aw = F.oW[c];
        if (qi == 0) d -= 2.0f * au; else u += au;
        if (qj == 0) d -= 2.0f * av; else v += av;
        if (qk == 0) d -= 2.0f * aw; else w += aw;
    }
    C.diag[cc] = d;
    C.invD[cc] = (d > 0.0f) ? 1.0f / d : 0.0f;
    C.oU[cc] = u; C.oV[cc] = v; C.oW[cc] = w;
}

__global__ void k_mg_invdiag(MgLevel L) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L.n) return;
    float d = L.diag[c];
    L.invD[c] = (d > 0.0f) ? 1.0f / d : 0.0f;
}

// s = z on rows (start of PCG with a preconditioner that produced z in separate passes)
__global__ void k_pcg_copy_zs(const int *__restrict__ segCell, const unsigned int *__restrict__ segMask,
                              const double *__restrict__ z, double *__restrict__ s, const DeviceScalars *S) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= S->numSegments) return;
    if ((segMask[warp] >> lane) & 1u) {
        int c = segCell[warp] + lane;
        s[c] = z[c];
    }
}


// ------------------------------------------------------------------------------------------------
// The whole PCG solve as ONE persistent cooperative kernel.
//
// At n ~ 2 M rows every pass of the solver is a 10-30 us kernel, so the multi-launch formulation
// above is bound by launch gaps, cold L1 and host polling rather than by HBM.  Here one grid that
// fills the machine (148 SMs x resident CTAs) walks the same phases with grid-wide barriers in
// between: the iteration loop, alpha/beta, the convergence test (||r||_inf <= tol, tested right
// after the residual update as pcgsolver.h:286-288 does, i.e. before the next preconditioner
// application) and the multigrid V-cycle all live on the device; the host launches once and reads
// the scalars back once.  Levels small enough for one CTA are run by block 0 between two barriers.
// The arithmetic per row is the same as in the multi-launch kernels (same device functions).
// ------------------------------------------------------------------------------------------------
struct PcgDev {
    const int *segCell;
    const unsigned int *segMask;
    PGrid g;
    double factor;
    const double *Adiag;
    const float *oU, *oV, *oW;
    const double *b;
    double *x, *r, *s, *z;
    Mg0 m0;
    MgLevel lv[MG_MAX_LEVELS];
    int numLevels, firstSmall;
    MgParams mp;
    int useMg, maxIter;
    int warm;          // x holds an initial guess
    DeviceScalars *S;
};

__device__ __forceinline__ double ldcg_d(const double *p) { return __ldcg(p); }

// dense-level sweep / restrict over a grid-stride range
__device__ __forceinline__ void pg_sweep_level(const MgLevel &L, const MgLevel &C, const float *xin, const float *e,
                                               float *xout, float omega, float scale, int mode, int gtid, int gthreads) {
    for (int c = gtid; c < L.n; c += gthreads) {
        int i = c % L.I, j = (c / L.I) % L.J, k = c / L.sk;
        xout[c] = mg_sweep_cell(L, C, xin, e, omega, scale, mode, c, i, j, k);
    }
}
__device__ __forceinline__ void pg_restrict_level(const MgLevel &F, const MgLevel &C, const float *x, int gtid, int gthreads) {
    for (int c = gtid; c < C.n; c += gthreads) {
        if (C.invD[c] == 0.0f) { C.b[c] = 0.0f; continue; }
        int i = c % C.I, j = (c / C.I) % C.J, k = c / C.sk;
        C.b[c] = mg_restrict_cell(F, x, i, j, k);
    }
}

__global__ void __launch_bounds__(TPB) k_pcg_persistent(PcgDev P) {
    cg::grid_group grid = cg::this_grid();
    DeviceScalars *S = P.S;
    const int lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gthreads = gridDim.x * blockDim.x;
    const int gwarp = gtid >> 5, nwarps = gthreads >> 5;
    const int nseg = S->numSegments;
    const float scale = P.mp.scale;
    const int nu = P.mp.nu;
    const float omCoarse = P.mp.omegaCoarse;
    auto preOm = [&](int sw) { return P.mp.om[sw]; };
    auto postOm = [&](int sw) { return P.mp.om[nu - 1 - sw]; };
    const int L = P.numLevels;
    const int fs = P.firstSmall ? P.firstSmall : L;
    float *x0a = P.lv[0].x, *x0b = P.lv[0].x2;

    // ---- one V-cycle on r; the level-0 pre-sweep from the zero guess has already been written to x0a.
    // Leaves z and accumulates rho[rhoSlot].  Every path through here is grid-uniform.
    auto vcycle_after_first_sweep = [&](int rhoSlot) {
        float *xa = x0a, *xb = x0b;
        for (int sw = 1; sw < nu; sw++) {
            const float omega = preOm(sw);
            for (int seg = gwarp; seg < nseg; seg += nwarps) {
                if ((P.segMask[seg] >> lane) & 1u) {
                    int c = P.segCell[seg] + lane;
                    float inv = P.m0.invD[c];
                    float bb = (float)P.r[c];
                    xb[c] = (inv == 0.0f) ? 0.0f : (1.0f - omega) * xa[c] + omega * inv * (bb + mg0_offsum(P.m0, xa, c));
                }
            }
            grid.sync();
            float *t = xa; xa = xb; xb = t;
        }
        // restrict level 0 -> 1
        {
            const MgLevel &C = P.lv[1];
            for (int cc = gtid; cc < C.n; cc += gthreads) {
                if (C.invD[cc] == 0.0f) { C.b[cc] = 0.0f; continue; }
                int ci = cc % C.I, cj = (cc / C.I) % C.J, ck = cc / C.sk;
                float acc = 0.0f;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    int i = 2 * ci + (q & 1), j = 2 * cj + ((q >> 1) & 1), k = 2 * ck + (q >> 2);
                    if (i >= P.g.I || j >= P.g.J || k >= P.g.K) continue;
                    int c = i + P.g.sj * j + P.g.sk * k;
                    if (!((P.m0.rowBits[c >> 5] >> (c & 31)) & 1u)) continue;
                    acc += (float)P.r[c] - ((float)P.Adiag[c] * xa[c] - mg0_offsum(P.m0, xa, c));
                }
                C.b[cc] = acc;
            }
        }
        grid.sync();
        // down through the big coarse levels
        float *lx[MG_MAX_LEVELS];
        for (int l = 1; l < fs && l < L - 1; l++) {
            const MgLevel &Lv = P.lv[l];
            float *ya = Lv.x, *yb = Lv.x2;
            for (int sw = 0; sw < nu; sw++) {
                pg_sweep_level(Lv, Lv, ya, nullptr, yb, preOm(sw), scale, sw == 0 ? 0 : 1, gtid, gthreads);
                grid.sync();
                float *t = ya; ya = yb; yb = t;
            }
            lx[l] = ya;
            pg_restrict_level(Lv, P.lv[l + 1], ya, gtid, gthreads);
            grid.sync();
        }
        // small levels: block 0 alone, CTA barriers only
        if (fs < L) {
            if (blockIdx.x == 0) {
                const int tid = threadIdx.x, nt = blockDim.x;
                for (int l = fs; l <= L - 1; l++) {
                    const MgLevel &Lv = P.lv[l];
                    int sweeps = (l == L - 1) ? P.mp.coarseSweeps : nu;
                    float *ya = Lv.x, *yb = Lv.x2;
                    for (int sw = 0; sw < sweeps; sw++) {
                        pg_sweep_level(Lv, Lv, ya, nullptr, yb, (l == L - 1) ? omCoarse : preOm(sw), scale, sw == 0 ? 0 : 1, tid, nt);
                        __syncthreads();
                        float *t = ya; ya = yb; yb = t;
                    }
                    lx[l] = ya;
                    if (l < L - 1) {
                        pg_restrict_level(Lv, P.lv[l + 1], ya, tid, nt);
                        __syncthreads();
                    }
                }
                for (int l = L - 2; l >= fs; l--) {
                    const MgLevel &Lv = P.lv[l];
                    float *ya = lx[l], *yb = (ya == Lv.x) ? Lv.x2 : Lv.x;
                    for (int sw = 0; sw < nu; sw++) {
                        pg_sweep_level(Lv, P.lv[l + 1], ya, lx[l + 1], yb, postOm(sw), scale, sw == 0 ? 2 : 1, tid, nt);
                        __syncthreads();
                        float *t = ya; ya = yb; yb = t;
                    }
                    lx[l] = ya;
                }
                // publish where the result of level fs sits: always copy into Lv.x so that every block agrees
                const MgLevel &Lf = P.lv[fs];
                if (lx[fs] != Lf.x) {
                    for (int c = tid; c < Lf.n; c += nt) Lf.x[c] = lx[fs][c];
                }
            }
            lx[fs] = P.lv[fs].x;
            grid.sync();
        } else {
            const MgLevel &Lv = P.lv[L - 1];
            float *ya = Lv.x, *yb = Lv.x2;
            for (int sw = 0; sw < P.mp.coarseSweeps; sw++) {
                pg_sweep_level(Lv, Lv, ya, nullptr, yb, omCoarse, scale, sw == 0 ? 0 : 1, gtid, gthreads);
                grid.sync();
                float *t = ya; ya = yb; yb = t;
            }
            lx[L - 1] = ya;
        }
        // up through the big coarse levels
        int top = (fs < L ? fs : L - 1) - 1;
        for (int l = top; l >= 1; l--) {
            const MgLevel &Lv = P.lv[l];
            float *ya = lx[l], *yb = (ya == Lv.x) ? Lv.x2 : Lv.x;
            for (int sw = 0; sw < nu; sw++) {
                pg_sweep_level(Lv, P.lv[l + 1], ya, lx[l + 1], yb, postOm(sw), scale, sw == 0 ? 2 : 1, gtid, gthreads);
                grid.sync();
                float *t = ya; ya = yb; yb = t;
            }
            lx[l] = ya;
        }
        // level 0 post-smoothing; the last sweep writes z and accumulates rho
        const MgLevel &C1 = P.lv[1];
        const float *e1 = lx[1];
        double part = 0.0;
        for (int sw = 0; sw < nu; sw++) {
            const bool last = (sw == nu - 1);
            const float omega = postOm(sw);
            for (int seg = gwarp; seg < nseg; seg += nwarps) {
                if ((P.segMask[seg] >> lane) & 1u) {
                    int c = P.segCell[seg] + lane;
                    float inv = P.m0.invD[c];
                    double rd = P.r[c];
                    float bb = (float)rd;
                    float xn;
                    if (inv == 0.0f) xn = 0.0f;
                    else if (sw == 0) xn = (1.0f - omega) * mg0_corr(P.m0, C1, xa, e1, scale, c) + omega * inv * (bb + mg0_offsum_corr(P.m0, C1, xa, e1, scale, c));
                    else xn = (1.0f - omega) * xa[c] + omega * inv * (bb + mg0_offsum(P.m0, xa, c));
                    xb[c] = xn;
                    if (last) { P.z[c] = (double)xn; part += (double)xn * rd; }
                }
            }
            if (last) block_add(part, &S->rho[rhoSlot]);
            grid.sync();
            float *t = xa; xa = xb; xb = t;
        }
        x0a = xa; x0b = xb;
    };

    // ---- prologue: r = b, z = M^-1 r, s = z, rho[0] = z.r      (pcgsolver.h:258-276)
    {
        double part = 0.0;
        for (int seg = gwarp; seg < nseg; seg += nwarps) {
            if ((P.segMask[seg] >> lane) & 1u) {
                int c = P.segCell[seg] + lane;
                double rv = P.b[c];
                if (P.warm) rv -= apply_row(P.g, c, P.factor, P.Adiag, P.oU, P.oV, P.oW, P.x);
                P.r[c] = rv;
                if (P.useMg) {
                    x0a[c] = preOm(0) * P.m0.invD[c] * (float)rv;
                } else {
                    double d = P.Adiag[c];
                    double zv = (d != 0.0) ? rv / d : 0.0;
                    P.z[c] = zv;
                    P.s[c] = zv;
                    part += zv * rv;
                }
            }
        }
        if (!P.useMg) block_add(part, &S->rho[0]);
        grid.sync();
        if (P.useMg) {
            vcycle_after_first_sweep(0);
            for (int seg = gwarp; seg < nseg; seg += nwarps) {
                if ((P.segMask[seg] >> lane) & 1u) {
                    int c = P.segCell[seg] + lane;
                    P.s[c] = P.z[c];
                }
            }
            grid.sync();
        }
    }

    const double tol = S->pcgTol;
    int it = 0;
    int done = 0;
    double rmax = S->pcgError;
    for (; it < P.maxIter; it++) {
        const int sl = it % 3, sn = (it + 1) % 3, sp = (it + 2) % 3;
        // ---- z = A s, dotSZ = s.z
        {
            double part = 0.0;
            for (int seg = gwarp; seg < nseg; seg += nwarps) {
                if ((P.segMask[seg] >> lane) & 1u) {
                    int c = P.segCell[seg] + lane;
                    double zv = apply_row(P.g, c, P.factor, P.Adiag, P.oU, P.oV, P.oW, P.s);
                    P.z[c] = zv;
                    part += P.s[c] * zv;
                }
            }
            block_add(part, &S->dotSZ[sl]);
        }
        grid.sync();
        // ---- x += alpha s, r -= alpha z, ||r||_inf; first pre-sweep of the V-cycle (or Jacobi z, rho)
        {
            const double alpha = ldcg_d(&S->rho[sl]) / ldcg_d(&S->dotSZ[sl]);
            double rabs = 0.0, part = 0.0;
            for (int seg = gwarp; seg < nseg; seg += nwarps) {
                if ((P.segMask[seg] >> lane) & 1u) {
                    int c = P.segCell[seg] + lane;
                    P.x[c] += alpha * P.s[c];
                    double rv = P.r[c] - alpha * P.z[c];
                    P.r[c] = rv;
                    rabs = fmax(rabs, fabs(rv));
                    if (P.useMg) {
                        x0a[c] = preOm(0) * P.m0.invD[c] * (float)rv;
                    } else {
                        double d = P.Adiag[c];
                        double zv = (d != 0.0) ? rv / d : 0.0;
                        P.z[c] = zv;
                        part += zv * rv;
                    }
                }
            }
            block_max(rabs, &S->rMaxBits[sl]);
            if (!P.useMg) block_add(part, &S->rho[sn]);
        }
        grid.sync();
        rmax = __longlong_as_double((long long)__ldcg(&S->rMaxBits[sl]));
        if (rmax <= tol) { done = 1; it++; break; }
        if (P.useMg) vcycle_after_first_sweep(sn);
        // ---- beta, s = z + beta s
        {
            const double rho = ldcg_d(&S->rho[sl]), rhoNew = ldcg_d(&S->rho[sn]);
            if (rhoNew == 0.0 || rhoNew != rhoNew) { done = 3; it++; break; }
            const double beta = rhoNew / rho;
            for (int seg = gwarp; seg < nseg; seg += nwarps) {
                if ((P.segMask[seg] >> lane) & 1u) {
                    int c = P.segCell[seg] + lane;
                    P.s[c] = P.z[c] + beta * P.s[c];
                }
            }
            if (gtid == 0) { S->dotSZ[sn] = 0.0; S->rMaxBits[sn] = 0ull; S->rho[sp] = 0.0; }
        }
        grid.sync();
    }
    if (gtid == 0) {
        S->pcgIterations = it;
        S->pcgError = rmax;
        S->pcgDone = done;
    }
}

// ---- 3. velocity update  (_applySolutionToVelocityField, pressuresolver.cpp:842-1047)
struct ApplyParams {
    PGrid g;
    float factor;   // (float)(_deltaTime/_dx)
};

__device__ __forceinline__ float row_pressure(const float *__restrict__ phi, const double *__restrict__ x,
                                              const PGrid &g, int c) {
    // pressureGrid is 0 except at pressure cells, where it is (float)soln  (:843-847)
    return is_row(phi, g, c) ? (float)x[c] : 0.0f;
}

// One CTA row per (j, k) line of faces, threads along i (no index divisions; the row test takes the coordinates).
template <int DIR>
__global__ void k_apply_pressure(ApplyParams ap, const float *__restrict__ phi, const double *__restrict__ x,
                                 const float *__restrict__ wgt, float *__restrict__ vel,
                                 unsigned char *__restrict__ valid) {
    const PGrid &g = ap.g;
    const int gi = g.I + (DIR == 0), gj = g.J + (DIR == 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;     // k: local plane
    if (i >= gi) return;
    const size_t t = (size_t)i + (size_t)gi * ((size_t)j + (size_t)gj * k);
    const int a = (DIR == 0) ? i : (DIR == 1 ? j : k + g.kOff);          // global index along DIR
    const int amax = (DIR == 0) ? g.I : (DIR == 1 ? g.J : g.Kg);
    unsigned char vflag = 0;
    // z-slab: the lowest local face plane has no cell below it in local storage (halo, re-imported later)
    const bool noLowerCell = (DIR == 2) && (k == 0);
    if (noLowerCell && a != 0) { valid[t] = 0; return; }
    if (!(a == 0 || a == amax - 1)) {
        const int stride = (DIR == 0) ? 1 : (DIR == 1 ? g.sj : g.sk);
        // cell "2" is (i,j,k), cell "1" is one step back along DIR
        const int c2 = i + g.I * (j + g.J * k);
        const int c1 = c2 - stride;
        const bool in2 = (DIR == 2) ? (k < g.K) : (a < amax);       // the face past the last (local) cell
        // a pressure row of the global system (is_row): liquid cell in [1, N-2]^3
        const int kg = k + g.kOff;
        auto interior = [&](int ci, int cj, int ckg) {
            return ci >= 1 && cj >= 1 && ckg >= 1 && ci < g.I - 1 && cj < g.J - 1 && ckg < g.Kg - 1;
        };
        const bool int2 = interior(i, j, kg);
        const bool int1 = interior(i - (DIR == 0), j - (DIR == 1), kg - (DIR == 2));
        // FluidMaterialGrid::isFaceBorderingMaterial{U,V,W}  fluidmaterialgrid.cpp:126-150
        const float phi1 = __ldg(phi + c1);
        const float phi2 = in2 ? __ldg(phi + c2) : 0.0f;   // (w>0 never occurs on the outermost face of a walled domain)
        const bool f1 = phi1 < 0.0f;
        const bool f2 = in2 ? (phi2 < 0.0f) : false;
        const float w = wgt[t];
        if (w > 0.0f && (f1 || f2)) {
            // pressureGrid is 0 except at pressure cells, where it is (float)soln  (:843-847)
            float p1 = 0.0f, p2 = 0.0f;
            if (f1 && f2) {
                p1 = int1 ? (float)x[c1] : 0.0f;
                p2 = int2 ? (float)x[c2] : 0.0f;
            } else {
                const float eps = 1e-6f;
                if (f1) {
                    float theta = __fdiv_rn(phi2, fadd(phi1, eps));
                    theta = (float)fmax(-25.0, fmin((double)theta, 25.0));
                    p1 = int1 ? (float)x[c1] : 0.0f;
                    p2 = fmul(theta, p1);
                } else {
                    float theta = __fdiv_rn(phi1, fadd(phi2, eps));
                    theta = (float)fmax(-25.0, fmin((double)theta, 25.0));
                    p2 = int2 ? (float)x[c2] : 0.0f;
                    p1 = fmul(theta, p2);
                }
            }
            vel[t] = fadd(vel[t], fmul(-ap.factor, fsub(p2, p1)));
            vflag = 1;
        } else {
            vel[t] = 0.0f;
        }
    }
    valid[t] = vflag;
}

__global__ void k_pressure_to_float(const float *__restrict__ phi, const double *__restrict__ x, PGrid g, int nC,
                                    float *__restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    out[c] = row_pressure(phi, x, g, c);
}

__global__ void k_reset_pressure_scalars(DeviceScalars *S) {
    S->numRows = 0; S->numSegments = 0; S->rhsMaxBits = 0ull; S->pad[0] = 0;
    S->pcgIterations = 0; S->pcgDone = 0; S->pcgError = 0.0;
}

// ------------------------------------------------------------------------------------------------
struct PressureScratch {
    unsigned int *maskAll = nullptr;
    unsigned int *maskPrev = nullptr;  // row bits of the previous solve (warm start)
    int *flagAll = nullptr;
    int *posAll = nullptr;
    int nSegAll = 0;
    int lastIterations = 0;            // PCG iterations of the previous solve (sizes the first poll-free batch)
    // multigrid hierarchy
    int numLevels = 0;                 // including level 0
    int firstSmall = 0;                // levels firstSmall..numLevels-1 run in one CTA (0: none)
    MgLevel lv[MG_MAX_LEVELS];         // lv[0]: only invD, x, x2 are used (dense over cells)
    float *pool = nullptr;             // one allocation behind all level arrays
    // active 32-cell segments of the levels >= 1 (single-GPU solver): lists and their sizes on the device
    int *segPool = nullptr;
    int *lvSeg[MG_MAX_LEVELS] = {nullptr};
    int *lvSegCount = nullptr;         // [MG_MAX_LEVELS]
    // z-slabs: the same lists for the replicated (global) levels
    int *gSegPool = nullptr;
    int *gSeg[MG_MAX_LEVELS] = {nullptr};
    int *gSegCount = nullptr;          // [MG_MAX_LEVELS]
    size_t coarseSmemBytes = 0;        // dynamic shared memory of k_mg_coarse (0: not usable, per-pass launches instead)
    int coarseBlocks = 0;              // its cooperative grid: one CTA per SM
    double *vq = nullptr, *vs2 = nullptr, *vr2 = nullptr;   // fused PCG passes: q = A s, ping-pong partners of s and r
    int coarseGroup = 1 << 20;         // CTAs of that grid that run the levels >= 2 (k_mg_coarse); default: all of them
                                       // (measured at 2 M rows: V-cycle 0.293 / 0.263 / 0.248 / 0.242 / 0.238 ms with
                                       // 4 / 8 / 16 / 32 / 148 CTAs -- the passes are bound by their chains of dependent
                                       // loads, which more CTAs walk in fewer rounds, not by the barrier)
    unsigned int *groupBar = nullptr;  // their barrier counter
    float *coarseBase = nullptr;       // levels >= 1 inside `pool` (L2 persistence window)
    size_t coarseBytes = 0, l2SetAside = 0, l2Window = 0;
    unsigned long long *trace = nullptr;   // FLIP_MG_TRACE developer probe
    float *smallInv = nullptr;             // dense inverse of the first level of <= SM_WARP_CELLS cells, and its validity flag
    int *smallInvOk = nullptr;
    int smallLevel = 0;
    int coopBlocks = 0;                // grid of the persistent solver (SMs x resident CTAs)
    // z-slabs: levels >= Lc span the WHOLE domain and are held (redundantly) by every rank; the
    // restricted residual of level Lc is all-gathered once per V-cycle.  Lc == 0: every level is local.
    int Lc = 0;
    int gLevels = 0;                   // total number of levels when Lc > 0 (levels Lc..gLevels-1 are in glv)
    int gFirstSmall = 0;
    MgLevel glv[MG_MAX_LEVELS];
    float *gpool = nullptr;
};

void pressure_alloc(flip_ctx *c) {
    const Dims &d = c->d;
    int nSeg = cdiv(d.nC, 32);
    c->maxSegments = nSeg;
    FLIP_CUDA_CHECK(cudaMalloc(&c->segCell, sizeof(int) * nSeg));
    FLIP_CUDA_CHECK(cudaMalloc(&c->segMask, sizeof(unsigned int) * nSeg));
    size_t pad = (size_t)d.nC + 64;
    FLIP_CUDA_CHECK(cudaMalloc(&c->Adiag, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->AoffU, sizeof(float) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->AoffV, sizeof(float) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->AoffW, sizeof(float) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->vx_, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->vr, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->vs, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->vz, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMalloc(&c->vb, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMemset(c->AoffU, 0, sizeof(float) * pad));
    FLIP_CUDA_CHECK(cudaMemset(c->AoffV, 0, sizeof(float) * pad));
    FLIP_CUDA_CHECK(cudaMemset(c->AoffW, 0, sizeof(float) * pad));
    FLIP_CUDA_CHECK(cudaMemset(c->vx_, 0, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMemset(c->vs, 0, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMemset(c->vz, 0, sizeof(double) * pad));
    FLIP_CUDA_CHECK(cudaMemset(c->vr, 0, sizeof(double) * pad));
    PressureScratch *ps = new PressureScratch();
    ps->nSegAll = nSeg;
    FLIP_CUDA_CHECK(cudaMalloc(&ps->maskAll, sizeof(unsigned int) * nSeg));
    FLIP_CUDA_CHECK(cudaMalloc(&ps->maskPrev, sizeof(unsigned int) * nSeg));
    FLIP_CUDA_CHECK(cudaMemset(ps->maskAll, 0, sizeof(unsigned int) * nSeg));
    FLIP_CUDA_CHECK(cudaMemset(ps->maskPrev, 0, sizeof(unsigned int) * nSeg));
    for (double **v : {&ps->vq, &ps->vs2, &ps->vr2}) {
        FLIP_CUDA_CHECK(cudaMalloc(v, sizeof(double) * pad));
        FLIP_CUDA_CHECK(cudaMemset(*v, 0, sizeof(double) * pad));
    }
    FLIP_CUDA_CHECK(cudaMalloc(&ps->flagAll, sizeof(int) * (nSeg + 1)));
    FLIP_CUDA_CHECK(cudaMalloc(&ps->posAll, sizeof(int) * (nSeg + 1)));
    c->mg = ps;

    // multigrid levels: halve (round up) until the grid is at most 2 cells wide
    {
        int I = d.I, J = d.J, K = d.K;
        int L = 0;
        size_t total = 0;
        while (L < MG_MAX_LEVELS) {
            MgLevel &lv = ps->lv[L];
            lv.I = I; lv.J = J; lv.K = K; lv.sj = I; lv.sk = I * J; lv.n = I * J * K; lv.cut = 0;
            size_t np = (size_t)lv.n + 64 + (L == 0 ? 0 : 2 * (size_t)mg_pad(lv.sk));
            total += (L == 0) ? 3 * np : 8 * np;
            L++;
            if (I <= 2 && J <= 2 && K <= 2) break;
            I = (I + 1) / 2; J = (J + 1) / 2; K = (K + 1) / 2;
        }
        ps->numLevels = L;
        FLIP_CUDA_CHECK(cudaMalloc(&ps->pool, sizeof(float) * total));
        FLIP_CUDA_CHECK(cudaMemset(ps->pool, 0, sizeof(float) * total));
        float *q = ps->pool;
        ps->firstSmall = 0;
        for (int l = 0; l < L; l++) {
            MgLevel &lv = ps->lv[l];
            // levels >= 1: every array is zero-padded by at least one plane on both sides (the branch-free passes
            // of k_mg_coarse read neighbours unconditionally); the pads are never written
            const size_t pad = (l == 0) ? 0 : (size_t)mg_pad(lv.sk);
            size_t np = (size_t)lv.n + 64 + 2 * pad;
            lv.diag = lv.oU = lv.oV = lv.oW = lv.b = nullptr;
            if (l == 1) { ps->coarseBase = q; ps->coarseBytes = sizeof(float) * (total - (size_t)(q - ps->pool)); }
            lv.invD = q + pad; q += np;
            lv.x = q + pad; q += np;
            lv.x2 = q + pad; q += np;
            if (l > 0) {
                lv.diag = q + pad; q += np;
                lv.oU = q + pad; q += np;
                lv.oV = q + pad; q += np;
                lv.oW = q + pad; q += np;
                lv.b = q + pad; q += np;
                if (ps->firstSmall == 0 && lv.n <= MG_SMALL) ps->firstSmall = l;
            }
        }
        // L2 set-aside for the coarse levels (see stage_pressure)
        {
            int maxPersist = 0, maxWindow = 0;
            cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
            cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
            ps->l2SetAside = 0;
            if (maxPersist > 0 && maxWindow > 0 && ps->coarseBase && !getenv("FLIP_MG_NO_L2PIN")) {
                size_t want = std::min((size_t)maxPersist, (size_t)48 << 20);
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                    ps->l2SetAside = want;
                    ps->l2Window = std::min(ps->coarseBytes, (size_t)maxWindow);
                }
            }
            cudaGetLastError();
        }
    }
    // active-segment lists of the coarse levels, and the shared-memory budget of the single-CTA levels
    {
        size_t ints = 0;
        for (int l = 1; l < ps->numLevels; l++) ints += (size_t)cdiv(ps->lv[l].n, 32) + 1;
        FLIP_CUDA_CHECK(cudaMalloc(&ps->segPool, sizeof(int) * (ints + MG_MAX_LEVELS)));
        FLIP_CUDA_CHECK(cudaMemset(ps->segPool, 0, sizeof(int) * (ints + MG_MAX_LEVELS)));
        ps->lvSegCount = ps->segPool;
        int *q = ps->segPool + MG_MAX_LEVELS;
        for (int l = 1; l < ps->numLevels; l++) { ps->lvSeg[l] = q; q += cdiv(ps->lv[l].n, 32) + 1; }
        ps->coarseSmemBytes = 0;
        if (ps->firstSmall && !getenv("FLIP_MG_NO_COOP")) {
            size_t bytes = 0;
            for (int l = ps->firstSmall; l < ps->numLevels; l++) bytes += sizeof(float) * sm_level_floats(ps->lv[l].n, ps->lv[l].sk);
            int maxOptin = 0, coop = 0, sms = 0, perSm = 0;
            FLIP_CUDA_CHECK(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
            FLIP_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
            FLIP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
            if (coop && bytes + 4096 <= (size_t)maxOptin) {
                FLIP_CUDA_CHECK(cudaFuncSetAttribute(k_mg_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
                FLIP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_mg_coarse, 1024, bytes));
                if (perSm >= 1) {
                    ps->coarseSmemBytes = bytes;
                    ps->coarseBlocks = sms;
                    if (const char *e = getenv("FLIP_MG_COARSE_BLOCKS")) ps->coarseBlocks = std::max(2, std::min(sms, atoi(e)));   // tuning knob
                    if (const char *e = getenv("FLIP_MG_GROUP")) ps->coarseGroup = std::max(1, std::min(sms, atoi(e)));            // tuning knob
                    FLIP_CUDA_CHECK(cudaMalloc(&ps->groupBar, 64));
                    FLIP_CUDA_CHECK(cudaMemset(ps->groupBar, 0, 64));
                    for (int l = ps->firstSmall; l < ps->numLevels; l++)
                        if (ps->lv[l].n <= SM_WARP_CELLS) { ps->smallLevel = l; break; }
                    if (ps->smallLevel && !getenv("FLIP_MG_NO_DIRECT")) {
                        FLIP_CUDA_CHECK(cudaMalloc(&ps->smallInv, sizeof(float) * SM_WARP_CELLS * SM_WARP_CELLS));
                        FLIP_CUDA_CHECK(cudaMemset(ps->smallInv, 0, sizeof(float) * SM_WARP_CELLS * SM_WARP_CELLS));
                        FLIP_CUDA_CHECK(cudaMalloc(&ps->smallInvOk, sizeof(int)));
                        FLIP_CUDA_CHECK(cudaMemset(ps->smallInvOk, 0, sizeof(int)));
                    }
                }
                if (getenv("FLIP_MG_TRACE")) {
                    FLIP_CUDA_CHECK(cudaMalloc(&ps->trace, 64 * sizeof(unsigned long long)));
                    FLIP_CUDA_CHECK(cudaMemset(ps->trace, 0, 64 * sizeof(unsigned long long)));
                }
            }
        }
    }
    // z-slabs: global coarse levels from level Lc on, if the slab boundaries are aligned to 2^Lc planes
    ps->Lc = 0;
    if (slab_on(c)) {
        int Lc = 3;
        if (const char *e = getenv("FLIP_MG_GLOBAL_FROM")) Lc = std::max(0, std::min(6, atoi(e)));   // tuning knob
        auto aligned = [&](int l) {
            int m = (1 << l) - 1;
            return !(d.kOff & m) && !(d.kOwn0 & m) && !(d.kOwn1 & m) && !(d.K & m) && !(d.Kg & m) && (d.Kg % c->nranks) == 0 &&
                   !((d.Kg / c->nranks) & m);
        };
        while (Lc > 0 && (!aligned(Lc) || Lc >= ps->numLevels)) Lc--;
        ps->Lc = Lc;
        for (int l = 0; l < ps->numLevels; l++) ps->lv[l].cut = (Lc == 0) ? 1 : 0;
        if (Lc > 0) {
            int I = ps->lv[Lc].I, J = ps->lv[Lc].J, K = d.Kg >> Lc;
            int L = Lc;
            size_t total = 0;
            while (L < MG_MAX_LEVELS) {
                MgLevel &lv = ps->glv[L];
                lv.I = I; lv.J = J; lv.K = K; lv.sj = I; lv.sk = I * J; lv.n = I * J * K; lv.cut = 0;
                total += 8 * ((size_t)lv.n + 64);
                L++;
                if (I <= 2 && J <= 2 && K <= 2) break;
                I = (I + 1) / 2; J = (J + 1) / 2; K = (K + 1) / 2;
            }
            ps->gLevels = L;
            FLIP_CUDA_CHECK(cudaMalloc(&ps->gpool, sizeof(float) * total));
            FLIP_CUDA_CHECK(cudaMemset(ps->gpool, 0, sizeof(float) * total));
            float *q = ps->gpool;
            ps->gFirstSmall = 0;
            for (int l = Lc; l < L; l++) {
                MgLevel &lv = ps->glv[l];
                size_t np = (size_t)lv.n + 64;
                lv.invD = q; q += np; lv.x = q; q += np; lv.x2 = q; q += np; lv.diag = q; q += np;
                lv.oU = q; q += np; lv.oV = q; q += np; lv.oW = q; q += np; lv.b = q; q += np;
                if (ps->gFirstSmall == 0 && lv.n <= MG_SMALL) ps->gFirstSmall = l;
            }
            size_t ints = MG_MAX_LEVELS;
            for (int l = Lc; l < L; l++) ints += (size_t)cdiv(ps->glv[l].n, 32) + 1;
            FLIP_CUDA_CHECK(cudaMalloc(&ps->gSegPool, sizeof(int) * ints));
            FLIP_CUDA_CHECK(cudaMemset(ps->gSegPool, 0, sizeof(int) * ints));
            ps->gSegCount = ps->gSegPool;
            int *qi = ps->gSegPool + MG_MAX_LEVELS;
            for (int l = Lc; l < L; l++) { ps->gSeg[l] = qi; qi += cdiv(ps->glv[l].n, 32) + 1; }
        }
    }
}

void pressure_free(flip_ctx *c) {
    cudaFree(c->segCell); cudaFree(c->segMask); cudaFree(c->Adiag);
    cudaFree(c->AoffU); cudaFree(c->AoffV); cudaFree(c->AoffW);
    cudaFree(c->vx_); cudaFree(c->vr); cudaFree(c->vs); cudaFree(c->vz); cudaFree(c->vb);
    // (a failed re-allocation in flip_set_slab must not leave pointers that flip_destroy would free again)
    c->segCell = nullptr; c->segMask = nullptr; c->Adiag = nullptr;
    c->AoffU = c->AoffV = c->AoffW = nullptr;
    c->vx_ = c->vr = c->vs = c->vz = c->vb = nullptr;
    if (c->mg) {
        PressureScratch *ps = (PressureScratch *)c->mg;
        cudaFree(ps->maskAll); cudaFree(ps->maskPrev); cudaFree(ps->flagAll); cudaFree(ps->posAll); cudaFree(ps->pool); cudaFree(ps->gpool);
        cudaFree(ps->segPool); cudaFree(ps->gSegPool); cudaFree(ps->trace); cudaFree(ps->groupBar);
        cudaFree(ps->smallInv); cudaFree(ps->smallInvOk);
        cudaFree(ps->vq); cudaFree(ps->vs2); cudaFree(ps->vr2);
        delete ps;
        c->mg = nullptr;
    }
}

void stage_pressure(flip_ctx *c, double dt) {
    const Dims &d = c->d;
    cudaStream_t st = c->stream;
    PressureScratch *ps = (PressureScratch *)c->mg;
    PGrid g{d.I, d.J, d.K, d.I, d.I * d.J, d.kOff, d.Kg, d.kOwn0, d.kOwn1, 0};
    int nSeg = ps->nSegAll;

    // moving solids: enclosed pockets first (PressureSolver::solve, pressuresolver.cpp:46)
    if (c->solU) {
        if (!c->wC) throw ApiError(FLIP_ERR_RUNTIME, "solid velocities without centre weights");
        condition_solid_velocities(c, g);
    }
    // warm start: the row bits of the previous solve say where vx_ holds a pressure worth starting from
    const bool warm = c->pressureWarmStart != 0;
    std::swap(ps->maskAll, ps->maskPrev);
    size_t ktBuild = kt_begin(c);
    k_reset_pressure_scalars<<<1, 1, 0, st>>>(c->dS); c->launches++;
    k_seg_flag<<<cdiv((long long)nSeg * 32, TPB), TPB, 0, st>>>(c->phiL, g, d.nC, nSeg, ps->maskAll, ps->flagAll, c->dS);
    c->launches++;
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, ps->flagAll, ps->posAll, nSeg, st);
    if (bytes > c->scanTempBytes) {
        cudaFree(c->scanTemp);
        FLIP_CUDA_CHECK(cudaMalloc(&c->scanTemp, bytes));
        c->scanTempBytes = bytes;
    }
    cub::DeviceScan::ExclusiveSum(c->scanTemp, c->scanTempBytes, ps->flagAll, ps->posAll, nSeg, st); c->launches++;
    k_seg_compact<<<cdiv(nSeg, TPB), TPB, 0, st>>>(nSeg, ps->maskAll, ps->flagAll, ps->posAll, c->segCell, c->segMask, c->dS);
    c->launches++;

    // The number of active segments is only known on the device; launches are sized for the upper
    // bound we can derive on the host without a sync: every particle's cell could be its own segment.
    int segBound = nSeg;
    {
        long long byParticles = (long long)c->npStore * 27;   // a particle makes at most 27 cells liquid
        if (byParticles < segBound) segBound = (int)byParticles;
        if (segBound < 1) segBound = 1;
    }
    int segBlocks = cdiv((long long)segBound * 32, TPB);

    BuildParams bp;
    bp.g = g;
    bp.invdx = 1.0 / d.dx;
    bp.factor = dt / (d.dx * d.dx);
    k_build_system<<<segBlocks, TPB, 0, st>>>(c->segCell, c->segMask, bp, c->phiL, c->U, c->V, c->W, c->wU, c->wV,
                                             c->wW, c->Adiag, c->AoffU, c->AoffV, c->AoffW, c->vb, c->vx_, c->dS,
                                             warm ? ps->maskPrev : nullptr, c->solU, c->solV, c->solW, c->wC);
    c->launches++;
    kt_end(c, FLIP_KERNEL_PRESSURE_BUILD, ktBuild);
    const bool slab = slab_on(c);
    if (slab) {
        // the system is global: ||b||_inf and the row count over all slabs
        slab_allreduce_scalar(c, &c->dS->rhsMaxBits, COMM_MAX_U64);
        FLIP_CUDA_CHECK(cudaMemcpyAsync(&c->dS->globalRows, &c->dS->numRows, sizeof(int), cudaMemcpyDeviceToDevice, st));
        comm_allreduce(c->comm, &c->dS->globalRows, 1, COMM_SUM_I32, st);
    }
    scalars_to_host(c);
    int n = slab ? c->hS->globalRows : c->hS->numRows;
    int numSeg = c->hS->numSegments;
    double bmax;
    {
        unsigned long long bits = c->hS->rhsMaxBits;
        memcpy(&bmax, &bits, sizeof(double));
    }
    c->cur.pressure_rows = n;
    c->cur.fluid_cells = n;
    c->cur.rhs_max = bmax;
    c->cur.pcg_iterations = 0;
    c->cur.pcg_error = 0.0;
    c->cur.pcg_converged = 1;
    // early out (pressuresolver.cpp:52-62): velocities and the valid mask stay untouched
    if (n == 0 || bmax < c->pressureTol) return;

    // The coarse levels are a few MB that every V-cycle walks through ~17 latency-bound passes, while the level-0
    // passes stream ~0.5 GB through L2 in between: keep the coarse levels resident in the L2 set-aside so that
    // their dependent loads cost an L2 hit instead of a DRAM round trip.
    if (ps->l2SetAside && !slab) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = ps->coarseBase;
        attr.accessPolicyWindow.num_bytes = ps->l2Window;
        const double touched = 6.0 * (double)n + (double)(1 << 20);      // ~32 B per coarse cell, ~n/7 * 1.3 cells
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ps->l2SetAside / touched);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) {
            cudaGetLastError();
            ps->l2SetAside = 0;
        }
    }
    segBlocks = std::max(1, cdiv((long long)numSeg * 32, TPB));
    const int loopBlocks = std::min(segBlocks, 148 * 6);    // grid-stride passes: one resident wave
    // the fused kernels hold five CTAs per SM (48 registers): a grid of 148 x 6 would leave a second, nearly empty wave
    // that takes as long as the first (grid-stride loop: every CTA does the same share of the work)
    const int fusedBlocks = std::min(segBlocks, 148 * 5);
    PcgParams pp;
    pp.g = g;
    pp.factor = bp.factor;
    pp.tolFactor = fmax(c->pressureTol, 1e-30);
    const bool useMg = (c->preconditioner == 1) && ps->numLevels >= 2;
    if (slab && useMg && ps->Lc == 0) {
        // unaligned slabs: block-local V-cycle per slab; the correction is zero on the neighbours' rows,
        // which the level-0 sweeps read through the halo planes of the iterate buffers
        g.cut = 1;
        FLIP_CUDA_CHECK(cudaMemsetAsync(ps->lv[0].x, 0, sizeof(float) * (size_t)d.nC, st));
        FLIP_CUDA_CHECK(cudaMemsetAsync(ps->lv[0].x2, 0, sizeof(float) * (size_t)d.nC, st));
    }
    int jacobi = useMg ? 0 : 1;
    Mg0 m0;
    MgParams mp;
    for (int q = 0; q < MG_MAX_SWEEPS; q++) mp.om[q] = (float)c->mgOmegaSched[q];
    mp.omegaCoarse = (float)c->mgOmega;
    mp.scale = (float)c->mgScale; mp.nu = c->mgNu; mp.coarseSweeps = c->mgCoarseSweeps;
    auto pre_om = [&](int sw) { return mp.om[sw]; };
    auto post_om = [&](int sw) { return mp.om[mp.nu - 1 - sw]; };
    if (useMg) {
        m0.g = g; m0.fac = (float)bp.factor; m0.Adiag = c->Adiag;
        m0.oU = c->AoffU; m0.oV = c->AoffV; m0.oW = c->AoffW;
        m0.invD = ps->lv[0].invD; m0.rowBits = ps->maskAll;
        size_t ktB = kt_begin(c);
        k_mg0_build<<<segBlocks, TPB, 0, st>>>(c->segCell, c->segMask, c->Adiag, ps->lv[0].invD, c->dS); c->launches++;
        k_mg_coarsen0<<<cdiv(ps->lv[1].n, TPB), TPB, 0, st>>>(m0, ps->lv[1]); c->launches++;
        const int Lc = slab ? ps->Lc : 0;
        if (Lc == 0) {
            for (int l = 1; l + 1 < ps->numLevels; l++) {
                k_mg_coarsen<<<cdiv(ps->lv[l + 1].n, TPB), TPB, 0, st>>>(ps->lv[l], ps->lv[l + 1]); c->launches++;
            }
            if (!slab && ps->smallInv) {
                k_mg_invert_small<<<1, 1024, 0, st>>>(ps->lv[ps->smallLevel], ps->smallInv, ps->smallInvOk);
                c->launches++;
            }
            if (!slab) {
                // active segments of the coarse levels; cells outside them keep a zero iterate for the whole solve
                FLIP_CUDA_CHECK(cudaMemsetAsync(ps->lvSegCount, 0, sizeof(int) * MG_MAX_LEVELS, st));
                const int lastListed = ps->firstSmall ? ps->firstSmall : ps->numLevels - 1;
                for (int l = 1; l <= lastListed; l++) {
                    MgLevel &lv = ps->lv[l];
                    FLIP_CUDA_CHECK(cudaMemsetAsync(lv.x, 0, sizeof(float) * lv.n, st));
                    FLIP_CUDA_CHECK(cudaMemsetAsync(lv.x2, 0, sizeof(float) * lv.n, st));
                    k_mg_build_list<<<cdiv(lv.n, TPB), TPB, 0, st>>>(lv, ps->lvSeg[l], &ps->lvSegCount[l]); c->launches++;
                }
            }
        } else {
            for (int l = 1; l < Lc; l++) {
                k_mg_coarsen<<<cdiv(ps->lv[l + 1].n, TPB), TPB, 0, st>>>(ps->lv[l], ps->lv[l + 1]); c->launches++;
            }
            // the coupling of my lowest owned coarse plane to the slab below is stored with the cell below
            for (int l = 1; l < Lc; l++)
                slab_exchange_cell_plane_f32(c, ps->lv[l].oW, ps->lv[l].sk, d.kOwn0 >> l, d.kOwn1 >> l);
            // every rank contributes the operator rows of its owned coarse planes
            MgLevel &Ll = ps->lv[Lc], &Gl = ps->glv[Lc];
            size_t off = (size_t)(d.kOwn0 >> Lc) * Ll.sk, cnt = (size_t)((d.kOwn1 - d.kOwn0) >> Lc) * Ll.sk;
            comm_allgather_f32(c->comm, Ll.diag + off, Gl.diag, cnt, st);
            comm_allgather_f32(c->comm, Ll.oU + off, Gl.oU, cnt, st);
            comm_allgather_f32(c->comm, Ll.oV + off, Gl.oV, cnt, st);
            comm_allgather_f32(c->comm, Ll.oW + off, Gl.oW, cnt, st);
            k_mg_invdiag<<<cdiv(Gl.n, TPB), TPB, 0, st>>>(Gl); c->launches++;
            for (int l = Lc; l + 1 < ps->gLevels; l++) {
                k_mg_coarsen<<<cdiv(ps->glv[l + 1].n, TPB), TPB, 0, st>>>(ps->glv[l], ps->glv[l + 1]); c->launches++;
            }
            // active segments of the slab-local levels 1..Lc and of the replicated levels: the dam-break column
            // fills an eighth of the box, a dense pass over a level spends most of its time on air
            FLIP_CUDA_CHECK(cudaMemsetAsync(ps->lvSegCount, 0, sizeof(int) * MG_MAX_LEVELS, st));
            FLIP_CUDA_CHECK(cudaMemsetAsync(ps->gSegCount, 0, sizeof(int) * MG_MAX_LEVELS, st));
            for (int l = 1; l <= Lc; l++) {
                k_mg_build_list<<<cdiv(ps->lv[l].n, TPB), TPB, 0, st>>>(ps->lv[l], ps->lvSeg[l], &ps->lvSegCount[l]); c->launches++;
            }
            for (int l = Lc; l < ps->gLevels; l++) {
                k_mg_build_list<<<cdiv(ps->glv[l].n, TPB), TPB, 0, st>>>(ps->glv[l], ps->gSeg[l], &ps->gSegCount[l]); c->launches++;
            }
        }
        kt_end(c, FLIP_KERNEL_PRESSURE_BUILD, ktB);
    }
    // effective level table: local levels below Lc, global (replicated) levels from Lc on
    const int Lc = (slab && useMg) ? ps->Lc : 0;
    const int L = Lc ? ps->gLevels : ps->numLevels;
    MgLevel *LV[MG_MAX_LEVELS];
    for (int l = 0; l < L; l++) LV[l] = (Lc && l >= Lc) ? &ps->glv[l] : &ps->lv[l];
    const int fsCfg = Lc ? ps->gFirstSmall : ps->firstSmall;
    const int fs = fsCfg ? std::max(fsCfg, std::max(Lc, 1)) : L;    // first level run by the single-CTA kernel
    // what level l sees of level l+1: the descriptor its parent indices refer to, and the coarse solution
    auto up_desc = [&](int l) -> MgLevel & { return (Lc && l + 1 == Lc) ? ps->lv[Lc] : *LV[l + 1]; };
    auto up_x = [&](int l) -> const float * {
        if (Lc && l + 1 == Lc) return ps->glv[Lc].x + (size_t)(d.kOff >> Lc) * ps->lv[Lc].sk;
        return LV[l + 1]->x;
    };
    // after restricting into level l+1: the global level needs every rank's owned planes
    auto after_restrict = [&](int l) {
        if (Lc && l + 1 == Lc) {
            MgLevel &Ll = ps->lv[Lc], &Gl = ps->glv[Lc];
            size_t off = (size_t)(d.kOwn0 >> Lc) * Ll.sk, cnt = (size_t)((d.kOwn1 - d.kOwn0) >> Lc) * Ll.sk;
            comm_allgather_f32(c->comm, Ll.b + off, Gl.b, cnt, st);
        }
    };
    // aligned z-slabs: the local levels exchange one boundary plane of the iterate after every sweep, so the
    // V-cycle is the same operator as on a single GPU
    auto xchg = [&](int l, float *x) {
        if (Lc && l < Lc) slab_exchange_cell_plane_f32(c, x, ps->lv[l].sk, d.kOwn0 >> l, d.kOwn1 >> l);
    };
    // aligned z-slabs: every level >= 1 runs over its active-segment list (built in the set-up above)
    const bool sl = Lc > 0 && ps->gSegPool != nullptr;
    auto seg_of = [&](int l, bool localDesc) -> const int * { return (Lc && l >= Lc && !localDesc) ? ps->gSeg[l] : ps->lvSeg[l]; };
    auto cnt_of = [&](int l, bool localDesc) -> const int * {
        return (Lc && l >= Lc && !localDesc) ? &ps->gSegCount[l] : &ps->lvSegCount[l];
    };
    auto sl_blocks = [&](const MgLevel &lv) { return std::max(1, std::min(cdiv(cdiv(lv.n, 32), WPB), 148 * 8)); };
    auto sl_rblocks = [&](const MgLevel &coarse) { return std::max(1, std::min(cdiv(coarse.n, 32), 148 * 8)); };
    auto sweep_lvl = [&](int l, MgLevel &lv, MgLevel &C, const float *xin, const float *e, float *xout, float om, int mode) {
        if (sl)
            k_mg_sweep_list<<<sl_blocks(lv), TPB, 0, st>>>(lv, C, xin, e, xout, om, mp.scale, mode, seg_of(l, false), cnt_of(l, false),
                                                          c->dS, 0.0f);
        else
            k_mg_sweep<<<cdiv(lv.n, TPB), TPB, 0, st>>>(lv, C, xin, e, xout, om, mp.scale, mode, c->dS);
        c->launches++;
    };
    // one V-cycle: z = M^-1 r, rho[rhoSlot] += z.r
    auto vcycle = [&](int rhoSlot) {
        const int nu = mp.nu;
        size_t ktV = kt_begin(c);
        // ---- down
        {   // level 0 pre-smoothing
            float *xa = ps->lv[0].x, *xb = ps->lv[0].x2;
            for (int sw = 0; sw < nu; sw++) {
                launch_mg0_sweep(loopBlocks, st, c->segCell, c->segMask, m0, up_desc(0), c->vr, xa, nullptr, xb, pre_om(sw),
                                                      mp.scale, sw == 0 ? 0 : 1, 0, nullptr, c->dS, 0, 0.0f);
                c->launches++;
                std::swap(xa, xb);
                xchg(0, xa);
            }
            // xa holds the result
            if (sl)     // the target is the local descriptor when level 1 is the first replicated one
                k_mg0_restrict_list<<<sl_rblocks(up_desc(0)), 256, 0, st>>>(m0, up_desc(0), c->vr, xa, seg_of(1, Lc == 1),
                                                                           cnt_of(1, Lc == 1), c->dS);
            else
                k_mg0_restrict<<<cdiv(up_desc(0).n, TPB), TPB, 0, st>>>(m0, up_desc(0), c->vr, xa, c->dS);
            c->launches++;
            ps->lv[0].x = xa; ps->lv[0].x2 = xb;
            after_restrict(0);
        }
        for (int l = 1; l < fs && l < L - 1; l++) {
            MgLevel &lv = *LV[l];
            float *xa = lv.x, *xb = lv.x2;
            for (int sw = 0; sw < nu; sw++) {
                sweep_lvl(l, lv, lv, xa, nullptr, xb, pre_om(sw), sw == 0 ? 0 : 1);
                std::swap(xa, xb);
                xchg(l, xa);
            }
            lv.x = xa; lv.x2 = xb;
            if (sl)
                k_mg_restrict_list<<<sl_rblocks(up_desc(l)), 256, 0, st>>>(lv, up_desc(l), lv.x, seg_of(l + 1, l + 1 == Lc),
                                                                          cnt_of(l + 1, l + 1 == Lc), c->dS);
            else
                k_mg_restrict<<<cdiv(up_desc(l).n, TPB), TPB, 0, st>>>(lv, up_desc(l), lv.x, c->dS);
            c->launches++;
            after_restrict(l);
        }
        // ---- bottom: small levels in one CTA, or coarsest-level sweeps
        if (fs < L) {
            MgSmallArgs A;
            for (int l = 0; l < L; l++) A.lv[l] = *LV[l];
            A.first = fs; A.last = L - 1; A.p = mp;
            k_mg_small<<<1, 1024, 0, st>>>(A, c->dS); c->launches++;
        } else {
            MgLevel &lv = *LV[L - 1];
            float *xa = lv.x, *xb = lv.x2;
            for (int sw = 0; sw < mp.coarseSweeps; sw++) {
                sweep_lvl(L - 1, lv, lv, xa, nullptr, xb, mp.omegaCoarse, sw == 0 ? 0 : 1);
                std::swap(xa, xb);
            }
            lv.x = xa; lv.x2 = xb;
        }
        // ---- up
        int top = (fs < L ? fs : L - 1) - 1;     // highest-numbered level handled by per-level launches on the way up
        for (int l = top; l >= 1; l--) {
            MgLevel &lv = *LV[l];
            float *xa = lv.x, *xb = lv.x2;
            for (int sw = 0; sw < nu; sw++) {
                sweep_lvl(l, lv, up_desc(l), xa, up_x(l), xb, post_om(sw), sw == 0 ? 2 : 1);
                std::swap(xa, xb);
                xchg(l, xa);      // the next sweep, and the finer level's prolongation, read the neighbours' planes
            }
            lv.x = xa; lv.x2 = xb;
        }
        {
            float *xa = ps->lv[0].x, *xb = ps->lv[0].x2;
            for (int sw = 0; sw < nu; sw++) {
                int last = (sw == nu - 1) ? 1 : 0;
                launch_mg0_sweep(loopBlocks, st, c->segCell, c->segMask, m0, up_desc(0), c->vr, xa, up_x(0), xb, post_om(sw),
                                                      mp.scale, sw == 0 ? 2 : 1, last, c->vz, c->dS, rhoSlot, 0.0f);
                c->launches++;
                std::swap(xa, xb);
                if (!last) xchg(0, xa);
            }
            ps->lv[0].x = xa; ps->lv[0].x2 = xb;
        }
        kt_end(c, FLIP_KERNEL_PRECOND, ktV);
    };

    // Single-GPU V-cycle: the coarse levels run over their active-segment lists, the first two pre-sweeps of every
    // level are one pass (mode 3), restrictions use eight threads per coarse cell, and the single-CTA levels live
    // in shared memory.  Same operator as `vcycle` up to float summation order in the restrictions.
    auto list_blocks = [&](const MgLevel &lv) { return std::max(1, std::min(cdiv(cdiv(lv.n, 32), WPB), 148 * 8)); };
    auto restrict_blocks = [&](const MgLevel &coarse) { return std::max(1, std::min(cdiv(coarse.n, 32), 148 * 8)); };
    // fusedIt >= 0 (needs nu >= 2): the residual update of PCG iteration fusedIt is part of the first pass
    // (k_pcg_update_presweep) and the new residual lands in the ping-pong partner of c->vr
    auto vcycle_list = [&](int rhoSlot, int fusedIt) {
        const int nu = mp.nu;
        size_t ktV = kt_begin(c);
        // ---- down
        {
            float *xa = ps->lv[0].x, *xb = ps->lv[0].x2;
            int sw = 0;
            if (fusedIt >= 0) {
                k_pcg_update_presweep<<<fusedBlocks, TPB, 0, st>>>(c->segCell, c->segMask, m0, c->vs, ps->vq, c->vx_, c->vr, ps->vr2, xb,
                                                                 pre_om(1), pre_om(0), c->dS, fusedIt);
                c->launches++;
                std::swap(c->vr, ps->vr2);
                std::swap(xa, xb);
                sw = 2;
            } else if (nu >= 2) {
                launch_mg0_sweep(loopBlocks, st, c->segCell, c->segMask, m0, ps->lv[1], c->vr, xa, nullptr, xb, pre_om(1),
                                                      mp.scale, 3, 0, nullptr, c->dS, 0, pre_om(0));
                c->launches++;
                std::swap(xa, xb);
                sw = 2;
            }
            for (; sw < nu; sw++) {
                launch_mg0_sweep(loopBlocks, st, c->segCell, c->segMask, m0, ps->lv[1], c->vr, xa, nullptr, xb, pre_om(sw),
                                                      mp.scale, sw == 0 ? 0 : 1, 0, nullptr, c->dS, 0, 0.0f);
                c->launches++;
                std::swap(xa, xb);
            }
            k_mg0_restrict_list<<<restrict_blocks(ps->lv[1]), 256, 0, st>>>(m0, ps->lv[1], c->vr, xa, ps->lvSeg[1],
                                                                           &ps->lvSegCount[1], c->dS);
            c->launches++;
            ps->lv[0].x = xa; ps->lv[0].x2 = xb;
        }
        if (ps->coarseSmemBytes) {
            // levels 1..coarsest in one cooperative launch
            MgCoarseArgs A;
            for (int l = 0; l < MG_MAX_LEVELS; l++) { A.lv[l] = ps->lv[l < L ? l : 0]; A.seg[l] = ps->lvSeg[l < L ? l : 0]; }
            A.segCount = ps->lvSegCount;
            A.fs = fs; A.last = L - 1; A.p = mp;
            A.smemFloats = (int)(ps->coarseSmemBytes / sizeof(float));
            A.group = ps->coarseGroup; A.groupBar = ps->groupBar;
            A.trace = ps->trace;
            A.smallInv = ps->smallInv; A.smallInvOk = ps->smallInvOk;
            const DeviceScalars *Sdev = c->dS;
            void *args[] = {&A, &Sdev};
            FLIP_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_mg_coarse, dim3(ps->coarseBlocks), dim3(1024), args,
                                                        ps->coarseSmemBytes, st));
            c->launches++;
            // the kernel ping-pongs x/x2 of the list-driven levels: replay the parity to find the results
            const int swaps = (nu >= 2 ? nu - 1 : nu) + nu;
            if (swaps & 1)
                for (int l = 1; l < fs; l++) std::swap(ps->lv[l].x, ps->lv[l].x2);
        } else {
            for (int l = 1; l < fs && l < L - 1; l++) {
                MgLevel &lv = ps->lv[l];
                float *xa = lv.x, *xb = lv.x2;
                int sw = 0;
                if (nu >= 2) {
                    k_mg_sweep_list<<<list_blocks(lv), TPB, 0, st>>>(lv, lv, xa, nullptr, xb, pre_om(1), mp.scale, 3, ps->lvSeg[l],
                                                                    &ps->lvSegCount[l], c->dS, pre_om(0));
                    c->launches++;
                    std::swap(xa, xb);
                    sw = 2;
                }
                for (; sw < nu; sw++) {
                    k_mg_sweep_list<<<list_blocks(lv), TPB, 0, st>>>(lv, lv, xa, nullptr, xb, pre_om(sw), mp.scale, sw == 0 ? 0 : 1,
                                                                    ps->lvSeg[l], &ps->lvSegCount[l], c->dS, 0.0f);
                    c->launches++;
                    std::swap(xa, xb);
                }
                lv.x = xa; lv.x2 = xb;
                k_mg_restrict_list<<<restrict_blocks(ps->lv[l + 1]), 256, 0, st>>>(lv, ps->lv[l + 1], lv.x, ps->lvSeg[l + 1],
                                                                                  &ps->lvSegCount[l + 1], c->dS);
                c->launches++;
            }
            // ---- bottom: the small levels in one CTA
            {
                MgSmallArgs A;
                for (int l = 0; l < L; l++) A.lv[l] = ps->lv[l];
                A.first = fs; A.last = L - 1; A.p = mp;
                k_mg_small<<<1, 1024, 0, st>>>(A, c->dS);
                c->launches++;
            }
            // ---- up
            for (int l = fs - 1; l >= 1; l--) {
                MgLevel &lv = ps->lv[l];
                float *xa = lv.x, *xb = lv.x2;
                for (int sw = 0; sw < nu; sw++) {
                    k_mg_sweep_list<<<list_blocks(lv), TPB, 0, st>>>(lv, ps->lv[l + 1], xa, ps->lv[l + 1].x, xb, post_om(sw), mp.scale,
                                                                    sw == 0 ? 2 : 1, ps->lvSeg[l], &ps->lvSegCount[l], c->dS, 0.0f);
                    c->launches++;
                    std::swap(xa, xb);
                }
                lv.x = xa; lv.x2 = xb;
            }
        }
        {
            float *xa = ps->lv[0].x, *xb = ps->lv[0].x2;
            for (int sw = 0; sw < nu; sw++) {
                int last = (sw == nu - 1) ? 1 : 0;
                launch_mg0_sweep(loopBlocks, st, c->segCell, c->segMask, m0, ps->lv[1], c->vr, xa, ps->lv[1].x, xb, post_om(sw),
                                                      mp.scale, sw == 0 ? 2 : 1, last, c->vz, c->dS, rhoSlot, 0.0f);
                c->launches++;
                std::swap(xa, xb);
            }
            ps->lv[0].x = xa; ps->lv[0].x2 = xb;
        }
        kt_end(c, FLIP_KERNEL_PRECOND, ktV);
    };
    // the list-driven cycle needs the single-CTA levels to start above level 0 and below the top
    const bool useLists = useMg && !slab && fs >= 1 && fs < L;
    auto apply_precond = [&](int rhoSlot) { if (useLists) vcycle_list(rhoSlot, -1); else vcycle(rhoSlot); };
    // single GPU, multigrid with at least two pre-sweeps: the fused passes (two level-0 passes fewer per iteration)
    const bool fused = useLists && mp.nu >= 2 && !getenv("FLIP_PCG_UNFUSED");

    k_pcg_scalars_init<<<1, 1, 0, st>>>(c->dS, pp.tolFactor); c->launches++;
    if (c->pcgPersistent && !slab) {
        PcgDev P;
        P.segCell = c->segCell; P.segMask = c->segMask; P.g = g; P.factor = pp.factor;
        P.Adiag = c->Adiag; P.oU = c->AoffU; P.oV = c->AoffV; P.oW = c->AoffW;
        P.b = c->vb; P.x = c->vx_; P.r = c->vr; P.s = c->vs; P.z = c->vz;
        P.m0 = m0;
        if (!useMg) { P.m0.g = g; P.m0.fac = 0.f; P.m0.Adiag = c->Adiag; P.m0.oU = c->AoffU; P.m0.oV = c->AoffV; P.m0.oW = c->AoffW; P.m0.invD = ps->lv[0].invD; P.m0.rowBits = ps->maskAll; }
        for (int l = 0; l < MG_MAX_LEVELS; l++) P.lv[l] = ps->lv[l < ps->numLevels ? l : 0];
        P.numLevels = ps->numLevels; P.firstSmall = ps->firstSmall; P.mp = mp;
        P.useMg = useMg ? 1 : 0; P.maxIter = c->pressureMaxIter; P.S = c->dS; P.warm = warm ? 1 : 0;
        if (ps->coopBlocks == 0) {
            int perSm = 0, sms = 0;
            FLIP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_pcg_persistent, TPB, 0));
            FLIP_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
            if (perSm < 1) throw CudaError("k_pcg_persistent does not fit on an SM");
            ps->coopBlocks = perSm * sms;
        }
        void *args[] = {&P};
        size_t ktS = kt_begin(c);
        FLIP_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k_pcg_persistent, dim3(ps->coopBlocks), dim3(TPB), args, 0, st));
        kt_end(c, FLIP_KERNEL_PCG_SOLVE, ktS);
        c->launches++;
        scalars_to_host(c);
    } else {
    if (slab && warm) slab_exchange_vector_halo(c, c->vx_);     // the initial guess on the neighbours' boundary planes
    k_pcg_init<<<segBlocks, TPB, 0, st>>>(c->segCell, c->segMask, pp, c->vb, c->Adiag, c->vr, c->vz, c->vs, c->dS, jacobi,
                                          c->AoffU, c->AoffV, c->AoffW, warm ? c->vx_ : nullptr);
    c->launches++;
    if (useMg) {
        apply_precond(0);
        if (!fused) { k_pcg_copy_zs<<<segBlocks, TPB, 0, st>>>(c->segCell, c->segMask, c->vz, c->vs, c->dS); c->launches++; }
    }
    if (slab) slab_allreduce_scalar(c, &c->dS->rho[0], COMM_SUM_F64);

    int it = 0;
    // the host polls the convergence flag between batches of iterations (a stream sync each): the first batch
    // runs up to just below the previous solve's count (consecutive substeps need nearly the same number)
    const int batch = useMg ? 3 : 16;
    const int firstBatch = (useMg && ps->lastIterations > batch + 2) ? ps->lastIterations - 2 : batch;
    bool done = false;
    if (fused) {
        // iteration it = [dir_spmv(it)] update+presweep(it), rest of the V-cycle, dir_spmv(it+1): the convergence test of
        // iteration it is taken by dir_spmv(it+1), which returns at once when the residual is below the tolerance
        auto dir_spmv = [&](int i) {
            size_t ktSp = kt_begin(c);
            k_pcg_dir_spmv<<<fusedBlocks, TPB, 0, st>>>(c->segCell, c->segMask, pp, c->Adiag, c->AoffU, c->AoffV, c->AoffW, c->vz,
                                                      c->vs, ps->vs2, ps->vq, c->dS, i);
            kt_end(c, FLIP_KERNEL_PCG_DIR_SPMV, ktSp);
            c->launches++;
            std::swap(c->vs, ps->vs2);
        };
        dir_spmv(0);
        while (!done && it < c->pressureMaxIter) {
            int stop = it + (it == 0 ? firstBatch : batch);
            if (stop > c->pressureMaxIter) stop = c->pressureMaxIter;
            for (; it < stop; it++) {
                size_t ktIt = kt_begin(c);
                vcycle_list((it + 1) % 3, it);
                dir_spmv(it + 1);
                kt_end(c, FLIP_KERNEL_PCG_ITER, ktIt);
            }
            scalars_to_host(c);
            done = c->hS->pcgDone != 0;
        }
    } else {
    while (!done && it < c->pressureMaxIter) {
        int stop = it + (it == 0 ? firstBatch : batch);
        if (stop > c->pressureMaxIter) stop = c->pressureMaxIter;
        for (; it < stop; it++) {
            size_t ktIt = kt_begin(c);
            if (slab) slab_exchange_vector_halo(c, c->vs);     // the neighbours' boundary plane of the search vector
            size_t ktSp = kt_begin(c);
            k_pcg_spmv<<<loopBlocks, TPB, 0, st>>>(c->segCell, c->segMask, pp, c->Adiag, c->AoffU, c->AoffV, c->AoffW,
                                                  c->vs, c->vz, c->dS, it);
            kt_end(c, FLIP_KERNEL_PCG_SPMV, ktSp);
            if (slab) slab_allreduce_scalar(c, &c->dS->dotSZ[it % 3], COMM_SUM_F64);
            k_pcg_update<<<loopBlocks, TPB, 0, st>>>(c->segCell, c->segMask, c->Adiag, c->vs, c->vz, c->vx_, c->vr,
                                                    c->dS, it, jacobi);
            c->launches += 2;
            if (slab) slab_allreduce_scalar(c, &c->dS->rMaxBits[it % 3], COMM_MAX_U64);
            if (useMg) apply_precond((it + 1) % 3);
            if (slab) slab_allreduce_scalar(c, &c->dS->rho[(it + 1) % 3], COMM_SUM_F64);
            k_pcg_direction<<<loopBlocks, TPB, 0, st>>>(c->segCell, c->segMask, c->vz, c->vs, c->dS, it);
            kt_end(c, FLIP_KERNEL_PCG_ITER, ktIt);
            c->launches++;
        }
        scalars_to_host(c);
        done = c->hS->pcgDone != 0;
    }
    }
    }   // multi-launch path
    FLIP_CUDA_CHECK(cudaGetLastError());
    if (ps->trace) {
        unsigned long long h[64];
        FLIP_CUDA_CHECK(cudaMemcpy(h, ps->trace, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "k_mg_coarse phase ns:");
        for (int q = 1; q < 64 && h[q]; q++) fprintf(stderr, " %llu", h[q] - h[q - 1]);
        fprintf(stderr, "\n");
    }
    ps->lastIterations = c->hS->pcgIterations;
    c->cur.pcg_iterations = c->hS->pcgIterations;
    c->cur.pcg_error = c->hS->pcgError;
    bool success = c->hS->pcgDone == 1;
    if (success) c->cur.pcg_converged = 1;
    else if (c->hS->pcgIterations == c->pressureMaxIter && c->hS->pcgError < c->pressureAcceptableTol) c->cur.pcg_converged = 2;
    else c->cur.pcg_converged = 0;
    // _solveLinearSystem failure: solve() returns before applying (pressuresolver.cpp:70-72)
    if (c->cur.pcg_converged == 0) {
        // nothing worth a warm start from
        FLIP_CUDA_CHECK(cudaMemsetAsync(ps->maskAll, 0, sizeof(unsigned int) * ps->nSegAll, st));
        return;
    }

    if (slab) slab_exchange_vector_halo(c, c->vx_);     // pressure of the neighbours' boundary planes
    ApplyParams ap;
    ap.g = g;
    ap.factor = (float)(dt / d.dx);
    size_t ktAp = kt_begin(c);
    {
        // (one CTA per row of faces where the row fits: I + 1 = 257 faces would leave a second CTA of 256 with one thread)
        const int aT = std::min(1024, ((d.I + 1 + 31) / 32) * 32);
        k_apply_pressure<0><<<dim3(cdiv(d.I + 1, aT), d.J, d.K), aT, 0, st>>>(ap, c->phiL, c->vx_, c->wU, c->U, c->validU);
        k_apply_pressure<1><<<dim3(cdiv(d.I, aT), d.J + 1, d.K), aT, 0, st>>>(ap, c->phiL, c->vx_, c->wV, c->V, c->validV);
        k_apply_pressure<2><<<dim3(cdiv(d.I, aT), d.J, d.K + 1), aT, 0, st>>>(ap, c->phiL, c->vx_, c->wW, c->W, c->validW);
    }
    kt_end(c, FLIP_KERNEL_PRESSURE_APPLY, ktAp);
    c->launches += 3;
    FLIP_CUDA_CHECK(cudaGetLastError());
}

void pressure_to_float(flip_ctx *c, float *devOut) {
    const Dims &d = c->d;
    PGrid g{d.I, d.J, d.K, d.I, d.I * d.J, d.kOff, d.Kg, d.kOwn0, d.kOwn1, 0};
    k_pressure_to_float<<<cdiv(d.nC, TPB), TPB, 0, c->stream>>>(c->phiL, c->vx_, g, d.nC, devOut);
    c->launches++;
}

}  // namespace flip
