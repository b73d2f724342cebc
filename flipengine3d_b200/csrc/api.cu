// C-ABI of libflip_b200.so (include/flip_b200.h): context lifetime, configuration, the frame /
// substep loop of FluidSimulation::update and the stage dispatcher.
#include <algorithm>
#include <limits>
#include <cmath>
#include <cstring>
#include <new>
#include "flip_internal.h"

using namespace flip;

namespace flip {
void pressure_to_float(flip_ctx *c, float *devOut);
}

static thread_local std::string g_createError;

namespace flip {
size_t kt_begin(flip_ctx *c) {
    if (!c->ktEnabled) return 0;
    if (c->ktUsed + 2 > c->ktPool.size()) {
        size_t old = c->ktPool.size();
        c->ktPool.resize(old + 256);
        for (size_t q = old; q < c->ktPool.size(); q++) FLIP_CUDA_CHECK(cudaEventCreate(&c->ktPool[q]));
    }
    size_t slot = c->ktUsed;
    c->ktUsed += 2;
    FLIP_CUDA_CHECK(cudaEventRecord(c->ktPool[slot], c->stream));
    return slot;
}
void kt_end(flip_ctx *c, int cls, size_t slot) {
    if (!c->ktEnabled) return;
    FLIP_CUDA_CHECK(cudaEventRecord(c->ktPool[slot + 1], c->stream));
    c->ktPending.push_back({cls, slot, slot + 1});
}
void kt_collect(flip_ctx *c) {
    for (auto &p : c->ktPending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ktPool[p.a], c->ktPool[p.b]) == cudaSuccess) {
            c->ktSumMs[p.cls] += ms;
            c->ktCount[p.cls]++;
        }
    }
    c->ktPending.clear();
    c->ktUsed = 0;
}
}  // namespace flip

template <class F>
static int guarded(flip_ctx *c, F &&f) {
    try {
        if (c) cudaSetDevice(c->device);
        f();
        return FLIP_OK;
    } catch (const CudaError &e) {
        if (c) c->lastError = e.msg; else g_createError = e.msg;
        return FLIP_ERR_CUDA;
    } catch (const ApiError &e) {
        if (c) c->lastError = e.msg; else g_createError = e.msg;
        return e.code;
    } catch (const std::bad_alloc &) {
        if (c) c->lastError = "host allocation failed"; else g_createError = "host allocation failed";
        return FLIP_ERR_RUNTIME;
    }
}

template <class T>
static void dev_alloc(T *&p, size_t n, bool zero = true) {
    FLIP_CUDA_CHECK(cudaMalloc(&p, sizeof(T) * (n + 64)));
    if (zero) FLIP_CUDA_CHECK(cudaMemset(p, 0, sizeof(T) * (n + 64)));
}

static void free_grids(flip_ctx *c);

static void free_all(flip_ctx *c) {
    cudaSetDevice(c->device);
    particles_free(c);
    free_grids(c);
    mesher_free(c);
    for (auto &o : c->fluidObjects) cudaFree(o.dsdf);
    cudaFree(c->scanTemp);
    cudaFree(c->nearSolid); cudaFree(c->pressure);
    cudaFree(c->dS);
    cudaFree(c->sendBuf[0]); cudaFree(c->sendBuf[1]); cudaFree(c->occ);
    peer_free(c);
    comm_destroy(c->comm); c->comm = nullptr;
    if (c->hS) cudaFreeHost(c->hS);
    if (c->eventsCreated) for (auto &e : c->evStage) cudaEventDestroy(e);
    for (auto &e : c->ktPool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
}

// Local grid of this context: Klocal planes starting at global plane kOff, of which [kOwn0,kOwn1) are owned.
static void set_geometry(flip_ctx *c, int I, int J, int Klocal, double dx, int kOff, int Kg, int kOwn0, int kOwn1) {
    Dims &d = c->d;
    d.I = I; d.J = J; d.K = Klocal; d.dx = dx;
    d.kOff = kOff; d.Kg = Kg; d.kOwn0 = kOwn0; d.kOwn1 = kOwn1;
    d.nU = (I + 1) * J * Klocal;
    d.nV = I * (J + 1) * Klocal;
    d.nW = I * J * (Klocal + 1);
    d.nC = I * J * Klocal;
    d.nN = (I + 1) * (J + 1) * (Klocal + 1);
}

static void free_grids(flip_ctx *c) {
    pressure_free(c);
    cudaFree(c->cellCount); cudaFree(c->cellStart); cudaFree(c->cellStartA);
    cudaFree(c->U); cudaFree(c->V); cudaFree(c->W); cudaFree(c->sU); cudaFree(c->sV); cudaFree(c->sW);
    cudaFree(c->validU); cudaFree(c->validV); cudaFree(c->validW); cudaFree(c->status);
    cudaFree(c->frontier[0]); cudaFree(c->frontier[1]);
    cudaFree(c->phiL); cudaFree(c->phiS); cudaFree(c->wU); cudaFree(c->wV); cudaFree(c->wW);
    cudaFree(c->wC); cudaFree(c->solU); cudaFree(c->solV); cudaFree(c->solW); cudaFree(c->pocketFlag);
    cudaFree(c->fricU); cudaFree(c->fricV); cudaFree(c->fricW);
    c->fricU = c->fricV = c->fricW = nullptr;
    for (int m = 0; m < 3; m++) {
        cudaFree(c->solidWeightSum[m]); cudaFree(c->solidValid[m]);
        c->solidWeightSum[m] = nullptr; c->solidValid[m] = nullptr;
    }
    c->wC = c->solU = c->solV = c->solW = nullptr;
    c->pocketFlag = nullptr;
    for (auto &a : c->p2gAcc) { cudaFree(a); a = nullptr; }
    cudaFree(c->occBits); c->occBits = nullptr;
    cudaFree(c->p2gTiles); c->p2gTiles = nullptr;
    c->cellCount = c->cellStart = c->cellStartA = nullptr;
    c->U = c->V = c->W = c->sU = c->sV = c->sW = nullptr;
    c->validU = c->validV = c->validW = c->status = nullptr;
    c->frontier[0] = c->frontier[1] = nullptr;
    c->phiL = c->phiS = c->wU = c->wV = c->wW = nullptr;
}

static void allocate_grids(flip_ctx *c) {
    const Dims &d = c->d;
    dev_alloc(c->U, d.nU); dev_alloc(c->V, d.nV); dev_alloc(c->W, d.nW);
    dev_alloc(c->sU, d.nU); dev_alloc(c->sV, d.nV); dev_alloc(c->sW, d.nW);
    dev_alloc(c->validU, d.nU); dev_alloc(c->validV, d.nV); dev_alloc(c->validW, d.nW);
    // extrapolation scratch: level bytes and two frontier lists for each of the three components
    dev_alloc(c->status, 3 * ext_stride(d));
    dev_alloc(c->frontier[0], 3 * ext_stride(d)); dev_alloc(c->frontier[1], 3 * ext_stride(d));
    dev_alloc(c->phiL, d.nC); dev_alloc(c->phiS, d.nN);
    dev_alloc(c->wU, d.nU); dev_alloc(c->wV, d.nV); dev_alloc(c->wW, d.nW);
    dev_alloc(c->cellCount, (size_t)d.nC + 1); dev_alloc(c->cellStart, (size_t)d.nC + 1);
    dev_alloc(c->cellStartA, (size_t)d.nC + 1);
    // cell occupancy bitmaps (occupied / 3x3x3 / 5x5x5 neighbourhood / surface), one bit per cell, rows padded to whole words
    dev_alloc(c->occBits, 4 * (size_t)((d.I + 31) / 32) * d.J * d.K);
    c->occBitsValid = false;
    pressure_alloc(c);
    // phi_liquid starts at the "no particles" value 3dx (particlelevelset.cpp:295-301)
    std::vector<float> init((size_t)d.nC, (float)(3.0 * d.dx));
    FLIP_CUDA_CHECK(cudaMemcpy(c->phiL, init.data(), sizeof(float) * d.nC, cudaMemcpyHostToDevice));
}

// Static inputs.  The solid SDF handed in (or built) is GLOBAL; a z-slab keeps its planes only.
static void upload_static_inputs(flip_ctx *c) {
    const Dims &d = c->d;
    Dims gd = d;   // global geometry
    gd.K = d.Kg; gd.kOff = 0;
    gd.nN = (d.I + 1) * (d.J + 1) * (d.Kg + 1);
    gd.nC = d.I * d.J * d.Kg;
    FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));      // (the copies below run on the legacy stream)
    if (!c->userSolidPhi) build_box_solid_sdf(gd, c->hostSolidPhi);
    // the domain with the enabled obstacles merged in (_addStaticObjectsToSDF, fluidsimulation.cpp:2927-2975)
    const std::vector<float> *solid = &c->hostSolidPhi;
    std::vector<float> merged;
    for (auto &o : c->obstacles) {
        if (!o.enabled) continue;
        if (merged.empty()) { merged = c->hostSolidPhi; solid = &merged; }
        for (size_t q = 0; q < merged.size(); q++) merged[q] = std::min(merged[q], o.sdf[q]);
    }
    c->solidDirty = false;
    std::vector<unsigned char> ns;
    build_near_solid(gd, *solid, c->nearSolidFactor, c->solidExactBand, c->CFL, ns, c->nsI, c->nsJ, c->nsK);
    const size_t nodePlane = (size_t)(d.I + 1) * (d.J + 1);
    std::vector<float> local(solid->begin() + nodePlane * d.kOff, solid->begin() + nodePlane * (d.kOff + d.K + 1));
    std::vector<float> wU, wV, wW, wC;
    build_weights(d, local, wU, wV, wW, wC);
    FLIP_CUDA_CHECK(cudaMemcpy(c->phiS, local.data(), sizeof(float) * d.nN, cudaMemcpyHostToDevice));
    FLIP_CUDA_CHECK(cudaMemcpy(c->wU, wU.data(), sizeof(float) * d.nU, cudaMemcpyHostToDevice));
    FLIP_CUDA_CHECK(cudaMemcpy(c->wV, wV.data(), sizeof(float) * d.nV, cudaMemcpyHostToDevice));
    FLIP_CUDA_CHECK(cudaMemcpy(c->wW, wW.data(), sizeof(float) * d.nW, cudaMemcpyHostToDevice));
    if (c->solU) {      // moving solids: the centre weights multiply their velocities in the divergence
        build_center_weights(d, local, wC);
        if (!c->wC) dev_alloc(c->wC, d.nC);
        FLIP_CUDA_CHECK(cudaMemcpy(c->wC, wC.data(), sizeof(float) * d.nC, cudaMemcpyHostToDevice));
    }
    if (!c->userFaceFriction) {
        // face friction of the merged solids (0 everywhere, the default: no arrays, the constraint takes its short path)
        bool any = c->boundaryFriction != 0.0f;
        for (auto &o : c->obstacles) any = any || (o.enabled && o.friction != 0.0f);
        if (any && !slab_on(c)) {
            std::vector<FrictionSolid> solids;
            solids.push_back({&c->hostSolidPhi, c->boundaryFriction});
            for (int pass = 0; pass < 2; pass++)            // static obstacles first, then the animated ones (:3062-3063)
                for (auto &o : c->obstacles)
                    if (o.enabled && o.animated == (pass == 1)) solids.push_back({&o.sdf, o.friction});
            std::vector<float> fU, fV, fW;
            build_face_friction(d, c->solidExactBand, solids, fU, fV, fW);
            if (!c->fricU) { dev_alloc(c->fricU, d.nU); dev_alloc(c->fricV, d.nV); dev_alloc(c->fricW, d.nW); }
            FLIP_CUDA_CHECK(cudaMemcpy(c->fricU, fU.data(), sizeof(float) * d.nU, cudaMemcpyHostToDevice));
            FLIP_CUDA_CHECK(cudaMemcpy(c->fricV, fV.data(), sizeof(float) * d.nV, cudaMemcpyHostToDevice));
            FLIP_CUDA_CHECK(cudaMemcpy(c->fricW, fW.data(), sizeof(float) * d.nW, cudaMemcpyHostToDevice));
        } else if (c->fricU) {
            cudaFree(c->fricU); cudaFree(c->fricV); cudaFree(c->fricW);
            c->fricU = c->fricV = c->fricW = nullptr;
        }
    }
    cudaFree(c->nearSolid); c->nearSolid = nullptr;
    dev_alloc(c->nearSolid, ns.size());
    FLIP_CUDA_CHECK(cudaMemcpy(c->nearSolid, ns.data(), ns.size(), cudaMemcpyHostToDevice));
}

extern "C" {

const char *flip_create_error(void) { return g_createError.c_str(); }

int flip_create(flip_ctx **out, int isize, int jsize, int ksize, double dx, int device) {
    if (!out) return FLIP_ERR_RUNTIME;
    *out = nullptr;
    // FluidSimulation::FluidSimulation: dims and dx must be positive (fluidsimulation.cpp:44-56)
    if (isize <= 0 || jsize <= 0 || ksize <= 0 || !(dx > 0.0)) {
        g_createError = "Error: dimensions and cell size must be greater than 0.";
        return FLIP_ERR_DOMAIN;
    }
    if ((long long)(isize + 1) * (jsize + 1) * (long long)(ksize + 1) > 2000000000ll) {
        g_createError = "grid too large for 32-bit cell indices";
        return FLIP_ERR_DOMAIN;
    }
    flip_ctx *c = new (std::nothrow) flip_ctx();
    if (!c) return FLIP_ERR_RUNTIME;
    c->device = device;
    int rc = guarded(nullptr, [&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CudaError(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                            "): libflip_b200 has no CPU fallback");
        if (device < 0 || device >= ndev) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "bad CUDA device ordinal");
        FLIP_CUDA_CHECK(cudaSetDevice(device));
        c->KgCfg = ksize;
        set_geometry(c, isize, jsize, ksize, dx, 0, ksize, 0, ksize);
        c->liquidRadius = 0.5 * 1.0 * dx * sqrt(3.0);   // _initializeParticleRadii fluidsimulation.cpp:2649
        FLIP_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto &e2 : c->evStage) FLIP_CUDA_CHECK(cudaEventCreate(&e2));
        c->eventsCreated = true;
        dev_alloc(c->dS, 1);
        FLIP_CUDA_CHECK(cudaMallocHost(&c->hS, sizeof(DeviceScalars)));
        memset(c->hS, 0, sizeof(DeviceScalars));
        allocate_grids(c);
    });
    if (rc != FLIP_OK) {
        free_all(c);
        delete c;
        return rc;
    }
    *out = c;
    return FLIP_OK;
}

void flip_destroy(flip_ctx *c) {
    if (!c) return;
    free_all(c);
    delete c;
}

const char *flip_last_error(const flip_ctx *c) { return c ? c->lastError.c_str() : "null context"; }

int flip_reset_body_force(flip_ctx *c) {       // resetBodyForce  fluidsimulation.cpp:1598
    return guarded(c, [&] { c->gravity[0] = c->gravity[1] = c->gravity[2] = 0.0; });
}
int flip_set_extreme_velocity_removal(flip_ctx *c, int on) {      // enable / disableExtremeVelocityRemoval  :1869-1881
    return guarded(c, [&] { c->extremeVelocityRemoval = on != 0; });
}
int flip_set_marker_particle_scale(flip_ctx *c, double s) {       // setMarkerParticleScale  :168-179
    return guarded(c, [&] {
        if (s < 0.0) throw ApiError(FLIP_ERR_DOMAIN, "Error: marker particle scale must be greater than or equal to 0.");
        if (s != c->markerParticleScale) c->stepCounter++;          // a cached mesh no longer applies
        c->markerParticleScale = s;
    });
}
int flip_add_body_force(flip_ctx *c, double fx, double fy, double fz) {
    // _constantBodyForces is summed by _getConstantBodyForce (fluidsimulation.cpp:3436-3444), in vec3 floats
    return guarded(c, [&] {
        c->gravity[0] = (double)((float)c->gravity[0] + (float)fx);
        c->gravity[1] = (double)((float)c->gravity[1] + (float)fy);
        c->gravity[2] = (double)((float)c->gravity[2] + (float)fz);
    });
}

int flip_set_pic_flip_ratio(flip_ctx *c, double r) {
    return guarded(c, [&] {
        if (r < 0.0 || r > 1.0) throw ApiError(FLIP_ERR_DOMAIN, "Error: PICFLIP ratio must be in range [0.0, 1.0].");
        c->ratioPICFLIP = r;
    });
}
int flip_set_cfl(flip_ctx *c, double cfl) {
    return guarded(c, [&] {
        if (cfl < 1.0) throw ApiError(FLIP_ERR_DOMAIN, "Error: CFL must be greater than or equal to 1.");
        const double oldCfl = c->CFL;
        const int oldLayers = c->extrapolationLayers;
        c->CFL = cfl;
        c->extrapolationLayers = (int)ceil(cfl) + 2;
        if (slab_on(c) && c->halo < slab_required_halo(c)) {
            c->CFL = oldCfl; c->extrapolationLayers = oldLayers;
            throw ApiError(FLIP_ERR_DOMAIN, "Error: this CFL number needs more z-slab halo planes than flip_set_halo configured "
                                            "(ceil(CFL) + 1 + extrapolation layers + 1).");
        }
    });
}
int flip_set_substep_limits(flip_ctx *c, int mn, int mx) {
    return guarded(c, [&] {
        if (mn < 1 || mx < mn || mx > 8) throw ApiError(FLIP_ERR_DOMAIN, "Error: bad time steps per frame range.");
        c->minSubsteps = mn; c->maxSubsteps = mx;
    });
}
int flip_set_pressure_solver(flip_ctx *c, double tol, double acc, int maxIter) {
    return guarded(c, [&] {
        if (!(tol > 0) || maxIter < 1) throw ApiError(FLIP_ERR_DOMAIN, "Error: bad pressure solver parameters.");
        c->pressureTol = tol; c->pressureAcceptableTol = acc; c->pressureMaxIter = maxIter;
    });
}
int flip_set_preconditioner(flip_ctx *c, int kind) {
    return guarded(c, [&] {
        if (kind < 0 || kind > 1) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "unknown preconditioner");
        c->preconditioner = kind;
    });
}

int flip_set_multigrid(flip_ctx *c, int sweeps, double damping, double weight, int coarsest) {
    return guarded(c, [&] {
        if (sweeps < 1 || sweeps > 8 || !(damping > 0.0 && damping <= 1.0) || !(weight > 0.0) || coarsest < 1)
            throw ApiError(FLIP_ERR_DOMAIN, "Error: bad multigrid parameters.");
        c->mgNu = sweeps; c->mgOmega = damping; c->mgScale = weight; c->mgCoarseSweeps = coarsest;
        for (double &w : c->mgOmegaSched) w = damping;
    });
}

int flip_set_multigrid_schedule(flip_ctx *c, int sweeps, const double *damping) {
    if (!c) return FLIP_ERR_RUNTIME;
    return guarded(c, [&] {
        if (sweeps < 1 || sweeps > 8 || !damping) throw ApiError(FLIP_ERR_DOMAIN, "Error: bad multigrid parameters.");
        for (int q = 0; q < sweeps; q++)
            if (!(damping[q] > 0.0 && damping[q] < 2.0)) throw ApiError(FLIP_ERR_DOMAIN, "Error: bad multigrid parameters.");
        c->mgNu = sweeps;
        for (int q = 0; q < 8; q++) c->mgOmegaSched[q] = damping[q < sweeps ? q : sweeps - 1];
    });
}

int flip_set_pressure_warm_start(flip_ctx *c, int on) {
    if (!c) return FLIP_ERR_RUNTIME;
    return guarded(c, [&] { c->pressureWarmStart = on ? 1 : 0; });
}

int flip_set_sampling_mode(flip_ctx *c, int mode) {
    if (!c) return FLIP_ERR_RUNTIME;
    return guarded(c, [&] {
        if (mode != FLIP_SAMPLING_EXACT && mode != FLIP_SAMPLING_FAST) throw ApiError(FLIP_ERR_DOMAIN, "sampling mode must be FLIP_SAMPLING_EXACT or FLIP_SAMPLING_FAST");
        c->samplingMode = mode;
    });
}

int flip_set_solver_mode(flip_ctx *c, int persistent) {
    return guarded(c, [&] { c->pcgPersistent = persistent ? 1 : 0; });
}

int flip_load_particles(flip_ctx *c, int n, const float *pos, const float *vel) {
    return guarded(c, [&] {
        if (n < 0) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "negative particle count");
        if (n == 0) return;
        c->loadQueuePos.insert(c->loadQueuePos.end(), pos, pos + 3ll * n);
        c->loadQueueVel.insert(c->loadQueueVel.end(), vel, vel + 3ll * n);
    });
}

static flip_ctx::FluidObject make_fluid_object(flip_ctx *c, const float *sdf, const int cellLo[3], const int cellHi[3], const double lo[3],
                                              const double hi[3], const double vel[3]) {
    flip_ctx::FluidObject o;
    const int n[3] = {c->d.I, c->d.J, c->d.Kg};
    for (int a = 0; a < 3; a++) {
        o.vel[a] = vel ? vel[a] : 0.0;
        if (sdf) {
            o.boxLo[a] = o.boxHi[a] = 0.0;
            o.lo[a] = cellLo ? std::max(0, cellLo[a]) : 0;
            o.hi[a] = cellHi ? std::min(n[a], cellHi[a]) : n[a];
        } else {
            o.boxLo[a] = lo[a]; o.boxHi[a] = hi[a];
            // the cells that can have a corner node inside the box
            o.lo[a] = std::max(0, (int)floor(lo[a] / c->d.dx) - 1);
            o.hi[a] = std::min(n[a], (int)ceil(hi[a] / c->d.dx) + 1);
        }
    }
    if (sdf) o.sdf.assign(sdf, sdf + (size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1));
    return o;
}

// The level-set grid a MeshFluidSource keeps for itself (MeshFluidSource::update, meshfluidsource.cpp:178-198): the
// bounding box of the mesh grown by _gridpad = 4 cell widths in total (AABB::expand, aabb.cpp:122-128: half on either
// side), in cells [gmin, gmax] clamped to the domain, with gmax - gmin + 2 nodes per axis.  The velocity constraint of
// an inflow reads this grid at the WORLD position (no offset subtracted, fluidsimulation.cpp:3401,4138 -- seeding does
// subtract it, :4549), and nodes outside the grid read as 0: stage_inflow_* reproduce exactly that.
static void source_level_set_grid(flip_ctx *c, flip_ctx::FluidObject &o, const double meshLo[3], const double meshHi[3]) {
    const int n[3] = {c->d.I, c->d.J, c->d.Kg};
    const double dx = c->d.dx, invdx = 1.0 / dx;
    const double pad = 4.0 * dx;
    for (int a = 0; a < 3; a++) {
        const float vmin = (float)meshLo[a], vmax = (float)meshHi[a];                   // mesh vertices are floats
        double width = (double)vmax - (double)vmin + 1e-9;                              // AABB(points), aabb.cpp:54-82
        const float pmin = vmin - (float)(0.5 * pad);                                   // position -= vec3(h, h, h)
        width += pad;
        const float pmax = pmin + (float)width;                                         // getMaxPoint
        int gmin = (int)floor((double)pmin * invdx), gmax = (int)floor((double)pmax * invdx);
        gmin = std::max(gmin, 0); gmax = std::min(gmax, n[a] - 1);
        o.sdfOrigin[a] = gmin;
        o.sdfNodes[a] = std::max(gmax - gmin + 1, 1) + 1;
    }
}

int flip_add_fluid_box(flip_ctx *c, const double lo[3], const double hi[3], const double vel[3]) {
    return guarded(c, [&] { c->fluidObjects.push_back(make_fluid_object(c, nullptr, nullptr, nullptr, lo, hi, vel)); });
}

int flip_add_fluid_sdf(flip_ctx *c, const float *sdf, const int cellLo[3], const int cellHi[3], const double vel[3]) {
    return guarded(c, [&] {
        if (!sdf) throw ApiError(FLIP_ERR_RUNTIME, "null signed distance field");
        c->fluidObjects.push_back(make_fluid_object(c, sdf, cellLo, cellHi, nullptr, nullptr, vel));
    });
}

int flip_add_fluid_source_box(flip_ctx *c, int outflow, const double lo[3], const double hi[3], const double vel[3], int *id) {
    return guarded(c, [&] {
        flip_ctx::FluidObject o = make_fluid_object(c, nullptr, nullptr, nullptr, lo, hi, vel);
        source_level_set_grid(c, o, lo, hi);
        o.kind = outflow ? 2 : 1; o.id = c->nextSourceId++;
        if (id) *id = o.id;
        c->fluidObjects.push_back(std::move(o));
    });
}

int flip_add_fluid_source_sdf(flip_ctx *c, int outflow, const float *sdf, const int cellLo[3], const int cellHi[3],
                              const double meshLo[3], const double meshHi[3], const double vel[3], int *id) {
    return guarded(c, [&] {
        if (!sdf) throw ApiError(FLIP_ERR_RUNTIME, "null signed distance field");
        flip_ctx::FluidObject o = make_fluid_object(c, sdf, cellLo, cellHi, nullptr, nullptr, vel);
        if (meshLo && meshHi) source_level_set_grid(c, o, meshLo, meshHi);
        else {
            // no mesh bounds given: the box of the nodes inside the object
            const int ni = c->d.I + 1, nj = c->d.J + 1, nk = c->d.Kg + 1;
            double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300};
            for (int k = 0; k < nk; k++)
                for (int j = 0; j < nj; j++)
                    for (int i = 0; i < ni; i++)
                        if (sdf[(size_t)i + (size_t)ni * (j + (size_t)nj * k)] <= 0.0f) {
                            const int q[3] = {i, j, k};
                            for (int a = 0; a < 3; a++) { blo[a] = std::min(blo[a], q[a] * c->d.dx); bhi[a] = std::max(bhi[a], q[a] * c->d.dx); }
                        }
            if (blo[0] > bhi[0]) for (int a = 0; a < 3; a++) blo[a] = bhi[a] = 0.0;
            source_level_set_grid(c, o, blo, bhi);
        }
        o.kind = outflow ? 2 : 1; o.id = c->nextSourceId++;
        if (id) *id = o.id;
        c->fluidObjects.push_back(std::move(o));
    });
}

int flip_enable_fluid_source(flip_ctx *c, int id, int on) {
    return guarded(c, [&] {
        for (auto &o : c->fluidObjects)
            if (o.kind != 0 && o.id == id) { o.enabled = on != 0; return; }
        throw ApiError(FLIP_ERR_RUNTIME, "Error: could not find mesh fluid source to remove.");
    });
}

int flip_constrain_fluid_source_velocity(flip_ctx *c, int id, int on) {
    return guarded(c, [&] {
        for (auto &o : c->fluidObjects)
            if (o.kind != 0 && o.id == id) { o.constrained = on != 0; return; }
        throw ApiError(FLIP_ERR_RUNTIME, "Error: could not find mesh fluid source.");
    });
}

int flip_remove_fluid_source(flip_ctx *c, int id) {
    return guarded(c, [&] {
        for (size_t q = 0; q < c->fluidObjects.size(); q++)
            if (c->fluidObjects[q].kind != 0 && c->fluidObjects[q].id == id) {
                FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
                cudaFree(c->fluidObjects[q].dsdf);
                c->fluidObjects.erase(c->fluidObjects.begin() + q);
                return;
            }
        throw ApiError(FLIP_ERR_RUNTIME, "Error: could not find mesh fluid source to remove.");      // :1979-1982
    });
}

int flip_add_marker_particle(flip_ctx *c, const float p[3], const float v[3]) {
    return guarded(c, [&] {
        // before initialize the particle joins the load queue; afterwards it is appended to the store
        if (!c->initialized) {
            c->loadQueuePos.insert(c->loadQueuePos.end(), p, p + 3);
            c->loadQueueVel.insert(c->loadQueueVel.end(), v, v + 3);
            return;
        }
        int n = c->np;
        std::vector<float> aos((size_t)6 * (n + 1));
        particles_download_aos(c, aos.data());
        float *a = aos.data() + 6ll * n;
        a[0] = p[0]; a[1] = p[1]; a[2] = p[2]; a[3] = v[0]; a[4] = v[1]; a[5] = v[2];
        particles_upload_aos(c, aos.data(), n + 1);
    });
}

int flip_set_solid_sdf(flip_ctx *c, const float *phi) {
    return guarded(c, [&] {
        // always the GLOBAL nodal array, (I+1)(J+1)(Kglobal+1) floats
        size_t nGlobal = (size_t)(c->d.I + 1) * (c->d.J + 1) * (c->d.Kg + 1);
        c->hostSolidPhi.assign(phi, phi + nGlobal);
        c->userSolidPhi = true;
        if (c->initialized) upload_static_inputs(c);   // re-derive the static inputs
    });
}

// ---- static obstacles ---------------------------------------------------------------------------
static void obstacles_changed(flip_ctx *c) {
    if (c->initialized) c->solidDirty = true;      // picked up by the next substep's obstacle stage (_isSolidLevelSetUpToDate = false)
    c->stepCounter++;                              // a cached surface mesh was clamped against the old solid
}

int flip_add_obstacle_sdf(flip_ctx *c, const float *sdf, int *id) {
    return guarded(c, [&] {
        if (!sdf) throw ApiError(FLIP_ERR_RUNTIME, "null signed distance field");
        flip_ctx::Obstacle o;
        o.id = c->nextObstacleId++;
        o.sdf.assign(sdf, sdf + (size_t)(c->d.I + 1) * (c->d.J + 1) * (c->d.Kg + 1));
        if (id) *id = o.id;
        c->obstacles.push_back(std::move(o));
        obstacles_changed(c);
    });
}

// the distances MeshLevelSet computes for a box mesh: exact within the band of solidExactBand cells around the mesh's
// index box (meshlevelset.cpp:572-601), untouched (here: the largest float) elsewhere
static void box_obstacle_sdf(const flip_ctx *c, const double lo[3], const double hi[3], std::vector<float> &sdf) {
    const int n[3] = {c->d.I + 1, c->d.J + 1, c->d.Kg + 1};
    const double dx = c->d.dx;
    sdf.assign((size_t)n[0] * n[1] * n[2], std::numeric_limits<float>::max());
    int a0[3], a1[3];
    for (int a = 0; a < 3; a++) {
        a0[a] = std::max(0, (int)floor(lo[a] / dx) - c->solidExactBand);
        a1[a] = std::min(n[a] - 1, (int)ceil(hi[a] / dx) + c->solidExactBand);
    }
    const float cx[3] = {(float)(0.5 * (lo[0] + hi[0])), (float)(0.5 * (lo[1] + hi[1])), (float)(0.5 * (lo[2] + hi[2]))};
    const float hx[3] = {(float)(0.5 * (hi[0] - lo[0])), (float)(0.5 * (hi[1] - lo[1])), (float)(0.5 * (hi[2] - lo[2]))};
    for (int k = a0[2]; k <= a1[2]; k++)
        for (int j = a0[1]; j <= a1[1]; j++)
            for (int i = a0[0]; i <= a1[0]; i++) {
                const float p[3] = {(float)(i * dx), (float)(j * dx), (float)(k * dx)};
                float q[3], out2 = 0.0f, in = -std::numeric_limits<float>::max();
                for (int a = 0; a < 3; a++) {
                    q[a] = std::fabs(p[a] - cx[a]) - hx[a];
                    const float e = std::max(q[a], 0.0f);
                    out2 += e * e;
                    in = std::max(in, q[a]);
                }
                sdf[(size_t)i + (size_t)n[0] * (j + (size_t)n[1] * k)] = std::sqrt(out2) + std::min(in, 0.0f);
            }
}

int flip_add_obstacle_box(flip_ctx *c, const double lo[3], const double hi[3], int *id) {
    return guarded(c, [&] {
        flip_ctx::Obstacle o;
        o.id = c->nextObstacleId++;
        o.isBox = true;
        for (int a = 0; a < 3; a++) { o.lo[a] = lo[a]; o.hi[a] = hi[a]; }
        box_obstacle_sdf(c, lo, hi, o.sdf);
        if (id) *id = o.id;
        c->obstacles.push_back(std::move(o));
        obstacles_changed(c);
    });
}

// MeshObject::updateMeshAnimated(previous, current, next) (meshobject.cpp:61-95) for a box that translates rigidly: the
// three meshes are the box of flip_add_obstacle_box moved by the three offsets.  From then on the obstacle stage of every
// substep (_updateSolidLevelSet, fluidsimulation.cpp:3028-3071, which rebuilds the solid SDF while an animated mesh
// changes, :3002) places the box at current + t (next - current), t = the part of the frame completed when the substep
// starts (:2892-2893, MeshObject::getMesh(t) meshobject.cpp:158-177), gives it the velocity
// ((current - previous) + t ((next - current) - (current - previous))) / frame dt (getVertexVelocities :199-215), and
// re-derives solid SDF, weights, near-solid mask and the solids' face velocities.
int flip_set_obstacle_box_motion(flip_ctx *c, int id, const double offPrev[3], const double offCur[3], const double offNext[3]) {
    return guarded(c, [&] {
        if (!offPrev || !offCur || !offNext) throw ApiError(FLIP_ERR_RUNTIME, "null offset");
        if (slab_on(c)) throw ApiError(FLIP_ERR_UNSUPPORTED, "animated obstacles are not supported in a z-slab run");
        for (auto &o : c->obstacles)
            if (o.id == id) {
                if (!o.isBox) throw ApiError(FLIP_ERR_UNSUPPORTED, "only obstacles added with flip_add_obstacle_box can be animated");
                o.animated = true;
                for (int a = 0; a < 3; a++) { o.offPrev[a] = offPrev[a]; o.offCur[a] = offCur[a]; o.offNext[a] = offNext[a]; }
                return;
            }
        throw ApiError(FLIP_ERR_RUNTIME, "Error: could not find mesh obstacle.");
    });
}

// Friction: setBoundaryFriction (fluidsimulation.cpp:1747-1759) and MeshObject::setFriction (meshobject.cpp:300-304) of an
// obstacle, both in [0, 1]; the face friction of the constraint is re-derived with the static inputs.
int flip_set_boundary_friction(flip_ctx *c, double f) {
    return guarded(c, [&] {
        if (!(f >= 0.0 && f <= 1.0)) throw ApiError(FLIP_ERR_DOMAIN, "Error: boundary friction must be in range [0.0, 1.0].");
        if (slab_on(c) && f != 0.0) throw ApiError(FLIP_ERR_UNSUPPORTED, "friction is not supported in a z-slab run");
        c->boundaryFriction = (float)f;
        obstacles_changed(c);
    });
}
int flip_set_obstacle_friction(flip_ctx *c, int id, double f) {
    return guarded(c, [&] {
        if (slab_on(c) && f != 0.0) throw ApiError(FLIP_ERR_UNSUPPORTED, "friction is not supported in a z-slab run");
        for (auto &o : c->obstacles)
            if (o.id == id) {
                o.friction = (float)std::fmax(std::fmin(f, 1.0), 0.0);        // clamped, as setFriction does
                obstacles_changed(c);
                return;
            }
        throw ApiError(FLIP_ERR_RUNTIME, "Error: could not find mesh obstacle.");
    });
}
// The face friction itself (parity seam, and for callers that keep their own solids): three HOST arrays in the MAC layout,
// copied; three NULLs: back to the friction derived from the boundary and the obstacles.
int flip_set_face_friction(flip_ctx *c, const float *U, const float *V, const float *W) {
    return guarded(c, [&] {
        const Dims &d = c->d;
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (!U && !V && !W) {
            c->userFaceFriction = false;
            cudaFree(c->fricU); cudaFree(c->fricV); cudaFree(c->fricW);
            c->fricU = c->fricV = c->fricW = nullptr;
            obstacles_changed(c);
            return;
        }
        if (!U || !V || !W) throw ApiError(FLIP_ERR_RUNTIME, "flip_set_face_friction: three arrays or three null pointers");
        if (slab_on(c)) throw ApiError(FLIP_ERR_UNSUPPORTED, "friction is not supported in a z-slab run");
        if (!c->fricU) { dev_alloc(c->fricU, d.nU); dev_alloc(c->fricV, d.nV); dev_alloc(c->fricW, d.nW); }
        FLIP_CUDA_CHECK(cudaMemcpy(c->fricU, U, sizeof(float) * d.nU, cudaMemcpyHostToDevice));
        FLIP_CUDA_CHECK(cudaMemcpy(c->fricV, V, sizeof(float) * d.nV, cudaMemcpyHostToDevice));
        FLIP_CUDA_CHECK(cudaMemcpy(c->fricW, W, sizeof(float) * d.nW, cudaMemcpyHostToDevice));
        c->userFaceFriction = true;
    });
}
int flip_get_face_friction(flip_ctx *c, float *U, float *V, float *W) {
    return guarded(c, [&] {
        const Dims &d = c->d;
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        float *out[3] = {U, V, W};
        const float *src[3] = {c->fricU, c->fricV, c->fricW};
        const int n[3] = {d.nU, d.nV, d.nW};
        for (int m = 0; m < 3; m++) {
            if (!out[m]) continue;
            if (src[m]) FLIP_CUDA_CHECK(cudaMemcpy(out[m], src[m], sizeof(float) * n[m], cudaMemcpyDeviceToHost));
            else memset(out[m], 0, sizeof(float) * n[m]);
        }
    });
}

// FluidSimulation::addMeshObstacle with a closed triangle mesh the library keeps (so that it can be animated):
// vertices as xyz triplets, triangles as vertex index triplets.  Static until flip_set_obstacle_mesh_motion is called.
int flip_add_obstacle_mesh(flip_ctx *c, const float *vertices, int numVertices, const int *triangles, int numTriangles, int *id) {
    return guarded(c, [&] {
        if (!vertices || !triangles || numVertices <= 0 || numTriangles <= 0) throw ApiError(FLIP_ERR_RUNTIME, "empty mesh");
        flip_ctx::Obstacle o;
        o.id = c->nextObstacleId++;
        o.isMesh = true;
        o.vertsCur.assign(vertices, vertices + 3 * (size_t)numVertices);
        o.vertsPrev = o.vertsCur;
        o.vertsNext = o.vertsCur;
        o.triangles.assign(triangles, triangles + 3 * (size_t)numTriangles);
        o.sdf.resize((size_t)(c->d.I + 1) * (c->d.J + 1) * (c->d.Kg + 1));
        const int rc = flip_mesh_sdf(c->d.I, c->d.J, c->d.Kg, c->d.dx, vertices, numVertices, triangles, numTriangles, c->solidExactBand,
                                     std::numeric_limits<float>::max(), o.sdf.data(), nullptr, nullptr);
        if (rc != FLIP_OK) throw ApiError(rc, "bad triangle mesh");
        if (id) *id = o.id;
        c->obstacles.push_back(std::move(o));
        obstacles_changed(c);
    });
}

// MeshObject::updateMeshAnimated(previous, current, next) (meshobject.cpp:61-95) for an obstacle added with
// flip_add_obstacle_mesh: the vertices of the three frames (same count and order: fixed topology).  Per substep the mesh
// stands at current + t (next - current) and vertex v moves with ((cur - prev)_v + t ((next - cur)_v - (cur - prev)_v)) /
// frame dt; the solids' face velocities take the velocity of the nearest surface point (flip::mesh_velocity_data).
int flip_set_obstacle_mesh_motion(flip_ctx *c, int id, const float *prev, const float *cur, const float *next) {
    return guarded(c, [&] {
        if (!prev || !cur || !next) throw ApiError(FLIP_ERR_RUNTIME, "null vertex array");
        if (slab_on(c)) throw ApiError(FLIP_ERR_UNSUPPORTED, "animated obstacles are not supported in a z-slab run");
        for (auto &o : c->obstacles)
            if (o.id == id) {
                if (!o.isMesh) throw ApiError(FLIP_ERR_UNSUPPORTED, "only obstacles added with flip_add_obstacle_mesh take vertex animations");
                const size_t n = o.vertsCur.size();
                o.vertsPrev.assign(prev, prev + n);
                o.vertsCur.assign(cur, cur + n);
                o.vertsNext.assign(next, next + n);
                o.animated = true;
                return;
            }
        throw ApiError(FLIP_ERR_RUNTIME, "Error: could not find mesh obstacle.");
    });
}

// The obstacle stage with animated obstacles (see flip_set_obstacle_box_motion).  Host work per substep: the SDF of every
// moving box and the solid fractions of all solids; the device normalises and extrapolates the face velocities.
static void update_animated_obstacles(flip_ctx *c) {
    bool any = false;
    for (auto &o : c->obstacles) any = any || (o.enabled && o.animated);
    if (!any) {
        if (c->solidVelFromAnimation) {         // the last animated obstacle was disabled or removed: solids at rest again
            FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            cudaFree(c->solU); cudaFree(c->solV); cudaFree(c->solW);
            c->solU = c->solV = c->solW = nullptr;
            c->solidVelFromAnimation = false;
        }
        return;
    }
    if (slab_on(c)) throw ApiError(FLIP_ERR_UNSUPPORTED, "animated obstacles are not supported in a z-slab run");
    c->solidVelFromAnimation = true;
    const Dims &d = c->d;
    // fluidsimulation.cpp:2892-2893 (floats)
    const float frameTime = (float)(c->frameRemaining + c->substepDt);
    float t = 1.0f - frameTime / (float)c->frameDt;
    t = std::fmin(1.0f, std::fmax(0.0f, t));
    const double invdt = c->frameDt < 1e-10 ? 0.0 : 1.0 / c->frameDt;
    // per moving mesh: its solid fractions and fraction x nearest-surface velocity on every face
    struct MeshData { const flip_ctx::Obstacle *o; std::vector<float> fraction[3], field[3]; };
    std::vector<MeshData> meshData;
    for (auto &o : c->obstacles) {
        if (!o.enabled || !o.animated || !o.isMesh) continue;
        const size_t n3 = o.vertsCur.size();
        std::vector<float> vt(n3), vel(n3);
        for (size_t q = 0; q < n3; q++) {
            const float p0 = o.vertsPrev[q], p1 = o.vertsCur[q], p2 = o.vertsNext[q];
            vt[q] = p1 + t * (p2 - p1);
            const float m1 = p1 - p0, m2 = p2 - p1;
            vel[q] = (float)((double)(m1 + t * (m2 - m1)) * invdt);
        }
        meshData.emplace_back();
        meshData.back().o = &o;
        const int rc = mesh_velocity_data(d, vt.data(), (int)(n3 / 3), o.triangles.data(), (int)(o.triangles.size() / 3), vel.data(),
                                          c->solidExactBand, std::numeric_limits<float>::max(), o.sdf, meshData.back().fraction,
                                          meshData.back().field);
        if (rc != FLIP_OK) throw ApiError(rc, "bad animated triangle mesh");
    }
    for (auto &o : c->obstacles) {
        if (!o.enabled || !o.animated || o.isMesh) continue;
        double lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            // the lower corner vertex of the three meshes as floats, then the reference's float interpolation
            const float p0 = (float)(o.lo[a] + o.offPrev[a]), p1 = (float)(o.lo[a] + o.offCur[a]), p2 = (float)(o.lo[a] + o.offNext[a]);
            const float pt = p1 + t * (p2 - p1);
            const float m1 = p1 - p0, m2 = p2 - p1;
            o.velocity[a] = (float)((double)(m1 + t * (m2 - m1)) * invdt);
            lo[a] = pt;
            hi[a] = (double)pt + (o.hi[a] - o.lo[a]);
        }
        box_obstacle_sdf(c, lo, hi, o.sdf);
    }
    upload_static_inputs(c);        // merged solid SDF, face weights, near-solid mask (and the centre weights below)
    // the velocity data of the merged solid SDF: weights of all solids, values of the moving ones
    std::vector<float> weightSum[3], fieldSum[3];
    const int n[3] = {d.nU, d.nV, d.nW};
    for (int m = 0; m < 3; m++) { weightSum[m].assign(n[m], 0.0f); fieldSum[m].assign(n[m], 0.0f); }
    add_solid_fractions(d, c->hostSolidPhi, nullptr, weightSum, fieldSum);
    for (auto &o : c->obstacles)
        if (o.enabled && !(o.animated && o.isMesh)) add_solid_fractions(d, o.sdf, o.animated ? o.velocity : nullptr, weightSum, fieldSum);
    for (auto &md : meshData)
        for (int m = 0; m < 3; m++)
            for (int q = 0; q < n[m]; q++) { weightSum[m][q] += md.fraction[m][q]; fieldSum[m][q] += md.field[m][q]; }
    if (!c->solU) {
        dev_alloc(c->solU, d.nU); dev_alloc(c->solV, d.nV); dev_alloc(c->solW, d.nW);
        if (!c->pocketFlag) dev_alloc(c->pocketFlag, d.nC);     // (kept when the velocities are withdrawn)
    }
    float *sol[3] = {c->solU, c->solV, c->solW};
    for (int m = 0; m < 3; m++) {
        if (!c->solidWeightSum[m]) { dev_alloc(c->solidWeightSum[m], n[m]); dev_alloc(c->solidValid[m], n[m]); }
        FLIP_CUDA_CHECK(cudaMemcpy(sol[m], fieldSum[m].data(), sizeof(float) * n[m], cudaMemcpyHostToDevice));
        FLIP_CUDA_CHECK(cudaMemcpy(c->solidWeightSum[m], weightSum[m].data(), sizeof(float) * n[m], cudaMemcpyHostToDevice));
    }
    if (!c->wC) {
        std::vector<float> phi((size_t)d.nN), wC;
        FLIP_CUDA_CHECK(cudaMemcpy(phi.data(), c->phiS, sizeof(float) * d.nN, cudaMemcpyDeviceToHost));
        build_center_weights(d, phi, wC);
        dev_alloc(c->wC, d.nC);
        FLIP_CUDA_CHECK(cudaMemcpy(c->wC, wC.data(), sizeof(float) * d.nC, cudaMemcpyHostToDevice));
    }
    solid_velocity_normalize_extrapolate(c, 5);      // _numVelocityExtrapolationLayers, meshlevelset.h:364
    c->stepCounter++;
}

int flip_enable_obstacle(flip_ctx *c, int id, int on) {
    return guarded(c, [&] {
        for (auto &o : c->obstacles)
            if (o.id == id) {
                if (o.enabled != (on != 0)) { o.enabled = on != 0; obstacles_changed(c); }
                return;
            }
        throw ApiError(FLIP_ERR_RUNTIME, "Error: could not find mesh obstacle.");
    });
}

int flip_remove_obstacle(flip_ctx *c, int id) {
    return guarded(c, [&] {
        for (size_t q = 0; q < c->obstacles.size(); q++)
            if (c->obstacles[q].id == id) {
                c->obstacles.erase(c->obstacles.begin() + q);
                obstacles_changed(c);
                return;
            }
        throw ApiError(FLIP_ERR_DOMAIN, "Error: could not find mesh obstacle to remove.");      // std::invalid_argument, :2020-2024
    });
}

// Face velocities of the solids.  The reference keeps them with the solid SDF (VelocityDataGrid, meshlevelset.h:69-87) and
// fills them from the vertex velocities of animated meshes; here the caller hands in the three face arrays.
int flip_set_solid_velocity(flip_ctx *c, const float *U, const float *V, const float *W) {
    return guarded(c, [&] {
        const Dims &d = c->d;
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (!U && !V && !W) {       // every solid at rest again
            cudaFree(c->solU); cudaFree(c->solV); cudaFree(c->solW);
            c->solU = c->solV = c->solW = nullptr;
            return;
        }
        if (!U || !V || !W) throw ApiError(FLIP_ERR_RUNTIME, "flip_set_solid_velocity: three arrays or three null pointers");
        if (slab_on(c)) throw ApiError(FLIP_ERR_UNSUPPORTED, "solid velocities are not supported in a z-slab run");
        if (!c->solU) {
            dev_alloc(c->solU, d.nU); dev_alloc(c->solV, d.nV); dev_alloc(c->solW, d.nW);
            if (!c->pocketFlag) dev_alloc(c->pocketFlag, d.nC);     // (kept when the velocities are withdrawn)
        }
        FLIP_CUDA_CHECK(cudaMemcpy(c->solU, U, sizeof(float) * d.nU, cudaMemcpyHostToDevice));
        FLIP_CUDA_CHECK(cudaMemcpy(c->solV, V, sizeof(float) * d.nV, cudaMemcpyHostToDevice));
        FLIP_CUDA_CHECK(cudaMemcpy(c->solW, W, sizeof(float) * d.nW, cudaMemcpyHostToDevice));
        if (c->initialized && !c->wC) {
            // the centre weights of the solid SDF as it stands on the device
            std::vector<float> phi((size_t)d.nN), wC;
            FLIP_CUDA_CHECK(cudaMemcpy(phi.data(), c->phiS, sizeof(float) * d.nN, cudaMemcpyDeviceToHost));
            build_center_weights(d, phi, wC);
            dev_alloc(c->wC, d.nC);
            FLIP_CUDA_CHECK(cudaMemcpy(c->wC, wC.data(), sizeof(float) * d.nC, cudaMemcpyHostToDevice));
        }
    });
}

int flip_initialize(flip_ctx *c) {
    return guarded(c, [&] {
        if (c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "Error: FluidSimulation is already initialized.");
        upload_static_inputs(c);
        // _loadParticles (fluidsimulation.cpp:2791); queued fluid objects are seeded at the end of the first substep,
        // as the reference's _updateFluidObjects does (:5504)
        int n = (int)(c->loadQueuePos.size() / 3);
        c->nextParticleId = c->particleIdBase + n;
        // (a z-slab keeps the particles of its own planes only; every rank may be handed the whole scene)
        particles_upload_split(c, c->loadQueuePos.data(), c->loadQueueVel.data(), n);
        c->loadQueuePos.clear(); c->loadQueuePos.shrink_to_fit();
        c->loadQueueVel.clear(); c->loadQueueVel.shrink_to_fit();
        c->initialized = true;
    });
}

// ---- frame / substep bookkeeping (fluidsimulation.cpp:5768-5824) -------------------------------

static double max_particle_speed(flip_ctx *c) {
    // _getMaximumMarkerParticleSpeed (:5553-5565): sqrt(max dot(v,v)); the max over survivors is
    // maintained by the sort's gather kernel
    unsigned int bits = c->hS->maxSpeedSqBits;
    float f;
    memcpy(&f, &bits, sizeof(float));
    return sqrt((double)f);
}

static double next_time_step(flip_ctx *c, double dt) {
    // _calculateNextTimeStep (:5593-5613)
    double maxu;
    if (c->currentFrame == 0 && c->substepNumber == 0) {
        // _predictMaximumMarkerParticleSpeed (:5529-5551): the largest fluid velocity of the queued objects
        // (_getMaximumMeshObjectFluidVelocity of a static object: |velocity| as a float vec3 length) plus the
        // body-force term |g| * dt, with |g| the float vec3 length
        maxu = 0.0;
        for (auto &o : c->fluidObjects) {
            if (o.kind == 2 || !o.enabled) continue;       // queued objects and enabled inflow sources (:5536-5545)
            float vx = (float)o.vel[0], vy = (float)o.vel[1], vz = (float)o.vel[2];
            maxu = std::max(maxu, (double)sqrtf(vx * vx + vy * vy + vz * vz));
        }
        float gx = (float)c->gravity[0], gy = (float)c->gravity[1], gz = (float)c->gravity[2];
        float len = sqrtf(gx * gx + gy * gy + gz * gz);
        maxu += (double)len * dt;
    } else {
        maxu = max_particle_speed(c);
    }
    double eps = 1e-6;
    return c->CFL * c->d.dx / (maxu + eps);
}

int flip_begin_frame(flip_ctx *c, double dt) {
    return guarded(c, [&] {
        if (!c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "Error: FluidSimulation must be initialized before update.");
        if (dt < 0.0) throw ApiError(FLIP_ERR_DOMAIN, "Error: delta time must be greater than or equal to 0.");
        double epsdt = 1e-6;
        dt = std::max(dt, epsdt);
        c->frameDt = dt;
        c->frameRemaining = dt;
        c->substepNumber = 0;
        c->stats.clear();
    });
}

int flip_begin_substep(flip_ctx *c, double *out) {
    return guarded(c, [&] {
        double substepTime = c->frameDt / (double)c->minSubsteps;
        double step = fmin(next_time_step(c, c->frameDt), c->frameRemaining);
        double timeCompleted = c->frameDt - c->frameRemaining;
        double stepLimit = (c->substepNumber + 1) * substepTime;
        if (timeCompleted + step > stepLimit) step = fmin(substepTime, c->frameRemaining);
        if (c->substepNumber == c->maxSubsteps - 1) step = c->frameRemaining;
        c->frameRemaining -= step;
        c->substepDt = step;
        memset(&c->cur, 0, sizeof(c->cur));
        c->cur.dt = step;
        c->cur.pcg_converged = 1;
        if (out) *out = step;
    });
}

static void run_stage(flip_ctx *c, int stage, double dt) {
    FLIP_CUDA_CHECK(cudaEventRecord(c->evStage[stage], c->stream));
    switch (stage) {
        case FLIP_STAGE_OBSTACLES:
            update_animated_obstacles(c);                       // every substep while an animated obstacle is enabled (:3002)
            if (c->solidDirty) upload_static_inputs(c);         // static solids: only after a change (:2007)
            break;
        case FLIP_STAGE_LIQUID_SDF: stage_liquid_sdf(c); break;
        case FLIP_STAGE_P2G: stage_p2g(c); break;
        case FLIP_STAGE_EXTRAPOLATE_A: if (c->np_global > 0 || c->npStore > 0) stage_extrapolate(c); break;   // :3262 guards on !empty()
        case FLIP_STAGE_SAVE: stage_save(c); break;
        case FLIP_STAGE_BODY_FORCE: stage_body_force(c, dt); stage_inflow_body_force_exclusion(c); break;
        case FLIP_STAGE_PRESSURE:
            stage_pressure(c, dt);
            if (slab_on(c)) {
                // projected velocities and masks of the halo planes come from the slabs that own them
                const Dims &d = c->d;
                slab_exchange_planes(c, c->U, (d.I + 1) * d.J, 0);
                slab_exchange_planes(c, c->V, d.I * (d.J + 1), 0);
                slab_exchange_planes(c, c->W, d.I * d.J, 1);
                slab_exchange_planes_u8(c, c->validU, (d.I + 1) * d.J, 0);
                slab_exchange_planes_u8(c, c->validV, d.I * (d.J + 1), 0);
                slab_exchange_planes_u8(c, c->validW, d.I * d.J, 1);
            }
            break;
        case FLIP_STAGE_EXTRAPOLATE_B: stage_extrapolate(c); break;
        case FLIP_STAGE_CONSTRAIN: stage_constrain(c); break;
        case FLIP_STAGE_G2P:
            if (c->fuseAdvance && !has_constrained_inflow(c) && stage_g2p_advance_fused(c, dt)) c->fusedAdvanceDone = true;
            else { stage_g2p(c); stage_inflow_constrain_particles(c); }
            break;
        case FLIP_STAGE_ADVANCE: stage_advance(c, dt); break;
        case FLIP_STAGE_TAIL: stage_fluid_objects(c); break;       // _updateFluidObjects  :5504
        default: throw ApiError(FLIP_ERR_OUT_OF_RANGE, "bad stage id");
    }
    FLIP_CUDA_CHECK(cudaEventRecord(c->evStage[stage + 1], c->stream));
}

int flip_run_stage(flip_ctx *c, int stage, double dt) {
    return guarded(c, [&] {
        if (!c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "Error: FluidSimulation must be initialized before update.");
        if (stage < 0 || stage >= FLIP_NUM_STAGES) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "bad stage id");
        run_stage(c, stage, dt);
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        kt_collect(c);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->evStage[stage], c->evStage[stage + 1]);
        c->stageMs[stage] = ms;
    });
}

int flip_end_substep(flip_ctx *c, int *more) {
    return guarded(c, [&] {
        c->cur.particles = slab_on(c) ? c->np_global : c->np;
        c->cur.removed_solid = c->hS->removedSolid;
        c->cur.removed_crowded = c->hS->removedCrowded;
        c->cur.removed_fast = c->hS->removedFast;
        c->stats.push_back(c->cur);
        c->substepNumber++;
        if (more) *more = c->frameRemaining > 1e-9 ? 1 : 0;
    });
}

int flip_end_frame(flip_ctx *c) {
    return guarded(c, [&] { c->currentFrame++; });
}

int flip_update(flip_ctx *c, double dt) {
    int rc = flip_begin_frame(c, dt);
    if (rc) return rc;
    return guarded(c, [&] {
        int more = 1;
        while (more) {
            double step = 0;
            int r2 = flip_begin_substep(c, &step);
            if (r2) throw ApiError(r2, c->lastError);
            // all stages are enqueued back to back; the only host syncs are the scalar read-backs
            // inside the pressure stage and at the end of the advance stage
            c->fuseAdvance = true;          // G2P and RK3 run as one pass where that applies (stage_g2p_advance_fused)
            for (int s = 0; s < FLIP_NUM_STAGES; s++) run_stage(c, s, step);
            c->fuseAdvance = false;
            FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            kt_collect(c);
            for (int s = 0; s < FLIP_NUM_STAGES; s++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, c->evStage[s], c->evStage[s + 1]);
                c->stageMs[s] = ms;
            }
            r2 = flip_end_substep(c, &more);
            if (r2) throw ApiError(r2, c->lastError);
        }
        c->currentFrame++;
    });
}

int flip_get_current_frame(const flip_ctx *c, int *f) { if (!c || !f) return FLIP_ERR_RUNTIME; *f = c->currentFrame; return FLIP_OK; }
int flip_set_current_frame(flip_ctx *c, int f) {
    return guarded(c, [&] {
        if (f < 0) throw ApiError(FLIP_ERR_DOMAIN, "frame number must be >= 0");
        c->currentFrame = f;
    });
}
int flip_get_num_substeps(const flip_ctx *c, int *n) { if (!c || !n) return FLIP_ERR_RUNTIME; *n = (int)c->stats.size(); return FLIP_OK; }
int flip_get_step_stats(const flip_ctx *c, int s, flip_step_stats *out) {
    if (!c || !out) return FLIP_ERR_RUNTIME;
    if (s < 0 || s >= (int)c->stats.size()) return FLIP_ERR_OUT_OF_RANGE;
    *out = c->stats[s];
    return FLIP_OK;
}
int flip_get_num_particles(const flip_ctx *c, int *n) { if (!c || !n) return FLIP_ERR_RUNTIME; *n = c->np; return FLIP_OK; }

int flip_get_particles(flip_ctx *c, float *aos, int capacity) {
    return guarded(c, [&] {
        if (capacity < c->np) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "Error: particle buffer too small.");
        particles_download_aos(c, aos);
    });
}
int flip_set_particles(flip_ctx *c, int n, const float *aos) {
    return guarded(c, [&] {
        if (!c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "Error: FluidSimulation must be initialized first.");
        if (n < 0) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "negative particle count");
        particles_upload_aos(c, aos, n);
    });
}
int flip_get_particle_positions(flip_ctx *c, float *xyz, int capacity) {
    return guarded(c, [&] {
        if (capacity < c->np) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "Error: particle buffer too small.");
        particles_download_component(c, xyz, 0);
    });
}
int flip_get_particle_velocities(flip_ctx *c, float *xyz, int capacity) {
    return guarded(c, [&] {
        if (capacity < c->np) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "Error: particle buffer too small.");
        particles_download_component(c, xyz, 1);
    });
}

int flip_enable_particle_ids(flip_ctx *c, int on) {
    return guarded(c, [&] { c->trackIds = on != 0; });
}
int flip_set_particle_id_base(flip_ctx *c, int base) {
    return guarded(c, [&] { c->particleIdBase = base; });
}
int flip_get_particle_ids(flip_ctx *c, int32_t *ids, int capacity) {
    return guarded(c, [&] {
        if (!c->trackIds) throw ApiError(FLIP_ERR_RUNTIME, "particle ids are not enabled");
        if (capacity < c->np) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "Error: particle buffer too small.");
        particles_download_ids(c, (int *)ids);
    });
}

int flip_set_surface_subdivision_level(flip_ctx *c, int n) {
    return guarded(c, [&] {
        if (n < 1) throw ApiError(FLIP_ERR_DOMAIN, "Error: subdivision level must be greater than or equal to 1.");
        if (n != c->surfaceSubdivision) c->stepCounter++;       // a cached mesh no longer applies
        c->surfaceSubdivision = n;
    });
}
int flip_set_surface_smoothing(flip_ctx *c, double value, int iterations) {
    return guarded(c, [&] {
        if (iterations < 0) throw ApiError(FLIP_ERR_DOMAIN, "Error: smoothing iterations must be greater than or equal to 0.");
        c->surfaceSmoothingValue = value; c->surfaceSmoothingIterations = iterations;
        c->stepCounter++;
    });
}
int flip_get_isomesh_size(flip_ctx *c, int *numVertices, int *numTriangles) {
    return guarded(c, [&] {
        if (!c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "Error: FluidSimulation must be initialized before its surface exists.");
        mesher_get(c, numVertices, numTriangles, nullptr, nullptr);
    });
}
int flip_get_isomesh(flip_ctx *c, float *verticesXyz, int *triangles) {
    return guarded(c, [&] {
        if (!c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "Error: FluidSimulation must be initialized before its surface exists.");
        mesher_get(c, nullptr, nullptr, verticesXyz, triangles);
    });
}

int flip_get_isomesh_field(flip_ctx *c, float *values, unsigned char *inside, unsigned char *need) {
    return guarded(c, [&] {
        if (!c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "Error: FluidSimulation must be initialized before its surface exists.");
        mesher_debug_field(c, values, inside, need);
    });
}

int flip_get_velocity_field(flip_ctx *c, float *U, float *V, float *W) {
    return guarded(c, [&] {
        const Dims &d = c->d;
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (U) FLIP_CUDA_CHECK(cudaMemcpy(U, c->U, sizeof(float) * d.nU, cudaMemcpyDeviceToHost));
        if (V) FLIP_CUDA_CHECK(cudaMemcpy(V, c->V, sizeof(float) * d.nV, cudaMemcpyDeviceToHost));
        if (W) FLIP_CUDA_CHECK(cudaMemcpy(W, c->W, sizeof(float) * d.nW, cudaMemcpyDeviceToHost));
    });
}

static void *array_ptr(const flip_ctx *c, int which, int64_t *bytes) {
    const Dims &d = c->d;
    switch (which) {
        case FLIP_ARRAY_U: *bytes = 4ll * d.nU; return c->U;
        case FLIP_ARRAY_V: *bytes = 4ll * d.nV; return c->V;
        case FLIP_ARRAY_W: *bytes = 4ll * d.nW; return c->W;
        case FLIP_ARRAY_VALID_U: *bytes = d.nU; return c->validU;
        case FLIP_ARRAY_VALID_V: *bytes = d.nV; return c->validV;
        case FLIP_ARRAY_VALID_W: *bytes = d.nW; return c->validW;
        case FLIP_ARRAY_LIQUID_PHI: *bytes = 4ll * d.nC; return c->phiL;
        case FLIP_ARRAY_SOLID_PHI: *bytes = 4ll * d.nN; return c->phiS;
        case FLIP_ARRAY_WEIGHT_U: *bytes = 4ll * d.nU; return c->wU;
        case FLIP_ARRAY_WEIGHT_V: *bytes = 4ll * d.nV; return c->wV;
        case FLIP_ARRAY_WEIGHT_W: *bytes = 4ll * d.nW; return c->wW;
        case FLIP_ARRAY_WEIGHT_C: if (!c->wC) break; *bytes = 4ll * d.nC; return c->wC;
        case FLIP_ARRAY_SOLID_VEL_U: *bytes = 4ll * d.nU; return c->solU;
        case FLIP_ARRAY_SOLID_VEL_V: *bytes = 4ll * d.nV; return c->solV;
        case FLIP_ARRAY_SOLID_VEL_W: *bytes = 4ll * d.nW; return c->solW;
        case FLIP_ARRAY_SAVED_U: *bytes = 4ll * d.nU; return c->sU;
        case FLIP_ARRAY_SAVED_V: *bytes = 4ll * d.nV; return c->sV;
        case FLIP_ARRAY_SAVED_W: *bytes = 4ll * d.nW; return c->sW;
        case FLIP_ARRAY_NEAR_SOLID: *bytes = (int64_t)c->nsI * c->nsJ * c->nsK; return c->nearSolid;
        case FLIP_ARRAY_PRESSURE: *bytes = 4ll * d.nC; return nullptr;
    }
    *bytes = -1;
    return nullptr;
}

int flip_array_bytes(const flip_ctx *c, int which, int64_t *bytes) {
    if (!c || !bytes) return FLIP_ERR_RUNTIME;
    array_ptr(c, which, bytes);
    if (*bytes < 0) return (which == FLIP_ARRAY_WEIGHT_C) ? FLIP_ERR_UNSUPPORTED : FLIP_ERR_OUT_OF_RANGE;
    return FLIP_OK;
}

int flip_get_array(flip_ctx *c, int which, void *out) {
    return guarded(c, [&] {
        int64_t bytes = 0;
        void *p = array_ptr(c, which, &bytes);
        if (bytes < 0) {
            if (which == FLIP_ARRAY_WEIGHT_C)
                throw ApiError(FLIP_ERR_UNSUPPORTED, "the cell-centre weight only multiplies solid velocities; it is built once flip_set_solid_velocity has been called");
            throw ApiError(FLIP_ERR_OUT_OF_RANGE, "bad array id");
        }
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (which == FLIP_ARRAY_PRESSURE) {
            float *tmp = nullptr;
            FLIP_CUDA_CHECK(cudaMalloc(&tmp, bytes));
            pressure_to_float(c, tmp);
            FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            FLIP_CUDA_CHECK(cudaMemcpy(out, tmp, bytes, cudaMemcpyDeviceToHost));
            cudaFree(tmp);
            return;
        }
        if (!p) throw ApiError(FLIP_ERR_RUNTIME, "array not allocated yet (call flip_initialize)");
        FLIP_CUDA_CHECK(cudaMemcpy(out, p, bytes, cudaMemcpyDeviceToHost));
    });
}

int flip_set_array(flip_ctx *c, int which, const void *in) {
    return guarded(c, [&] {
        int64_t bytes = 0;
        void *p = array_ptr(c, which, &bytes);
        if (bytes < 0 || !p || which == FLIP_ARRAY_PRESSURE) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "array cannot be set");
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        FLIP_CUDA_CHECK(cudaMemcpy(p, in, bytes, cudaMemcpyHostToDevice));
    });
}

int flip_get_stage_times_ms(const flip_ctx *c, float ms[FLIP_NUM_STAGES]) {
    if (!c || !ms) return FLIP_ERR_RUNTIME;
    for (int s = 0; s < FLIP_NUM_STAGES; s++) ms[s] = c->stageMs[s];
    return FLIP_OK;
}
int flip_enable_kernel_timing(flip_ctx *c, int on) {
    return guarded(c, [&] { FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream)); kt_collect(c); c->ktEnabled = on != 0; });
}
int flip_reset_kernel_timing(flip_ctx *c) {
    return guarded(c, [&] {
        FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        kt_collect(c);
        for (int q = 0; q < FLIP_NUM_KERNEL_CLASSES; q++) { c->ktSumMs[q] = 0; c->ktCount[q] = 0; }
    });
}
int flip_get_kernel_timing(const flip_ctx *c, int cls, double *ms, int64_t *n) {
    if (!c || cls < 0 || cls >= FLIP_NUM_KERNEL_CLASSES) return FLIP_ERR_OUT_OF_RANGE;
    if (ms) *ms = c->ktSumMs[cls];
    if (n) *n = c->ktCount[cls];
    return FLIP_OK;
}
int flip_get_kernel_launches(const flip_ctx *c, int64_t *n) { if (!c || !n) return FLIP_ERR_RUNTIME; *n = c->launches; return FLIP_OK; }
int flip_get_stream(const flip_ctx *c, void **s) { if (!c || !s) return FLIP_ERR_RUNTIME; *s = (void *)c->stream; return FLIP_OK; }
int flip_synchronize(flip_ctx *c) {
    return guarded(c, [&] { FLIP_CUDA_CHECK(cudaStreamSynchronize(c->stream)); });
}

int flip_set_slab(flip_ctx *c, int rank, int nranks, const void *id, int idBytes) {
    return guarded(c, [&] {
        if (c->initialized) throw ApiError(FLIP_ERR_RUNTIME, "flip_set_slab must precede flip_initialize");
        if (nranks < 1 || rank < 0 || rank >= nranks) throw ApiError(FLIP_ERR_OUT_OF_RANGE, "bad rank / nranks");
        if (nranks == 1) return;
        const int Kg = c->KgCfg, H = c->halo;
        int k0 = 0, k1 = 0;
        flip_slab_range(Kg, nranks, rank, &k0, &k1);
        if ((Kg / nranks) < H) throw ApiError(FLIP_ERR_DOMAIN, "z-slabs thinner than the halo width");
        if (H < slab_required_halo(c))
            throw ApiError(FLIP_ERR_DOMAIN, "z-slab halo narrower than the particle reach plus the extrapolation depth "
                                            "(ceil(CFL) + 1 + extrapolation layers + 1 planes): call flip_set_halo first");
        const int lo = rank > 0 ? H : 0, hi = rank < nranks - 1 ? H : 0;
        c->comm = comm_create(rank, nranks, id, idBytes);
        c->rank = rank; c->nranks = nranks;
        free_grids(c);
        set_geometry(c, c->d.I, c->d.J, (k1 - k0) + lo + hi, c->d.dx, k0 - lo, Kg, lo, lo + (k1 - k0));
        allocate_grids(c);
        peer_setup(c);
    });
}
int flip_static_inputs(int I, int J, int K, double dx, float *phi, int phiIsInput, float *wU, float *wV, float *wW,
                       unsigned char *nearSolid, int nearDims[3]) {
    if (I <= 0 || J <= 0 || K <= 0 || !(dx > 0.0)) return FLIP_ERR_DOMAIN;
    try {
        flip_ctx defaults;          // only its configuration defaults are read
        Dims d;
        d.I = I; d.J = J; d.K = K; d.dx = dx; d.kOff = 0; d.Kg = K; d.kOwn0 = 0; d.kOwn1 = K;
        d.nU = (I + 1) * J * K; d.nV = I * (J + 1) * K; d.nW = I * J * (K + 1); d.nC = I * J * K;
        d.nN = (I + 1) * (J + 1) * (K + 1);
        std::vector<float> p, a, b, w, cc;
        if (phiIsInput) {
            if (!phi) return FLIP_ERR_RUNTIME;
            p.assign(phi, phi + (size_t)d.nN);
        } else {
            build_box_solid_sdf(d, p);
            if (phi) memcpy(phi, p.data(), sizeof(float) * p.size());
        }
        if (wU || wV || wW) {
            build_weights(d, p, a, b, w, cc);
            if (wU) memcpy(wU, a.data(), sizeof(float) * a.size());
            if (wV) memcpy(wV, b.data(), sizeof(float) * b.size());
            if (wW) memcpy(wW, w.data(), sizeof(float) * w.size());
        }
        if (nearSolid || nearDims) {
            std::vector<unsigned char> ns;
            int gi = 0, gj = 0, gk = 0;
            build_near_solid(d, p, defaults.nearSolidFactor, defaults.solidExactBand, defaults.CFL, ns, gi, gj, gk);
            if (nearSolid) memcpy(nearSolid, ns.data(), ns.size());
            if (nearDims) { nearDims[0] = gi; nearDims[1] = gj; nearDims[2] = gk; }
        }
        return FLIP_OK;
    } catch (const std::bad_alloc &) { g_createError = "host allocation failed"; return FLIP_ERR_RUNTIME; }
}
// The face friction of a set of solids, on the HOST (no device needed): solids[0] is the domain (nodal field everywhere),
// the others obstacle fields carrying the largest float outside their band, in merge order.
int flip_face_friction(int I, int J, int K, double dx, int band, int numSolids, const float *const *phis, const float *frictions,
                       float *fU, float *fV, float *fW) {
    if (I <= 0 || J <= 0 || K <= 0 || !(dx > 0.0) || numSolids <= 0 || !phis || !frictions) return FLIP_ERR_DOMAIN;
    try {
        Dims d;
        d.I = I; d.J = J; d.K = K; d.dx = dx; d.kOff = 0; d.Kg = K; d.kOwn0 = 0; d.kOwn1 = K;
        d.nU = (I + 1) * J * K; d.nV = I * (J + 1) * K; d.nW = I * J * (K + 1); d.nC = I * J * K;
        d.nN = (I + 1) * (J + 1) * (K + 1);
        std::vector<std::vector<float>> fields(numSolids);
        std::vector<FrictionSolid> solids;
        for (int q = 0; q < numSolids; q++) {
            fields[q].assign(phis[q], phis[q] + (size_t)d.nN);
            solids.push_back({&fields[q], frictions[q]});
        }
        std::vector<float> a, b, w;
        build_face_friction(d, band, solids, a, b, w);
        if (fU) memcpy(fU, a.data(), sizeof(float) * a.size());
        if (fV) memcpy(fV, b.data(), sizeof(float) * b.size());
        if (fW) memcpy(fW, w.data(), sizeof(float) * w.size());
        return FLIP_OK;
    } catch (const std::bad_alloc &) { g_createError = "host allocation failed"; return FLIP_ERR_RUNTIME; }
}
// The nodal field the library gives a box obstacle (flip_add_obstacle_box), on the HOST: exact box distances within `band`
// cells of the box's index range, the largest float elsewhere.
int flip_box_obstacle_sdf(int I, int J, int K, double dx, int band, const double lo[3], const double hi[3], float *phi) {
    if (I <= 0 || J <= 0 || K <= 0 || !(dx > 0.0) || !lo || !hi || !phi) return FLIP_ERR_DOMAIN;
    try {
        flip_ctx tmp;       // geometry and band only
        tmp.d.I = I; tmp.d.J = J; tmp.d.K = K; tmp.d.Kg = K; tmp.d.dx = dx;
        tmp.solidExactBand = band;
        std::vector<float> sdf;
        box_obstacle_sdf(&tmp, lo, hi, sdf);
        memcpy(phi, sdf.data(), sizeof(float) * sdf.size());
        return FLIP_OK;
    } catch (const std::bad_alloc &) { g_createError = "host allocation failed"; return FLIP_ERR_RUNTIME; }
}
int flip_center_weights(int I, int J, int K, double dx, const float *phi, float *wC) {
    if (I <= 0 || J <= 0 || K <= 0 || !(dx > 0.0)) return FLIP_ERR_DOMAIN;
    if (!phi || !wC) return FLIP_ERR_RUNTIME;
    try {
        Dims d;
        d.I = I; d.J = J; d.K = K; d.dx = dx; d.kOff = 0; d.Kg = K; d.kOwn0 = 0; d.kOwn1 = K;
        d.nC = I * J * K;
        d.nN = (I + 1) * (J + 1) * (K + 1);
        std::vector<float> p(phi, phi + (size_t)d.nN), cc;
        build_center_weights(d, p, cc);
        memcpy(wC, cc.data(), sizeof(float) * cc.size());
        return FLIP_OK;
    } catch (const std::bad_alloc &) { g_createError = "host allocation failed"; return FLIP_ERR_RUNTIME; }
}
int flip_get_nccl_unique_id(void *out, int idBytes) {
    try {
        comm_unique_id(out, idBytes);
        return FLIP_OK;
    } catch (const CudaError &e) { g_createError = e.msg; return FLIP_ERR_CUDA; }
    catch (const ApiError &e) { g_createError = e.msg; return e.code; }
}
/* owned global planes of rank `rank`: a balanced split of K */
int flip_slab_range(int K, int nranks, int rank, int *k0, int *k1) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !k0 || !k1) return FLIP_ERR_OUT_OF_RANGE;
    *k0 = (int)((long long)K * rank / nranks);
    *k1 = (int)((long long)K * (rank + 1) / nranks);
    return FLIP_OK;
}
int flip_get_slab_info(const flip_ctx *c, int *kOff, int *Klocal, int *kOwn0, int *kOwn1) {
    if (!c) return FLIP_ERR_RUNTIME;
    if (kOff) *kOff = c->d.kOff;
    if (Klocal) *Klocal = c->d.K;
    if (kOwn0) *kOwn0 = c->d.kOwn0;
    if (kOwn1) *kOwn1 = c->d.kOwn1;
    return FLIP_OK;
}
int flip_set_halo(flip_ctx *c, int planes) {
    return guarded(c, [&] {
        if (c->nranks > 1) throw ApiError(FLIP_ERR_RUNTIME, "flip_set_halo must precede flip_set_slab");
        if (planes < slab_required_halo(c))
            throw ApiError(FLIP_ERR_DOMAIN, "the halo must cover the RK3 reach plus the extrapolation depth "
                                            "(ceil(CFL) + 1 + extrapolation layers + 1 planes)");
        c->halo = planes;
    });
}

}  // extern "C"
