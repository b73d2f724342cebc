// Internal declarations shared by the translation units of libflip_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <string>
#include <vector>
#include "../../include/flip_b200.h"

#define FLIP_CUDA_CHECK(call)                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            throw flip::CudaError(std::string(#call) + ": " + cudaGetErrorString(e__) + " at " + \
                                  __FILE__ + ":" + std::to_string(__LINE__));                   \
        }                                                                                       \
    } while (0)

namespace flip {

struct Comm;   // comm.cpp

struct CudaError {
    std::string msg;
    explicit CudaError(const std::string &m) : msg(m) {}
};
struct ApiError {
    int code;
    std::string msg;
    ApiError(int c, const std::string &m) : code(c), msg(m) {}
};

// Grid geometry handed to kernels by value.
// With a z-slab decomposition (flip_set_slab) every array is LOCAL: K planes = the owned planes
// [kOwn0,kOwn1) plus halo planes on the sides that have a neighbour; local plane kl is global plane
// kl + kOff.  Particle coordinates stay global.  Single GPU: kOff = 0, Kg = K, kOwn = [0,K).
struct Dims {
    int I, J, K;        // cells of the local grid
    int nU, nV, nW;     // face counts (local)
    int nC;             // cells (local)
    int nN;             // nodes (I+1)(J+1)(K+1) (local)
    double dx;
    int kOff;           // global k of local plane 0
    int Kg;             // global K
    int kOwn0, kOwn1;   // owned local planes
};

// Particle store: SoA, 6 float arrays (24 B/particle), double buffered for the per-step cell sort.
struct ParticleSoA {
    float *px = nullptr, *py = nullptr, *pz = nullptr;
    float *vx = nullptr, *vy = nullptr, *vz = nullptr;
};

// Device scalars written by kernels and read by later kernels / the host (one pinned mirror).
struct DeviceScalars {
    // particle bookkeeping
    int numParticles;          // survivors after the last sort/compaction
    int removedSolid, removedCrowded, removedFast;
    unsigned int maxSpeedSqBits; // max |v|^2 over survivors, float bits (non-negative => uint order)
    int speedHist[8];          // _getMarkerParticleSpeedLimit bins
    float maxSpeedLimit;       // result of the histogram walk
    int anyCrowded;            // some cell holds more than maxParticlesPerCell candidates
    // pressure
    int numRows;               // n
    int numFluidCells;         // phi<0 over [1,N-2]^3 (== numRows)
    int numSegments;
    unsigned long long rhsMaxBits;  // ||b||_inf as double bits
    // PCG (slots rotate by iteration)
    double dotSZ[3];
    double rho[3];
    unsigned long long rMaxBits[3];
    int pcgIterations;
    int pcgDone;               // 0 running, 1 converged, 3 breakdown
    double pcgError;
    double pcgTol;
    // extrapolation frontier sizes
    int frontierCount[2];
    // z-slab exchange: particle counts sent to / received from the lower [0] and upper [1] neighbour
    int sendCount[2], recvCount[2];
    int globalParticles;       // sum over ranks of the sort input (speed-limit rule)
    int globalRows;
    int extCount[96];          // extrapolation frontier sizes, [component][layer] (grid.cu: EXT_MAX_LAYERS + 1 per component)
    int deferredCount;         // particles the single-precision G2P / RK3 kernels left to the literal ones
    int slabError[3];          // z-slab exchange: [0] an emigrant jumped past the neighbouring slab, [1] a send buffer
                               // overflowed, [2] an RK3 sample left the halo planes; summed over the ranks before
                               // anybody acts on them
    int pocketChanged;         // enclosed-pocket search of the moving-solid path: a pass changed a flag
    // multigrid coarse solve etc.
    int pad[7];
};

}  // namespace flip

struct flip_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    flip::Dims d;
    std::string lastError;
    int64_t launches = 0;

    // ---- configuration (defaults: fluidsimulation.h:1518-1701)
    double gravity[3] = {0, 0, 0};
    double ratioPICFLIP = 0.05;
    double CFL = 5.0;
    int minSubsteps = 1, maxSubsteps = 6;
    double pressureTol = 1e-9, pressureAcceptableTol = 1.0;
    int pressureMaxIter = 1000;
    int preconditioner = 1;
    int pressureWarmStart = 1;   // PCG starts from the previous substep's pressure where the cell was a row then
    int mgNu = 2, mgCoarseSweeps = 8;
    int samplingMode = FLIP_SAMPLING_FAST;   // trilinear blend of G2P / RK3 in float (indices and weights stay exact)
    int pcgPersistent = 0;   // measured slower than the multi-launch solver at 2 M rows (occupancy-limited), kept selectable
    double mgOmega = 0.9, mgScale = 1.8;
    // damping of pre-sweep s (post-sweeps mirror it).  Default: the two sweeps form the degree-2 Chebyshev
    // polynomial of the interval [0.4, 2] of the Jacobi-scaled spectrum (1/nodes), measured 23 -> 21 PCG iterations
    // at 256^3 against twice 0.9 at the same cost
    double mgOmegaSched[8] = {0.5664, 1.5765, 0.9, 0.9, 0.9, 0.9, 0.9, 0.9};
    int maxParticlesPerCell = 250;
    double solidBufferWidth = 0.1f;          // float in the reference (fluidsimulation.h:1685)
    double maxExtremeVelocityRemovalPercent = 0.0005;
    int maxExtremeVelocityRemovalAbsolute = 35;
    bool extremeVelocityRemoval = true;      // _isExtremeVelocityRemovalEnabled (fluidsimulation.h; :4345)
    double markerParticleScale = 3.0;        // _markerParticleScale: the mesher's particle radius factor (:5083)
    float markerParticleStepDistanceFactor = 0.5f;
    int nearSolidFactor = 3;
    int solidExactBand = 3;
    int extrapolationLayers = 7;             // ceil(CFL)+2  fluidsimulation.cpp:3777
    double liquidRadius = 0.0;               // 0.5*dx*sqrt(3)  fluidsimulation.cpp:2649

    // ---- state machine of update()
    bool initialized = false;
    int currentFrame = 0;
    int substepNumber = 0;
    double frameDt = 0, frameRemaining = 0, substepDt = 0;
    bool firstStepEver = true;               // frame 0 / substep 0 uses the predicted speed (:5595)
    std::vector<flip_step_stats> stats;      // per substep of the last frame
    flip_step_stats cur;                     // being filled

    // ---- host staging
    std::vector<float> loadQueuePos, loadQueueVel;   // loadMarkerParticleData queue
    // addMeshFluid queue (_addedFluidMeshObjectQueue, fluidsimulation.h): cell range [lo,hi) of the object, its fluid
    // velocity, and its signed distance field -- a nodal array of the global grid, or (sdf empty) an axis-aligned box
    // kind 0: queued object, seeded once; 1: inflow source (emits at the end of every substep); 2: outflow source
    // (MeshFluidSource, meshfluidsource.h; _updateMeshFluidSources fluidsimulation.cpp:4700-4722)
    struct FluidObject {
        int lo[3], hi[3]; double boxLo[3], boxHi[3], vel[3]; std::vector<float> sdf;
        int kind = 0, id = 0; bool enabled = true, constrained = true; float *dsdf = nullptr;
        int sdfOrigin[3] = {0, 0, 0}, sdfNodes[3] = {0, 0, 0};   // the source's own level-set grid (meshfluidsource.cpp:178-198)
    };
    std::vector<FluidObject> fluidObjects;
    int nextSourceId = 1;
    int nextParticleId = 0;                   // ids of particles seeded after the load (enableParticleIds)
    std::vector<float> hostSolidPhi;          // nodal: the domain (built-in box or flip_set_solid_sdf), WITHOUT the obstacles
    // static obstacles (addMeshObstacle, fluidsimulation.cpp:1994): nodal SDFs merged into the solid SDF by minimum
    // (MeshLevelSet::calculateUnion, meshlevelset.cpp:1758-1795)
    struct Obstacle {
        int id = 0;
        bool enabled = true;
        std::vector<float> sdf;
        // a box that moves rigidly (MeshObject::updateMeshAnimated, meshobject.cpp:61-95): the box it was added as and the
        // translations of the previous, the current and the next frame's mesh against it
        bool isBox = false, animated = false;
        double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        double offPrev[3] = {0, 0, 0}, offCur[3] = {0, 0, 0}, offNext[3] = {0, 0, 0};
        float velocity[3] = {0, 0, 0};            // of the substep under way
        float friction = 0.0f;                    // MeshObject::setFriction (meshobject.cpp:300-304)
        // a closed triangle mesh that moves (rigidly or deforming, fixed topology): the three frames' vertices
        bool isMesh = false;
        std::vector<int> triangles;
        std::vector<float> vertsPrev, vertsCur, vertsNext;        // xyz triplets
    };
    // friction of the solids on the partly open faces (_getFaceFrictionU/V/W, fluidsimulation.cpp:3785-3853): null while
    // every friction is 0 (the default).  userFaceFriction: handed in with flip_set_face_friction, not derived here.
    float *fricU = nullptr, *fricV = nullptr, *fricW = nullptr;
    float boundaryFriction = 0.0f;            // setBoundaryFriction (:1747-1759)
    bool userFaceFriction = false;
    bool solidVelFromAnimation = false;       // solU/V/W are rebuilt every substep from the animated obstacles
    float *solidWeightSum[3] = {nullptr, nullptr, nullptr};    // device: summed solid fractions of the faces (U, V, W)
    unsigned char *solidValid[3] = {nullptr, nullptr, nullptr}; // device: faces whose solid velocity is defined
    std::vector<Obstacle> obstacles;
    int nextObstacleId = 1;
    bool solidDirty = false;                  // obstacles changed after initialize: re-derived at the next substep (:2007)
    bool userSolidPhi = false;
    std::vector<float> hU, hV, hW;            // host mirror handed out by getVelocityField

    // ---- device memory
    int capacity = 0;                         // particle capacity
    int np = 0;                               // live OWNED particles (host copy)
    int npStore = 0;                          // particles in the store (owned + ghosts of neighbouring slabs)
    flip::ParticleSoA P[2];                   // ping-pong
    int cur_buf = 0;
    int *cellOfParticle = nullptr;            // [capacity] destination cell (+ extreme-velocity flag in bit 30) or -1 (removed)
    int *sortIdx = nullptr;                   // [capacity] particle indices grouped by cell
    int *srcIdx = nullptr;                    // [capacity] final gather map
    int *pid[2] = {nullptr, nullptr};         // optional particle ids carried through the sorts
    bool trackIds = false;
    int particleIdBase = 0;                   // id of the first loaded particle (a z-slab rank that is handed its own part of a scene)
    unsigned char *occ = nullptr;             // 4x4x4-cell blocks that hold particles (rebuilt by every sort)
    size_t occBytes = 0;
    int *cellCount = nullptr;                 // [nC+1]
    int *cellStart = nullptr;                 // [nC+1] exclusive scan of kept counts (valid for current particles)
    int *cellStartA = nullptr;                // [nC+1] scan of candidate counts
    void *scanTemp = nullptr;
    size_t scanTempBytes = 0;

    float *U = nullptr, *V = nullptr, *W = nullptr;
    float *sU = nullptr, *sV = nullptr, *sW = nullptr;    // saved field
    unsigned char *validU = nullptr, *validV = nullptr, *validW = nullptr;
    unsigned char *status = nullptr;          // extrapolation level grid, max(nU,nV,nW)
    int *frontier[2] = {nullptr, nullptr};
    float *phiL = nullptr;                    // liquid SDF
    float *phiS = nullptr;                    // solid SDF, nodal
    float *wU = nullptr, *wV = nullptr, *wW = nullptr, *wC = nullptr;
    // face velocities of the solids (MeshLevelSet::getFaceVelocityU/V/W, meshlevelset.cpp:207-231): null while every
    // solid is at rest.  They enter the divergence (pressuresolver.cpp:608-613) and the solid constraint (:3884-3933).
    float *solU = nullptr, *solV = nullptr, *solW = nullptr;
    unsigned char *pocketFlag = nullptr;      // [nC] scratch of the enclosed-pocket search (pressuresolver.cpp:124-244)
    unsigned char *nearSolid = nullptr;
    int nsI = 0, nsJ = 0, nsK = 0;
    float *pressure = nullptr;                // (I,J,K) float, last solution
    void *p2gAcc[4] = {nullptr, nullptr, nullptr, nullptr};   // fixed-point accumulators of the P2G scatter: U, V, W faces; cell minima
    unsigned int *occBits = nullptr;          // cell occupancy bitmaps: occupied | 3x3x3 neighbourhood | 5x5x5 neighbourhood
    // surface reconstruction (mesher.cu): getIsomesh() settings of the reference (fluidsimulation.h:1605-1608)
    int surfaceSubdivision = 1;
    double surfaceSmoothingValue = 0.5;
    int surfaceSmoothingIterations = 2;
    void *mesher = nullptr;
    bool fuseAdvance = false;                 // set by flip_update around its stage loop
    bool fusedAdvanceDone = false;            // whole-step path: the advection kernels already ran with the G2P stage
    bool speedHistReady = false;              // ... and left the speed histogram of the removal rules in the device scalars
    int64_t stepCounter = 0;                  // bumped whenever the particle store is re-sorted (cache stamp of the mesh)
    int *p2gTiles = nullptr;                  // tile bookkeeping of the P2G scatter / SDF shell search (particles.cu)
    bool occBitsValid = false;                // the sort has just written the first of them (rows of whole words)

    // pressure system, dense-indexed vectors over cells + active 32-cell segments
    int *segCell = nullptr;                   // [maxSegments] first cell of the segment
    unsigned int *segMask = nullptr;          // [maxSegments] liquid lanes
    int maxSegments = 0;
    double *Adiag = nullptr;                  // [nC]
    float *AoffU = nullptr, *AoffV = nullptr, *AoffW = nullptr;  // [nC] masked upper-face weights
    double *vx_ = nullptr, *vr = nullptr, *vs = nullptr, *vz = nullptr, *vb = nullptr;   // [nC]
    void *mg = nullptr;                       // multigrid hierarchy (pressure.cu)

    flip::DeviceScalars *dS = nullptr;        // device
    flip::DeviceScalars *hS = nullptr;        // pinned host mirror

    cudaEvent_t evStage[FLIP_NUM_STAGES + 1];
    float stageMs[FLIP_NUM_STAGES] = {0};
    bool eventsCreated = false;

    // optional per-kernel-class device timing (CUDA events on the context's stream)
    bool ktEnabled = false;
    std::vector<cudaEvent_t> ktPool;
    size_t ktUsed = 0;
    struct KtPending { int cls; size_t a, b; };
    std::vector<KtPending> ktPending;
    double ktSumMs[FLIP_NUM_KERNEL_CLASSES] = {0};
    int64_t ktCount[FLIP_NUM_KERNEL_CLASSES] = {0};

    // multi-GPU slab
    int rank = 0, nranks = 1;
    int halo = 16;                            // halo planes towards each neighbour
    int KgCfg = 0;                            // global K given to flip_create
    flip::Comm *comm = nullptr;
    int ownedBegin = 0, ownedEnd = 0;         // owned particles inside the sorted (ghost-extended) store
    bool ghostsPresent = false;
    void *peer = nullptr;                     // peer-memory exchange state (peer.cu); NCCL is used when it is off
    float *sendBuf[2] = {nullptr, nullptr};   // emigrant staging, 7 arrays of sendCap entries each
    int sendCap = 0;
    int np_global = 0;
};

namespace flip {

// static_host.cpp
void build_box_solid_sdf(const Dims &d, std::vector<float> &phi);
void build_weights(const Dims &d, const std::vector<float> &phi, std::vector<float> &wU, std::vector<float> &wV,
                   std::vector<float> &wW, std::vector<float> &wC);
void build_center_weights(const Dims &d, const std::vector<float> &phi, std::vector<float> &wC);
struct FrictionSolid { const std::vector<float> *phi; float friction; };
void build_face_friction(const Dims &d, int band, const std::vector<FrictionSolid> &solids, std::vector<float> &fU,
                         std::vector<float> &fV, std::vector<float> &fW);
int mesh_velocity_data(const Dims &d, const float *vertices_xyz, int num_vertices, const int *triangles, int num_triangles,
                       const float *vertex_velocities_xyz, int band, float far_value, std::vector<float> &phi,
                       std::vector<float> fraction[3], std::vector<float> field[3]);
void add_solid_fractions(const Dims &d, const std::vector<float> &phi, const float velocity[3], std::vector<float> weightSum[3],
                         std::vector<float> fieldSum[3]);
void build_near_solid(const Dims &d, const std::vector<float> &phi, int factor, int band, double cfl,
                      std::vector<unsigned char> &grid, int &gi, int &gj, int &gk);

// particles.cu
void particles_alloc(flip_ctx *c, int capacity);
void particles_free(flip_ctx *c);
void particles_upload_aos(flip_ctx *c, const float *aos6, int n);      // filter + sort
void particles_upload_split(flip_ctx *c, const float *pos, const float *vel, int n);
void particles_download_aos(flip_ctx *c, float *aos6);
void particles_download_component(flip_ctx *c, float *xyz, int which);  // 0 pos, 1 vel
void particles_download_ids(flip_ctx *c, int *ids);
void particles_sort(flip_ctx *c, bool applyRemovalRules, double frameDt, int srcOffset, int count, bool ownedOnly);
void stage_liquid_sdf(flip_ctx *c);
void stage_p2g(flip_ctx *c);
void stage_g2p(flip_ctx *c);
void stage_advance(flip_ctx *c, double dt);
bool stage_g2p_advance_fused(flip_ctx *c, double dt);   // whole-step path: G2P + RK3 in one pass (false: not applicable)

// seed.cu
void stage_fluid_objects(flip_ctx *c);          // seeds the queued fluid objects / runs the sources (end of a substep)
bool has_constrained_inflow(const flip_ctx *c);          // an enabled inflow source that pins the velocity inside it
void stage_inflow_body_force_exclusion(flip_ctx *c);     // _getInflowConstrainedVelocityComponents :3372
void stage_inflow_constrain_particles(flip_ctx *c);      // _constrainMarkerParticleVelocities :4161

// mesher.cu
void mesher_get(flip_ctx *c, int *nv, int *nt, float *verts, int *tris);
void mesher_free(flip_ctx *c);
void mesher_debug_field(flip_ctx *c, float *values, unsigned char *inside, unsigned char *need);

// grid.cu
void stage_extrapolate(flip_ctx *c);
void stage_save(flip_ctx *c);
void stage_body_force(flip_ctx *c, double dt);
void stage_constrain(flip_ctx *c);
void solid_velocity_normalize_extrapolate(flip_ctx *c, int layers);

// pressure.cu
void pressure_alloc(flip_ctx *c);
void pressure_free(flip_ctx *c);
void stage_pressure(flip_ctx *c, double dt);

// comm.cpp: NCCL (loaded with dlopen) for the z-slab exchanges
Comm *comm_create(int rank, int nranks, const void *uniqueId, int idBytes);
void comm_destroy(Comm *);
int comm_unique_id(void *out, int idBytes);
void comm_group_begin(Comm *);
void comm_group_end(Comm *);
void comm_send(Comm *, const void *buf, size_t bytes, int peer, cudaStream_t st);
void comm_recv(Comm *, void *buf, size_t bytes, int peer, cudaStream_t st);
enum { COMM_SUM_F64 = 0, COMM_MAX_U64 = 1, COMM_SUM_I32 = 2, COMM_MAX_U32 = 3 };
void comm_allreduce(Comm *, void *buf, size_t count, int kind, cudaStream_t st);
void comm_allgather_f32(Comm *, const float *send, float *recv, size_t countPerRank, cudaStream_t st);

// peer.cu: halo planes and PCG scalars through CUDA-IPC peer memory (NVLink), one kernel per exchange
void peer_setup(flip_ctx *c);                           // collective, after the communicator exists
void peer_free(flip_ctx *c);
bool peer_on(const flip_ctx *c);
void peer_exchange_planes(flip_ctx *c, const void *sendLo, void *recvLo, const void *sendHi, void *recvHi, size_t bytes);
void peer_allreduce(flip_ctx *c, void *val, int count, int kind);   // COMM_SUM_F64 / COMM_MAX_U64, count <= 8

// slab.cu: halo exchanges between neighbouring slabs
void slab_allreduce_scalar(flip_ctx *c, void *val, int kind);      // one fp64 sum / u64 max on the solver's stream
void slab_exchange_ghosts(flip_ctx *c);                 // ghost particles within `halo` planes, then re-sort
void slab_drop_ghosts_and_migrate(flip_ctx *c);         // after advance: emigrants out, immigrants in
void slab_exchange_planes(flip_ctx *c, float *field, int planeElems, int facePlanes);   // halo planes of a float grid
void slab_exchange_planes_u8(flip_ctx *c, unsigned char *field, int planeElems, int facePlanes);
void slab_exchange_vector_halo(flip_ctx *c, double *v); // one cell plane each way (PCG search vector / pressure)
void slab_exchange_cell_plane_f32(flip_ctx *c, float *v, int planeElems, int k0, int k1);
inline bool slab_on(const flip_ctx *c);

// helpers
void scalars_to_host(flip_ctx *c);   // async copy + sync
size_t kt_begin(flip_ctx *c);                    // records a start event, returns its slot (or 0 when disabled)
void kt_end(flip_ctx *c, int cls, size_t slot);  // records the stop event
void kt_collect(flip_ctx *c);                    // after a stream sync: folds pending pairs into the sums
inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }
// per-component stride of the extrapolation scratch arrays (level bytes, frontier lists): the largest face
// count, padded and rounded so that every component's slice stays 256-byte aligned
inline size_t ext_stride(const Dims &d) {
    size_t m = (size_t)(d.nU > d.nV ? (d.nU > d.nW ? d.nU : d.nW) : (d.nV > d.nW ? d.nV : d.nW));
    return ((m + 63) / 64 + 1) * 64;
}
inline bool slab_on(const flip_ctx *c) { return c->nranks > 1; }
// Halo planes a z-slab needs towards a neighbour: an owned particle travels up to ceil(CFL) cells in a substep and its
// RK3 / G2P stencils reach one plane further; the redundantly extrapolated field of a slab is contaminated from its
// local boundary plane inwards by one plane per extrapolation layer (the boundary counts as a border face there)
inline int slab_required_halo(const flip_ctx *c) { return (int)ceil(c->CFL) + 1 + c->extrapolationLayers + 1; }

}  // namespace flip
