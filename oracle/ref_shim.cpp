/*
 * ref_shim.cpp — C-ABI driver around the UNMODIFIED reference engine (jklae/FLIPEngine3D,
 * src/engine, Blender-FLIP-Fluids 1.0.9).  TEST INFRASTRUCTURE ONLY.
 *
 * This file is ours; it is compiled together with the reference sources (where they lie under
 * /root/reference) into oracle/_ref/libflipref_{golden,fast}.so by oracle/Makefile.  It lets
 * tests/ and bench.py (a) run the reference end to end through its public API, and (b) run the
 * reference's own step stage by stage (the private member functions _stepFluid calls, in the same
 * order: fluidsimulation.cpp:5471-5508) so that every intermediate array can be dumped and
 * compared with the CUDA path.  Private members are reached with the `#define private public`
 * trick (no source edit; access specifiers do not change layout under the Itanium ABI).
 *
 * Nothing under flipengine3d_b200/ may link, load or call this library.
 */
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#define private public
#define protected public
#include "engine/fluidsimulation.h"
#include "engine/pressuresolver.h"
#include "engine/velocityadvector.h"
#include "engine/particlelevelset.h"
#include "engine/meshlevelset.h"
#include "engine/macvelocityfield.h"
#include "engine/particlemesher.h"
#include "engine/meshfluidsource.h"
#include "engine/trianglemesh.h"
#include "engine/threadutils.h"
#include "engine/stopwatch.h"
#undef private
#undef protected

namespace {

struct RefSim {
    FluidSimulation *sim = nullptr;
    std::string lastError;
    // timing of the stage-wise step, seconds, indexed by stage id
    double stageTime[16] = {0};
    TriangleMesh isomesh;      // ref_isomesh
    std::vector<MeshFluidSource *> sources;   // kept alive for the simulation (it stores the pointers)
    std::vector<MeshObject *> obstacles;      // likewise (addMeshObstacle keeps the pointer)
};

template <class F>
int guarded(RefSim *h, F &&f) {
    try {
        f();
        return 0;
    } catch (std::exception &e) {
        h->lastError = e.what();
        return 1;
    }
}

int parseIterations(const std::string &s) {
    // "Pressure Solver Iterations: N\nEstimated Error: E"  (pressuresolver.cpp:816-835)
    size_t p = s.find("Iterations: ");
    if (p == std::string::npos) return -1;
    return atoi(s.c_str() + p + 12);
}

double parseError(const std::string &s) {
    size_t p = s.find("Estimated Error: ");
    if (p == std::string::npos) return -1.0;
    return atof(s.c_str() + p + 17);
}

}  // namespace

extern "C" {

void *ref_create(int isize, int jsize, int ksize, double dx) {
    RefSim *h = new RefSim();
    h->sim = new FluidSimulation(isize, jsize, ksize, dx);
    h->sim->disableConsoleOutput();
    return h;
}

void ref_destroy(void *p) {
    RefSim *h = (RefSim *)p;
    delete h->sim;
    delete h;
}

const char *ref_last_error(void *p) { return ((RefSim *)p)->lastError.c_str(); }

void ref_set_threads(int n) { ThreadUtils::setMaxThreadCount(n); }
int ref_get_threads() { return ThreadUtils::getMaxThreadCount(); }

void ref_add_body_force(void *p, double fx, double fy, double fz) {
    ((RefSim *)p)->sim->addBodyForce(fx, fy, fz);
}

void ref_set_pressure_tolerance(void *p, double tol) {
    ((RefSim *)p)->sim->_pressureSolveTolerance = tol;
}

/* positions / velocities: n float triplets each.  Must be called before ref_initialize
 * (the load queue is only drained there: fluidsimulation.cpp:2669-2675,2791). */
int ref_load_particles(void *p, int n, const float *pos, const float *vel) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        FluidSimulationMarkerParticleData d;
        d.size = n;
        d.positions = (char *)pos;
        d.velocities = (char *)vel;
        h->sim->loadMarkerParticleData(d);
    });
}

int ref_initialize(void *p) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        h->sim->disableSurfaceReconstruction();
        h->sim->disableConsoleOutput();
        h->sim->initialize();
    });
}

/* Whole frame through the reference's public API (fluidsimulation.cpp:5755). */
int ref_update(void *p, double dt) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] { h->sim->update(dt); });
}

int ref_num_particles(void *p) { return (int)((RefSim *)p)->sim->_markerParticles.size(); }
int ref_current_frame(void *p) { return ((RefSim *)p)->sim->getCurrentFrame(); }
int ref_last_substeps(void *p) { return ((RefSim *)p)->sim->_currentFrameTimeStepNumber; }

/* The reference's LOGICAL particle view (operator[] of FragmentedVector),
 * as AoS {px,py,pz,vx,vy,vz}. */
void ref_get_particles(void *p, float *aos) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    unsigned int n = s->_markerParticles.size();
    for (unsigned int i = 0; i < n; i++) {
        MarkerParticle mp = s->_markerParticles[i];
        aos[6 * i + 0] = mp.position.x;
        aos[6 * i + 1] = mp.position.y;
        aos[6 * i + 2] = mp.position.z;
        aos[6 * i + 3] = mp.velocity.x;
        aos[6 * i + 4] = mp.velocity.y;
        aos[6 * i + 5] = mp.velocity.z;
    }
}

/* Overwrite the particle store (lock-step re-synchronisation). Same count only. */
int ref_set_particles(void *p, int n, const float *aos) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    if ((int)s->_markerParticles.size() != n) return 1;
    for (int i = 0; i < n; i++) {
        MarkerParticle mp(vmath::vec3(aos[6 * i], aos[6 * i + 1], aos[6 * i + 2]),
                          vmath::vec3(aos[6 * i + 3], aos[6 * i + 4], aos[6 * i + 5]));
        s->_markerParticles[i] = mp;
    }
    return 0;
}

int ref_pcg_iterations(void *p) { return parseIterations(((RefSim *)p)->sim->_pressureSolverStatus); }
double ref_pcg_error(void *p) { return parseError(((RefSim *)p)->sim->_pressureSolverStatus); }
int ref_num_fluid_cells(void *p) { return ((RefSim *)p)->sim->_getNumFluidCells(); }
double ref_liquid_sdf_radius(void *p) { return ((RefSim *)p)->sim->_liquidSDFParticleRadius; }

/* ---- frame / substep bookkeeping, restating the loop of update() (fluidsimulation.cpp:5768-5824)
 *      so that stages can be driven one at a time.  ---- */
void ref_begin_frame(void *p, double dt) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    s->_timingData = FluidSimulation::TimingData();
    double epsdt = 1e-6;
    s->_isZeroLengthDeltaTime = dt < epsdt;
    dt = std::max(dt, epsdt);
    s->_isCurrentFrameFinished = false;
    s->_currentFrameDeltaTime = dt;
    s->_currentFrameDeltaTimeRemaining = dt;
    s->_currentFrameTimeStepNumber = 0;
    s->_isSkippedFrame = s->_isZeroLengthDeltaTime && s->_outputData.isInitialized;
}

/* Returns the substep length the reference would take next and commits it. */
double ref_begin_substep(void *p) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    double dt = s->_currentFrameDeltaTime;
    double substepTime = s->_currentFrameDeltaTime / (double)s->_minFrameTimeSteps;
    double eps = 1e-9;
    s->_currentFrameTimeStep = fmin(s->_calculateNextTimeStep(dt), s->_currentFrameDeltaTimeRemaining);
    double timeCompleted = s->_currentFrameDeltaTime - s->_currentFrameDeltaTimeRemaining;
    double stepLimit = (s->_currentFrameTimeStepNumber + 1) * substepTime;
    if (timeCompleted + s->_currentFrameTimeStep > stepLimit) {
        s->_currentFrameTimeStep = fmin(substepTime, s->_currentFrameDeltaTimeRemaining);
    }
    if (s->_currentFrameTimeStepNumber == s->_maxFrameTimeSteps - 1) {
        s->_currentFrameTimeStep = s->_currentFrameDeltaTimeRemaining;
    }
    s->_currentFrameDeltaTimeRemaining -= s->_currentFrameTimeStep;
    s->_isLastFrameTimeStep = fabs(s->_currentFrameDeltaTimeRemaining) < eps;
    return s->_currentFrameTimeStep;
}

/* returns 1 when the frame has more substeps to run */
int ref_end_substep(void *p) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    s->_currentFrameTimeStepNumber++;
    return s->_currentFrameDeltaTimeRemaining > 1e-9 ? 1 : 0;
}

void ref_end_frame(void *p) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    s->_outputData.isInitialized = true;
    s->_currentFrame++;
    s->_isCurrentFrameFinished = true;
}

/* Stage ids follow _stepFluid (fluidsimulation.cpp:5471-5508). */
enum {
    ST_OBSTACLES = 0,     /* S1  _updateObstacleObjects                         */
    ST_LIQUID_SDF = 1,    /* S2  _updateLiquidLevelSet + postProcess            */
    ST_P2G = 2,           /* S3a reset valid, clear MAC, VelocityAdvector::advect */
    ST_EXTRAPOLATE_A = 3, /* S3b _extrapolateFluidVelocities                    */
    ST_SAVE = 4,          /* S4  _saveVelocityField                             */
    ST_BODY_FORCE = 5,    /* S5  _applyBodyForcesToVelocityField                */
    ST_PRESSURE = 6,      /* S7a _updateWeightGrid + PressureSolver::solve      */
    ST_EXTRAPOLATE_B = 7, /* S7b _extrapolateFluidVelocities                    */
    ST_CONSTRAIN = 8,     /* S8  _constrainVelocityFields                       */
    ST_G2P = 9,           /* S10 _updateMarkerParticleVelocities                */
    ST_ADVANCE = 10,      /* S11+S12 delete saved field, _advanceMarkerParticles */
    ST_TAIL = 11          /* S13 _updateFluidObjects (S14 output skipped: meshing is off) */
};

int ref_stage(void *p, int stage, double dt) {
    RefSim *h = (RefSim *)p;
    FluidSimulation *s = h->sim;
    StopWatch t;
    t.start();
    int rc = guarded(h, [&] {
        switch (stage) {
            case ST_OBSTACLES:
                s->_updateObstacleObjects(dt);
                break;
            case ST_LIQUID_SDF:
                s->_updateLiquidLevelSet();
                s->_liquidSDF.postProcessSignedDistanceField(s->_solidSDF);
                break;
            case ST_P2G: {
                /* body of _advectVelocityField up to the extrapolation (fluidsimulation.cpp:3259-3268) */
                s->_validVelocities.reset();
                s->_MACVelocity.clear();
                if (!s->_markerParticles.empty()) {
                    VelocityAdvectorParameters params;
                    params.particles = &s->_markerParticles;
                    params.vfield = &s->_MACVelocity;
                    params.validVelocities = &s->_validVelocities;
                    params.particleRadius = s->_liquidSDFParticleRadius;
                    s->_velocityAdvector.advect(params);
                }
                break;
            }
            case ST_EXTRAPOLATE_A:
                if (!s->_markerParticles.empty()) {
                    s->_extrapolateFluidVelocities(s->_MACVelocity, s->_validVelocities);
                }
                break;
            case ST_SAVE:
                s->_saveVelocityField();
                break;
            case ST_BODY_FORCE:
                s->_applyBodyForcesToVelocityField(dt);
                break;
            case ST_PRESSURE: {
                /* body of _pressureSolve up to the extrapolation (fluidsimulation.cpp:3738-3761) */
                s->_updateWeightGrid();
                PressureSolverParameters params;
                params.cellwidth = s->_dx;
                params.deltaTime = dt;
                params.tolerance = s->_pressureSolveTolerance;
                params.acceptableTolerance = s->_pressureSolveAcceptableTolerance;
                params.maxIterations = s->_maxPressureSolveIterations;
                params.velocityField = &s->_MACVelocity;
                params.validVelocities = &s->_validVelocities;
                params.liquidSDF = &s->_liquidSDF;
                params.solidSDF = &s->_solidSDF;
                params.weightGrid = &s->_weightGrid;
                params.isSurfaceTensionEnabled = false;
                PressureSolver psolver;
                psolver.solve(params);
                s->_pressureSolverStatus = psolver.getSolverStatus();
                break;
            }
            case ST_EXTRAPOLATE_B:
                s->_extrapolateFluidVelocities(s->_MACVelocity, s->_validVelocities);
                break;
            case ST_CONSTRAIN:
                s->_constrainVelocityFields();
                break;
            case ST_G2P:
                s->_updateMarkerParticleVelocities();
                break;
            case ST_ADVANCE:
                s->_deleteSavedVelocityField();
                s->_advanceMarkerParticles(dt);
                break;
            case ST_TAIL:
                s->_updateFluidObjects();
                break;
            default:
                throw std::runtime_error("ref_stage: bad stage id");
        }
    });
    t.stop();
    if (stage >= 0 && stage < 16) h->stageTime[stage] = t.getTime();
    return rc;
}

double ref_stage_time(void *p, int stage) { return ((RefSim *)p)->stageTime[stage]; }

/* Array ids for ref_get_array / ref_set_array. */
enum {
    AR_U = 0, AR_V = 1, AR_W = 2,                 /* float, MAC layout (macvelocityfield.cpp:46-54) */
    AR_VALID_U = 3, AR_VALID_V = 4, AR_VALID_W = 5, /* uint8 */
    AR_LIQUID_PHI = 6,                            /* float (I,J,K) */
    AR_SOLID_PHI = 7,                             /* float (I+1,J+1,K+1) nodal */
    AR_WEIGHT_U = 8, AR_WEIGHT_V = 9, AR_WEIGHT_W = 10, AR_WEIGHT_C = 11, /* float */
    AR_SAVED_U = 12, AR_SAVED_V = 13, AR_SAVED_W = 14,
    AR_NEAR_SOLID = 15,                           /* uint8, coarse grid */
    AR_SOLID_VEL_U = 16, AR_SOLID_VEL_V = 17, AR_SOLID_VEL_W = 18  /* float, MAC layout: the face velocities kept with the
                                                     solid SDF (VelocityDataGrid::field, meshlevelset.h:69-87) */
};

static void *arrayPtr(FluidSimulation *s, int which, size_t *bytes) {
    auto f = [&](Array3d<float> *a) { *bytes = sizeof(float) * (size_t)a->getNumElements(); return (void *)a->getRawArray(); };
    auto b = [&](Array3d<bool> *a) { *bytes = (size_t)a->getNumElements(); return (void *)a->getRawArray(); };
    switch (which) {
        case AR_U: return f(s->_MACVelocity.getArray3dU());
        case AR_V: return f(s->_MACVelocity.getArray3dV());
        case AR_W: return f(s->_MACVelocity.getArray3dW());
        case AR_VALID_U: return b(&s->_validVelocities.validU);
        case AR_VALID_V: return b(&s->_validVelocities.validV);
        case AR_VALID_W: return b(&s->_validVelocities.validW);
        case AR_LIQUID_PHI: return f(&s->_liquidSDF._phi);
        case AR_SOLID_PHI: return f(&s->_solidSDF._phi);
        case AR_WEIGHT_U: return f(&s->_weightGrid.U);
        case AR_WEIGHT_V: return f(&s->_weightGrid.V);
        case AR_WEIGHT_W: return f(&s->_weightGrid.W);
        case AR_WEIGHT_C: return f(&s->_weightGrid.center);
        case AR_SAVED_U: return f(s->_savedVelocityField.getArray3dU());
        case AR_SAVED_V: return f(s->_savedVelocityField.getArray3dV());
        case AR_SAVED_W: return f(s->_savedVelocityField.getArray3dW());
        case AR_NEAR_SOLID: return b(&s->_nearSolidGrid);
        case AR_SOLID_VEL_U: return f(s->_solidSDF._velocityData.field.getArray3dU());
        case AR_SOLID_VEL_V: return f(s->_solidSDF._velocityData.field.getArray3dV());
        case AR_SOLID_VEL_W: return f(s->_solidSDF._velocityData.field.getArray3dW());
    }
    *bytes = 0;
    return nullptr;
}

long ref_array_bytes(void *p, int which) {
    size_t bytes = 0;
    arrayPtr(((RefSim *)p)->sim, which, &bytes);
    return (long)bytes;
}

int ref_get_array(void *p, int which, void *out) {
    size_t bytes = 0;
    void *src = arrayPtr(((RefSim *)p)->sim, which, &bytes);
    if (!src) return 1;
    memcpy(out, src, bytes);
    return 0;
}

int ref_set_array(void *p, int which, const void *in) {
    size_t bytes = 0;
    void *dst = arrayPtr(((RefSim *)p)->sim, which, &bytes);
    if (!dst) return 1;
    memcpy(dst, in, bytes);
    return 0;
}

void ref_near_solid_dims(void *p, int *gi, int *gj, int *gk) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    *gi = s->_nearSolidGrid.width;
    *gj = s->_nearSolidGrid.height;
    *gk = s->_nearSolidGrid.depth;
}

/* Forces the weight grid to exist (it is otherwise built lazily inside the first pressure solve). */
void ref_update_weight_grid(void *p) { ((RefSim *)p)->sim->_updateWeightGrid(); }
/* ... again, after the solid SDF was overwritten through ref_set_array (the flag _updateSolidLevelSet would clear, :3070) */
void ref_invalidate_weight_grid(void *p) { ((RefSim *)p)->sim->_isWeightGridUpToDate = false; }

/* Single-point samplers used by the unit tests of the restated interpolation code. */
void ref_sample_velocity(void *p, int n, const float *pos, float *out) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    for (int i = 0; i < n; i++) {
        vmath::vec3 v = s->_MACVelocity.evaluateVelocityAtPositionLinear(
            vmath::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
        out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
    }
}

void ref_sample_solid_phi(void *p, int n, const float *pos, float *out) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    for (int i = 0; i < n; i++) {
        out[i] = s->_solidSDF.trilinearInterpolate(vmath::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
    }
}

/* FluidManager::initialize's fluid object (src/FluidManager.cpp:53-64): an axis-aligned box as a 12-triangle mesh in a
 * MeshObject, queued with addMeshFluid (seeded by the reference at the end of the next step). */
int ref_add_mesh_fluid_box(void *p, const double lo[3], const double hi[3], const double vel[3]) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        FluidSimulation *s = h->sim;
        vmath::vec3 q((float)lo[0], (float)lo[1], (float)lo[2]);
        const double w = hi[0] - lo[0], ht = hi[1] - lo[1], d = hi[2] - lo[2];
        TriangleMesh m;
        m.vertices = {vmath::vec3(q.x, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z + d),
                      vmath::vec3(q.x, q.y, q.z + d), vmath::vec3(q.x, q.y + ht, q.z), vmath::vec3(q.x + w, q.y + ht, q.z),
                      vmath::vec3(q.x + w, q.y + ht, q.z + d), vmath::vec3(q.x, q.y + ht, q.z + d)};
        m.triangles = {Triangle(0, 1, 2), Triangle(0, 2, 3), Triangle(4, 7, 6), Triangle(4, 6, 5), Triangle(0, 3, 7), Triangle(0, 7, 4),
                       Triangle(1, 5, 6), Triangle(1, 6, 2), Triangle(0, 4, 5), Triangle(0, 5, 1), Triangle(3, 2, 6), Triangle(3, 6, 7)};
        MeshObject obj(s->_isize, s->_jsize, s->_ksize, s->_dx);
        obj.updateMeshStatic(m);
        s->addMeshFluid(obj, vmath::vec3((float)vel[0], (float)vel[1], (float)vel[2]));
    });
}

/* MeshLevelSet::fastCalculateSignedDistanceField(mesh, band) (meshlevelset.cpp:773-828) of an arbitrary triangle mesh on a
 * grid of its own: the nodal phi, (I+1)(J+1)(K+1) floats. */
int ref_mesh_level_set(int isize, int jsize, int ksize, double dx, const float *verts, int nv, const int *tris, int nt, int band,
                       float *out) {
    try {
        TriangleMesh m;
        for (int v = 0; v < nv; v++) m.vertices.push_back(vmath::vec3(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]));
        for (int t = 0; t < nt; t++) m.triangles.push_back(Triangle(tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]));
        MeshLevelSet sdf(isize, jsize, ksize, dx);
        sdf.disableVelocityData();
        sdf.fastCalculateSignedDistanceField(m, band);
        size_t q = 0;
        for (int k = 0; k <= ksize; k++)
            for (int j = 0; j <= jsize; j++)
                for (int i = 0; i <= isize; i++) out[q++] = sdf(i, j, k);
        return 0;
    } catch (std::exception &) {
        return 1;
    }
}

/* addMeshFluid with an arbitrary static closed mesh (queued, seeded at the end of the next step) */
int ref_add_mesh_fluid_mesh(void *p, const float *verts, int nv, const int *tris, int nt, const double vel[3]) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        FluidSimulation *s = h->sim;
        TriangleMesh m;
        for (int v = 0; v < nv; v++) m.vertices.push_back(vmath::vec3(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]));
        for (int t = 0; t < nt; t++) m.triangles.push_back(Triangle(tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]));
        MeshObject obj(s->_isize, s->_jsize, s->_ksize, s->_dx);
        obj.updateMeshStatic(m);
        s->addMeshFluid(obj, vmath::vec3((float)vel[0], (float)vel[1], (float)vel[2]));
    });
}

/* setCFLConditionNumber (:1765), setPICFLIPRatio, setMin / MaxTimeStepsPerFrame (:1801-1831) */
int ref_set_step_settings(void *p, int cfl, double picflip, int min_steps, int max_steps) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        FluidSimulation *s = h->sim;
        if (cfl > 0) s->setCFLConditionNumber(cfl);
        if (picflip >= 0.0) s->setPICFLIPRatio(picflip);
        if (min_steps > 0) s->setMinTimeStepsPerFrame(min_steps);
        if (max_steps > 0) s->setMaxTimeStepsPerFrame(max_steps);
    });
}

/* enable / disableExtremeVelocityRemoval (:1869-1881), setMarkerParticleScale (:168-179) */
void ref_set_extreme_velocity_removal(void *p, int on) {
    FluidSimulation *s = ((RefSim *)p)->sim;
    if (on) s->enableExtremeVelocityRemoval(); else s->disableExtremeVelocityRemoval();
}
int ref_set_marker_particle_scale(void *p, double scale) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] { h->sim->setMarkerParticleScale(scale); });
}

/* FluidSimulation::addMeshObstacle (fluidsimulation.cpp:1994-2008) with a static box MeshObject; returns its index in this
 * shim's list.  ref_remove_obstacle: removeMeshObstacle (:2010-2031). */
int ref_add_obstacle_box(void *p, const double lo[3], const double hi[3]) {
    RefSim *h = (RefSim *)p;
    int idx = -1;
    guarded(h, [&] {
        FluidSimulation *s = h->sim;
        vmath::vec3 q((float)lo[0], (float)lo[1], (float)lo[2]);
        const double w = hi[0] - lo[0], ht = hi[1] - lo[1], d = hi[2] - lo[2];
        TriangleMesh m;
        m.vertices = {vmath::vec3(q.x, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z + d),
                      vmath::vec3(q.x, q.y, q.z + d), vmath::vec3(q.x, q.y + ht, q.z), vmath::vec3(q.x + w, q.y + ht, q.z),
                      vmath::vec3(q.x + w, q.y + ht, q.z + d), vmath::vec3(q.x, q.y + ht, q.z + d)};
        m.triangles = {Triangle(0, 1, 2), Triangle(0, 2, 3), Triangle(4, 7, 6), Triangle(4, 6, 5), Triangle(0, 3, 7), Triangle(0, 7, 4),
                       Triangle(1, 5, 6), Triangle(1, 6, 2), Triangle(0, 4, 5), Triangle(0, 5, 1), Triangle(3, 2, 6), Triangle(3, 6, 7)};
        MeshObject *obj = new MeshObject(s->_isize, s->_jsize, s->_ksize, s->_dx);
        obj->updateMeshStatic(m);
        s->addMeshObstacle(obj);
        h->obstacles.push_back(obj);
        idx = (int)h->obstacles.size() - 1;
    });
    return idx;
}
/* MeshObject::updateMeshAnimated (meshobject.cpp:61-95) for obstacle `idx`: the box [lo,hi] moved by the three offsets as
 * the previous, current and next frame's mesh (call once per frame, before update / the stage-wise step). */
static TriangleMesh boxMesh(const double lo[3], const double hi[3], const double off[3]) {
    vmath::vec3 q((float)(lo[0] + off[0]), (float)(lo[1] + off[1]), (float)(lo[2] + off[2]));
    const double w = hi[0] - lo[0], ht = hi[1] - lo[1], d = hi[2] - lo[2];
    TriangleMesh m;
    m.vertices = {vmath::vec3(q.x, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z + d),
                  vmath::vec3(q.x, q.y, q.z + d), vmath::vec3(q.x, q.y + ht, q.z), vmath::vec3(q.x + w, q.y + ht, q.z),
                  vmath::vec3(q.x + w, q.y + ht, q.z + d), vmath::vec3(q.x, q.y + ht, q.z + d)};
    m.triangles = {Triangle(0, 1, 2), Triangle(0, 2, 3), Triangle(4, 7, 6), Triangle(4, 6, 5), Triangle(0, 3, 7), Triangle(0, 7, 4),
                   Triangle(1, 5, 6), Triangle(1, 6, 2), Triangle(0, 4, 5), Triangle(0, 5, 1), Triangle(3, 2, 6), Triangle(3, 6, 7)};
    return m;
}
int ref_animate_obstacle_box(void *p, int idx, const double lo[3], const double hi[3], const double offPrev[3],
                             const double offCur[3], const double offNext[3]) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        if (idx < 0 || idx >= (int)h->obstacles.size() || !h->obstacles[idx]) throw std::runtime_error("no such obstacle");
        h->obstacles[idx]->updateMeshAnimated(boxMesh(lo, hi, offPrev), boxMesh(lo, hi, offCur), boxMesh(lo, hi, offNext));
    });
}
/* addMeshObstacle with an arbitrary closed mesh (static until ref_animate_obstacle_mesh), and updateMeshAnimated with the
 * vertices of the previous / current / next frame (same triangles). */
static TriangleMesh meshOf(const float *verts, int nv, const int *tris, int nt) {
    TriangleMesh m;
    for (int v = 0; v < nv; v++) m.vertices.push_back(vmath::vec3(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]));
    for (int t = 0; t < nt; t++) m.triangles.push_back(Triangle(tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]));
    return m;
}
int ref_add_obstacle_mesh(void *p, const float *verts, int nv, const int *tris, int nt) {
    RefSim *h = (RefSim *)p;
    int idx = -1;
    guarded(h, [&] {
        FluidSimulation *s = h->sim;
        MeshObject *obj = new MeshObject(s->_isize, s->_jsize, s->_ksize, s->_dx);
        obj->updateMeshStatic(meshOf(verts, nv, tris, nt));
        s->addMeshObstacle(obj);
        h->obstacles.push_back(obj);
        idx = (int)h->obstacles.size() - 1;
    });
    return idx;
}
int ref_animate_obstacle_mesh(void *p, int idx, const float *prev, const float *cur, const float *next, int nv, const int *tris, int nt) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        if (idx < 0 || idx >= (int)h->obstacles.size() || !h->obstacles[idx]) throw std::runtime_error("no such obstacle");
        h->obstacles[idx]->updateMeshAnimated(meshOf(prev, nv, tris, nt), meshOf(cur, nv, tris, nt), meshOf(next, nv, tris, nt));
    });
}
/* setBoundaryFriction (:1747-1759), MeshObject::setFriction of obstacle `idx`, and the face friction the constraint reads
 * (_getFaceFrictionU/V/W :3785-3853) for every face of component comp (0 U, 1 V, 2 W), in the MAC layout. */
int ref_set_boundary_friction(void *p, double f) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] { h->sim->setBoundaryFriction(f); });
}
int ref_set_obstacle_friction(void *p, int idx, double f) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        if (idx < 0 || idx >= (int)h->obstacles.size() || !h->obstacles[idx]) throw std::runtime_error("no such obstacle");
        h->obstacles[idx]->setFriction((float)f);
    });
}
int ref_face_friction(void *p, int comp, float *out) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        FluidSimulation *s = h->sim;
        const int I = s->_isize, J = s->_jsize, K = s->_ksize;
        const int ni = comp == 0 ? I + 1 : I, nj = comp == 1 ? J + 1 : J, nk = comp == 2 ? K + 1 : K;
        size_t q = 0;
        for (int k = 0; k < nk; k++)
            for (int j = 0; j < nj; j++)
                for (int i = 0; i < ni; i++, q++) {
                    GridIndex g(i, j, k);
                    out[q] = comp == 0 ? s->_getFaceFrictionU(g) : (comp == 1 ? s->_getFaceFrictionV(g) : s->_getFaceFrictionW(g));
                }
    });
}
int ref_remove_obstacle(void *p, int idx) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        if (idx < 0 || idx >= (int)h->obstacles.size() || !h->obstacles[idx]) throw std::runtime_error("no such obstacle");
        h->sim->removeMeshObstacle(h->obstacles[idx]);
        h->obstacles[idx] = nullptr;
    });
}

/* FluidSimulation::addMeshFluidSource with a static box MeshFluidSource: inflow (setInflow + setVelocity) or outflow
 * (setOutflow; fluid outflow is on by default).  Returns the index of the source in this shim's list. */
int ref_add_fluid_source_box(void *p, int outflow, const double lo[3], const double hi[3], const double vel[3]) {
    RefSim *h = (RefSim *)p;
    int idx = -1;
    guarded(h, [&] {
        FluidSimulation *s = h->sim;
        vmath::vec3 q((float)lo[0], (float)lo[1], (float)lo[2]);
        const double w = hi[0] - lo[0], ht = hi[1] - lo[1], d = hi[2] - lo[2];
        TriangleMesh m;
        m.vertices = {vmath::vec3(q.x, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z), vmath::vec3(q.x + w, q.y, q.z + d),
                      vmath::vec3(q.x, q.y, q.z + d), vmath::vec3(q.x, q.y + ht, q.z), vmath::vec3(q.x + w, q.y + ht, q.z),
                      vmath::vec3(q.x + w, q.y + ht, q.z + d), vmath::vec3(q.x, q.y + ht, q.z + d)};
        m.triangles = {Triangle(0, 1, 2), Triangle(0, 2, 3), Triangle(4, 7, 6), Triangle(4, 6, 5), Triangle(0, 3, 7), Triangle(0, 7, 4),
                       Triangle(1, 5, 6), Triangle(1, 6, 2), Triangle(0, 4, 5), Triangle(0, 5, 1), Triangle(3, 2, 6), Triangle(3, 6, 7)};
        MeshFluidSource *src = new MeshFluidSource(s->_isize, s->_jsize, s->_ksize, s->_dx);
        src->updateMeshStatic(m);
        if (outflow) src->setOutflow(); else src->setInflow();
        src->setVelocity(vmath::vec3((float)vel[0], (float)vel[1], (float)vel[2]));
        s->addMeshFluidSource(src);
        h->sources.push_back(src);
        idx = (int)h->sources.size() - 1;
    });
    return idx;
}
void ref_constrain_fluid_source_velocity(void *p, int idx, int on) {
    RefSim *h = (RefSim *)p;
    if (idx >= 0 && idx < (int)h->sources.size()) {
        if (on) h->sources[idx]->enableConstrainedFluidVelocity(); else h->sources[idx]->disableConstrainedFluidVelocity();
    }
}
void ref_enable_fluid_source(void *p, int idx, int on) {
    RefSim *h = (RefSim *)p;
    if (idx >= 0 && idx < (int)h->sources.size()) { if (on) h->sources[idx]->enable(); else h->sources[idx]->disable(); }
}

/* The surface FluidSimulation::getIsomesh() would hold for the CURRENT particles: the body of _outputSurfaceMeshThread
 * (fluidsimulation.cpp:5150-5218) run synchronously -- _polygonizeOutputSurface (ParticleMesher::meshParticles),
 * removeMinimumTriangleCountPolyhedra, _removeMeshNearDomain, smooth, _invertContactNormals -- with the engine's
 * settings except the subdivision level and the smoothing iterations given here (iterations < 0: the engine's). */
int ref_isomesh(void *p, int subdivisions, int smoothIterations, int *nv, int *nt) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        FluidSimulation *s = h->sim;
        std::vector<vmath::vec3> particles;
        particles.reserve(s->_markerParticles.size());
        for (size_t i = 0; i < s->_markerParticles.size(); i++) particles.push_back(s->_markerParticles[i].position);
        MeshLevelSet solid;
        solid.constructMinimalSignedDistanceField(s->_solidSDF);
        const int keepSub = s->_outputFluidSurfaceSubdivisionLevel, keepIt = s->_surfaceReconstructionSmoothingIterations;
        s->_outputFluidSurfaceSubdivisionLevel = subdivisions;
        if (smoothIterations >= 0) s->_surfaceReconstructionSmoothingIterations = smoothIterations;
        TriangleMesh isomesh, preview;
        s->_polygonizeOutputSurface(isomesh, preview, &particles, &solid);
        isomesh.removeMinimumTriangleCountPolyhedra(s->_minimumSurfacePolyhedronTriangleCount);
        s->_removeMeshNearDomain(isomesh);
        s->_smoothSurfaceMesh(isomesh);
        s->_invertContactNormals(isomesh);
        s->_outputFluidSurfaceSubdivisionLevel = keepSub;
        s->_surfaceReconstructionSmoothingIterations = keepIt;
        h->isomesh = isomesh;
        *nv = (int)isomesh.vertices.size();
        *nt = (int)isomesh.triangles.size();
    });
}

void ref_get_isomesh(void *p, float *verts, int *tris) {
    RefSim *h = (RefSim *)p;
    for (size_t i = 0; i < h->isomesh.vertices.size(); i++) {
        verts[3 * i] = h->isomesh.vertices[i].x; verts[3 * i + 1] = h->isomesh.vertices[i].y; verts[3 * i + 2] = h->isomesh.vertices[i].z;
    }
    for (size_t i = 0; i < h->isomesh.triangles.size(); i++)
        for (int q = 0; q < 3; q++) tris[3 * i + q] = h->isomesh.triangles[i].tri[q];
}

/* The mesher's scalar field of the current particles (ParticleMesher::_computeScalarField, one compute chunk), as the
 * polygonizer reads it (ScalarField::getScalarFieldValue: solid nodes clamped to the threshold): (K s+1)(J s+1)(I s+1)
 * floats, i fastest. */
int ref_mesher_scalar_field(void *p, int subdivisions, float *out) {
    RefSim *h = (RefSim *)p;
    return guarded(h, [&] {
        FluidSimulation *s = h->sim;
        std::vector<vmath::vec3> particles;
        for (size_t i = 0; i < s->_markerParticles.size(); i++) particles.push_back(s->_markerParticles[i].position);
        MeshLevelSet solid;
        solid.constructMinimalSignedDistanceField(s->_solidSDF);
        ParticleMesherParameters params;
        params.isize = s->_isize; params.jsize = s->_jsize; params.ksize = s->_ksize; params.dx = s->_dx;
        params.subdivisions = subdivisions;
        params.computechunks = 1;
        params.radius = s->_markerParticleRadius * s->_markerParticleScale;
        params.particles = &particles;
        params.solidSDF = &solid;
        params.isPreviewMesherEnabled = false;
        ParticleMesher mesher;
        mesher._initialize(params);
        ParticleMesher::MesherComputeChunkData data;
        mesher._generateComputeChunkData(data);
        if (data.computeChunks.size() != 1) throw std::runtime_error("expected one compute chunk");
        ParticleMesher::ScalarFieldData fieldData;
        mesher._initializeScalarFieldData(data.computeChunks[0], data, fieldData);
        const int ni = s->_isize * subdivisions + 1, nj = s->_jsize * subdivisions + 1, nk = s->_ksize * subdivisions + 1;
        if (fieldData.particles.empty()) {
            for (size_t q = 0; q < (size_t)ni * nj * nk; q++) out[q] = 0.0f;
            return;
        }
        mesher._computeScalarField(fieldData);
        // the chunk spans the bounding range of the blocks near particles; outside it the field keeps its fill value
        for (size_t q = 0; q < (size_t)ni * nj * nk; q++) out[q] = -mesher._getMaxDistanceValue();
        const ParticleMesher::MesherComputeChunk &c = data.computeChunks[0];
        for (int k = 0; k < c.ksize; k++)
            for (int j = 0; j < c.jsize; j++)
                for (int i = 0; i < c.isize; i++)
                    out[(size_t)(i + c.minGridIndex.i) + (size_t)ni * ((j + c.minGridIndex.j) + (size_t)nj * (k + c.minGridIndex.k))] =
                        (float)fieldData.fieldValues.getScalarFieldValue(i, j, k);
    });
}

}  // extern "C"
