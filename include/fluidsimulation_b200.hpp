// C++ façade over the C-ABI of flip_b200.h with the member names of the reference's FluidSimulation
// (src/engine/fluidsimulation.h) for the calls on the hot path: the ones FluidManager makes
// (src/FluidManager.cpp:12,51-67,76-79: constructor, setSurfaceSubdivisionLevel, getSimulationDimensions,
// addMeshFluid, addBodyForce, initialize, getCurrentFrame, update) and the ones the north star adds
// (getMarkerParticles, getNumMarkerParticles, loadMarkerParticleData, getVelocityField, snapshot blobs).
// Header only; link with -lflip_b200.  Error behaviour: the exception types the reference throws for the same
// misuse (std::runtime_error: update before initialize, fluidsimulation.cpp:5756; std::domain_error: dt < 0, :5761;
// std::out_of_range: bad ranges, :2059).  There is no CPU fallback: without a CUDA device the constructor throws.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "flip_b200.h"

namespace flipb200 {

// MarkerParticle (markerparticle.h:30-42): two vmath::vec3 of floats, 24 bytes -- the AoS wire format
struct vec3 { float x = 0.f, y = 0.f, z = 0.f; };
struct MarkerParticle { vec3 position, velocity; };
static_assert(sizeof(MarkerParticle) == 24, "MarkerParticle must be six packed floats");

// FluidSimulationMarkerParticleData (fluidsimulation.h:102-106): float xyz triplets behind char pointers
struct FluidSimulationMarkerParticleData {
    int size = 0;
    char *positions = nullptr;
    char *velocities = nullptr;
};

class FluidSimulation {
public:
    FluidSimulation(int isize, int jsize, int ksize, double dx, int device = 0) : _isize(isize), _jsize(jsize), _ksize(ksize), _dx(dx) {
        int rc = flip_create(&_c, isize, jsize, ksize, dx, device);
        if (rc != FLIP_OK) {
            std::string m = flip_create_error();
            if (rc == FLIP_ERR_DOMAIN) throw std::domain_error(m);
            throw std::runtime_error(m);
        }
    }
    ~FluidSimulation() { flip_destroy(_c); }
    FluidSimulation(const FluidSimulation &) = delete;
    FluidSimulation &operator=(const FluidSimulation &) = delete;

    void getGridDimensions(int *i, int *j, int *k) const { *i = _isize; *j = _jsize; *k = _ksize; }
    double getCellSize() const { return _dx; }
    void getSimulationDimensions(double *w, double *h, double *d) const { *w = _isize * _dx; *h = _jsize * _dx; *d = _ksize * _dx; }
    void setSurfaceSubdivisionLevel(int) {}          // surface reconstruction is outside the hot path
    void addBodyForce(double fx, double fy, double fz) { check(flip_add_body_force(_c, fx, fy, fz)); }
    // addMeshFluid(MeshObject) for the axis-aligned box FluidManager builds (FluidManager.cpp:56-64)
    void addMeshFluidBox(const double lo[3], const double hi[3], const double velocity[3]) { check(flip_add_fluid_box(_c, lo, hi, velocity)); }
    void loadMarkerParticleData(FluidSimulationMarkerParticleData d) {
        check(flip_load_particles(_c, d.size, reinterpret_cast<const float *>(d.positions), reinterpret_cast<const float *>(d.velocities)));
    }
    void setPICFLIPRatio(double r) { check(flip_set_pic_flip_ratio(_c, r)); }
    void setCFLConditionNumber(double n) { check(flip_set_cfl(_c, n)); }
    void initialize() { check(flip_initialize(_c)); }
    void update(double dt) { check(flip_update(_c, dt)); }
    int getCurrentFrame() { int f = 0; check(flip_get_current_frame(_c, &f)); return f; }
    void setCurrentFrame(int f) { check(flip_set_current_frame(_c, f)); }
    unsigned int getNumMarkerParticles() { int n = 0; check(flip_get_num_particles(_c, &n)); return (unsigned int)n; }
    std::vector<MarkerParticle> getMarkerParticles() {         // by value, as the reference does (:2053-2073)
        std::vector<MarkerParticle> p(getNumMarkerParticles());
        if (!p.empty()) check(flip_get_particles(_c, reinterpret_cast<float *>(p.data()), (int)p.size()));
        return p;
    }
    // getMarkerParticlePositionData / VelocityData (:2408-2420): float xyz triplets
    unsigned int getMarkerParticlePositionDataSize() { return getNumMarkerParticles() * 3u * (unsigned int)sizeof(float); }
    void getMarkerParticlePositionData(char *data) { check(flip_get_particle_positions(_c, reinterpret_cast<float *>(data), (int)getNumMarkerParticles())); }
    void getMarkerParticleVelocityData(char *data) { check(flip_get_particle_velocities(_c, reinterpret_cast<float *>(data), (int)getNumMarkerParticles())); }
    // MACVelocityField raw arrays (macvelocityfield.cpp:100-110): U (i+1,j,k), V (i,j+1,k), W (i,j,k+1), i fastest
    void getVelocityField(std::vector<float> &U, std::vector<float> &V, std::vector<float> &W) {
        U.resize((size_t)(_isize + 1) * _jsize * _ksize);
        V.resize((size_t)_isize * (_jsize + 1) * _ksize);
        W.resize((size_t)_isize * _jsize * (_ksize + 1));
        check(flip_get_velocity_field(_c, U.data(), V.data(), W.data()));
    }
    // per-substep log integers of the last frame (_logStepInfo :5710-5745)
    int getNumSubsteps() { int n = 0; check(flip_get_num_substeps(_c, &n)); return n; }
    flip_step_stats getStepStats(int substep) { flip_step_stats s; check(flip_get_step_stats(_c, substep, &s)); return s; }
    flip_ctx *handle() { return _c; }

private:
    void check(int rc) {
        if (rc == FLIP_OK) return;
        std::string m = flip_last_error(_c);
        if (rc == FLIP_ERR_DOMAIN) throw std::domain_error(m);
        if (rc == FLIP_ERR_OUT_OF_RANGE) throw std::out_of_range(m);
        throw std::runtime_error(m);
    }
    flip_ctx *_c = nullptr;
    int _isize, _jsize, _ksize;
    double _dx;
};

}  // namespace flipb200
