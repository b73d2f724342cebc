// C++ façade over the C-ABI of flip_b200.h with the type and member names of the reference for the calls on the hot
// path: everything FluidManager touches (src/FluidManager.cpp:12,47-83,102-171: FluidSimulation constructor,
// setSurfaceSubdivisionLevel, getSimulationDimensions, AABB / TriangleMesh / MeshObject::updateMeshStatic,
// addMeshFluid(MeshObject), addBodyForce, initialize, getCurrentFrame, update, getIsomesh) and what the north star
// adds (addMarkerParticle, getMarkerParticles, getNumMarkerParticles, loadMarkerParticleData, getVelocityField with the
// MACVelocityField raw arrays, snapshot blobs, the per-stage timing buckets).  The call sequence of
// FluidManager::initialize / iUpdate compiles against this header unchanged (examples/fluidmanager_headless.cpp).
// Header only; link with -lflip_b200.  Error behaviour: the exception types the reference throws for the same
// misuse (std::runtime_error: update before initialize, fluidsimulation.cpp:5756; std::domain_error: dt < 0, :5761;
// std::out_of_range: bad ranges, :2059).  There is no CPU fallback: without a CUDA device the constructor throws.
//
// The names live in namespace flipb200 and are also exported to the global namespace (vmath::vec3, AABB, Triangle,
// TriangleMesh, MeshObject, MACVelocityField, MarkerParticle, FluidSimulation) unless FLIPB200_NO_GLOBAL_NAMES is
// defined before the include.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "flip_b200.h"

namespace flipb200 {

namespace vmath {      // vmath.h:33-120, the members FluidManager uses
struct vec3 {
    float x = 0.f, y = 0.f, z = 0.f;
    vec3() = default;
    template <class A, class B, class C>      // (float and double arguments mix freely, as with the reference's overloads)
    vec3(A xx, B yy, C zz) : x((float)xx), y((float)yy), z((float)zz) {}
    vec3 &operator+=(const vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline float dot(const vec3 &a, const vec3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3 &a, const vec3 &b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(const vec3 &a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(const vec3 &a) { float l = length(a); return l > 0.f ? vec3(a.x / l, a.y / l, a.z / l) : a; }
}  // namespace vmath

// MarkerParticle (markerparticle.h:30-42): two vmath::vec3 of floats, 24 bytes -- the AoS wire format
struct MarkerParticle { vmath::vec3 position, velocity; };
static_assert(sizeof(MarkerParticle) == 24, "MarkerParticle must be six packed floats");

// FluidSimulationMarkerParticleData (fluidsimulation.h:102-106): float xyz triplets behind char pointers
struct FluidSimulationMarkerParticleData {
    int size = 0;
    char *positions = nullptr;
    char *velocities = nullptr;
};

// aabb.h:36-70
struct AABB {
    vmath::vec3 position;
    double width = 0.0, height = 0.0, depth = 0.0;
    AABB() = default;
    AABB(vmath::vec3 p, double w, double h, double d) : position(p), width(w), height(h), depth(d) {}
};

// triangle.h:33-46, trianglemesh.h:50-130 (the two containers FluidManager reads and writes)
struct Triangle {
    int tri[3] = {0, 0, 0};
    Triangle() = default;
    Triangle(int a, int b, int c) { tri[0] = a; tri[1] = b; tri[2] = c; }
};
struct TriangleMesh {
    std::vector<vmath::vec3> vertices;
    std::vector<Triangle> triangles;
};

// MeshObject (meshobject.h:60-170) for static closed meshes: what addMeshFluid needs of it is the cell range
// (getCells, meshobject.cpp:101-140) and the signed distance field that FluidSimulation computes from getMesh()
// (MeshLevelSet::fastCalculateSignedDistanceField, fluidsimulation.cpp:4743-4744).  Here the nodal field is computed on
// the host when the object is queued -- exact point-triangle distances within `band` cells of the mesh's bounding box,
// the sign by crossing parity along +x -- and handed to the device seeding (flip_add_fluid_sdf).  An axis-aligned
// box (the object FluidManager builds) is recognised and handed over analytically (flip_add_fluid_box).
class MeshObject {
public:
    MeshObject() = default;
    MeshObject(int isize, int jsize, int ksize, double dx) : _isize(isize), _jsize(jsize), _ksize(ksize), _dx(dx) {}
    void getGridDimensions(int *i, int *j, int *k) const { *i = _isize; *j = _jsize; *k = _ksize; }
    void updateMeshStatic(TriangleMesh meshCurrent) { _mesh = std::move(meshCurrent); _animated = false; }
    // MeshObject::updateMeshAnimated (meshobject.cpp:61-95), once per frame before update(): the previous, the current and
    // the next frame's mesh (fixed topology).  An axis-aligned box that translates rigidly goes through
    // flip_set_obstacle_box_motion, any other closed mesh (turning, deforming) through flip_set_obstacle_mesh_motion with
    // per-vertex velocities.  May be called before or after addMeshObstacle (an object added as a static box has to keep
    // translating as a box).
    void updateMeshAnimated(TriangleMesh meshPrevious, TriangleMesh meshCurrent, TriangleMesh meshNext) {
        if (meshPrevious.vertices.size() != meshCurrent.vertices.size() || meshNext.vertices.size() != meshCurrent.vertices.size())
            throw std::runtime_error("Error: animated mesh objects must keep their topology.\n");
        _meshPrev = std::move(meshPrevious);
        _meshNext = std::move(meshNext);
        _mesh = std::move(meshCurrent);
        MeshObject prev(_isize, _jsize, _ksize, _dx), next(_isize, _jsize, _ksize, _dx);
        prev._mesh = _meshPrev;
        next._mesh = _meshNext;
        _boxMotion = isAxisAlignedBox() && prev.isAxisAlignedBox() && next.isAxisAlignedBox();
        if (_ctx && _addedAsBox && !_boxMotion)
            throw std::runtime_error("Error: an obstacle added as a box can only translate; add it with its animation set.\n");
        if (_ctx && !_addedAsBox) _boxMotion = false;
        if (_boxMotion) {
            vmath::vec3 lo, hi, plo, phi, nlo, nhi;
            bounds(lo, hi); prev.bounds(plo, phi); next.bounds(nlo, nhi);
            for (int a = 0; a < 3; a++) { _curLo[a] = (&lo.x)[a]; _prevLo[a] = (&plo.x)[a]; _nextLo[a] = (&nlo.x)[a]; }
        }
        _animated = true;
        pushMotion();
    }
    TriangleMesh getMesh() const { return _mesh; }
    // MeshObject::enable / disable / isEnabled (meshobject.cpp:267-283): acts on the device once the object is an obstacle
    void enable() { _enabled = true; if (_ctx) flip_enable_obstacle(_ctx, _obstacleId, 1); }
    void disable() { _enabled = false; if (_ctx) flip_enable_obstacle(_ctx, _obstacleId, 0); }
    bool isEnabled() const { return _enabled; }
    bool isAnimated() const { return _animated; }
    // MeshObject::setFriction / getFriction (meshobject.cpp:300-308), clamped to [0, 1]
    void setFriction(float f) {
        _friction = std::max(0.0f, std::min(f, 1.0f));
        if (_ctx && flip_set_obstacle_friction(_ctx, _obstacleId, _friction) != FLIP_OK) throw std::runtime_error(flip_last_error(_ctx));
    }
    float getFriction() const { return _friction; }

    void bounds(vmath::vec3 &lo, vmath::vec3 &hi) const {
        lo = vmath::vec3(1e30f, 1e30f, 1e30f); hi = vmath::vec3(-1e30f, -1e30f, -1e30f);
        for (const auto &v : _mesh.vertices) {
            lo = vmath::vec3(std::min(lo.x, v.x), std::min(lo.y, v.y), std::min(lo.z, v.z));
            hi = vmath::vec3(std::max(hi.x, v.x), std::max(hi.y, v.y), std::max(hi.z, v.z));
        }
    }
    // eight vertices on the corners of their bounding box, twelve triangles: the mesh of getTriangleMeshFromAABB
    bool isAxisAlignedBox() const {
        if (_mesh.vertices.size() != 8 || _mesh.triangles.size() != 12) return false;
        vmath::vec3 lo, hi;
        bounds(lo, hi);
        unsigned seen = 0;
        for (const auto &v : _mesh.vertices) {
            const bool x0 = v.x == lo.x, x1 = v.x == hi.x, y0 = v.y == lo.y, y1 = v.y == hi.y, z0 = v.z == lo.z, z1 = v.z == hi.z;
            if (!((x0 || x1) && (y0 || y1) && (z0 || z1))) return false;
            seen |= 1u << ((x1 ? 1 : 0) | (y1 ? 2 : 0) | (z1 ? 4 : 0));
        }
        return seen == 0xffu;
    }
    // nodal signed distance field of the grid, (I+1)(J+1)(K+1) floats, i fastest; cells [lo,hi) worth scanning
    // (flip_mesh_sdf: what MeshObject::getMeshLevelSet hands the engine, meshobject.cpp:239)
    void signedDistanceField(std::vector<float> &phi, int lo[3], int hi[3], int band = 3, float farValue = 0.0f) const {
        phi.resize((size_t)(_isize + 1) * (_jsize + 1) * (_ksize + 1));
        static_assert(sizeof(vmath::vec3) == 12 && sizeof(Triangle) == 12, "packed float / int triplets");
        const int rc = flip_mesh_sdf(_isize, _jsize, _ksize, _dx, reinterpret_cast<const float *>(_mesh.vertices.data()),
                                     (int)_mesh.vertices.size(), reinterpret_cast<const int *>(_mesh.triangles.data()),
                                     (int)_mesh.triangles.size(), band, farValue, phi.data(), lo, hi);
        if (rc != FLIP_OK) throw std::domain_error("Error: bad triangle mesh for a signed distance field.\n");
    }

private:
    friend class FluidSimulation;
    int _isize = 0, _jsize = 0, _ksize = 0;
    double _dx = 0.0;
    TriangleMesh _mesh;
    bool _enabled = true;
    flip_ctx *_ctx = nullptr;     // the simulation this object is an obstacle of
    int _obstacleId = 0;
    bool _animated = false, _boxMotion = false, _addedAsBox = false;
    float _friction = 0.0f;
    TriangleMesh _meshPrev, _meshNext;
    double _baseLo[3] = {0, 0, 0};                     // lower corner of the box the obstacle was added as
    double _prevLo[3] = {0, 0, 0}, _curLo[3] = {0, 0, 0}, _nextLo[3] = {0, 0, 0};
    void pushMotion() {
        if (!_ctx || !_animated) return;
        if (!_boxMotion) {
            static_assert(sizeof(vmath::vec3) == 12, "packed float triplets");
            if (flip_set_obstacle_mesh_motion(_ctx, _obstacleId, reinterpret_cast<const float *>(_meshPrev.vertices.data()),
                                              reinterpret_cast<const float *>(_mesh.vertices.data()),
                                              reinterpret_cast<const float *>(_meshNext.vertices.data())) != FLIP_OK)
                throw std::runtime_error(flip_last_error(_ctx));
            return;
        }
        double a[3], b[3], c[3];
        for (int q = 0; q < 3; q++) { a[q] = _prevLo[q] - _baseLo[q]; b[q] = _curLo[q] - _baseLo[q]; c[q] = _nextLo[q] - _baseLo[q]; }
        if (flip_set_obstacle_box_motion(_ctx, _obstacleId, a, b, c) != FLIP_OK) throw std::runtime_error(flip_last_error(_ctx));
    }
};

// MeshFluidSource (meshfluidsource.h:40-117) for static closed meshes: an inflow that emits at the end of every substep
// or an outflow that removes the particles inside it.  enable() / disable() and removeMeshFluidSource act on the device
// object once the source has been added to a simulation.
class FluidSimulation;
class MeshFluidSource {
public:
    MeshFluidSource() = default;
    MeshFluidSource(int i, int j, int k, double dx) : _object(i, j, k, dx) {}
    void updateMeshStatic(TriangleMesh meshCurrent) { _object.updateMeshStatic(std::move(meshCurrent)); }
    void enable() { _enabled = true; push(); }
    void disable() { _enabled = false; push(); }
    bool isEnabled() const { return _enabled; }
    void setInflow() { _inflow = true; }
    bool isInflow() const { return _inflow; }
    void setOutflow() { _inflow = false; }
    bool isOutflow() const { return !_inflow; }
    void setVelocity(vmath::vec3 v) { _velocity = v; }
    vmath::vec3 getVelocity() const { return _velocity; }
    void enableConstrainedFluidVelocity() { _constrained = true; push(); }
    void disableConstrainedFluidVelocity() { _constrained = false; push(); }
    bool isConstrainedFluidVelocityEnabled() const { return _constrained; }
    MeshObject *getMeshObject() { return &_object; }

private:
    friend class FluidSimulation;
    void push() {
        if (!_ctx) return;
        flip_enable_fluid_source(_ctx, _id, _enabled ? 1 : 0);
        flip_constrain_fluid_source_velocity(_ctx, _id, _constrained ? 1 : 0);
    }
    MeshObject _object;
    bool _enabled = true, _inflow = true, _constrained = true;
    vmath::vec3 _velocity;
    flip_ctx *_ctx = nullptr;
    int _id = 0;
};

// MACVelocityField (macvelocityfield.cpp:46-54,100-110): the three raw face arrays, U (i+1,j,k), V (i,j+1,k),
// W (i,j,k+1), i fastest -- a host copy filled by FluidSimulation::getVelocityField()
class MACVelocityField {
public:
    void getGridDimensions(int *i, int *j, int *k) const { *i = _isize; *j = _jsize; *k = _ksize; }
    double getGridCellSize() const { return _dx; }
    float *getRawArrayU() { return _u.data(); }
    float *getRawArrayV() { return _v.data(); }
    float *getRawArrayW() { return _w.data(); }
    float U(int i, int j, int k) const { return _u[(size_t)i + (size_t)(_isize + 1) * (j + (size_t)_jsize * k)]; }
    float V(int i, int j, int k) const { return _v[(size_t)i + (size_t)_isize * (j + (size_t)(_jsize + 1) * k)]; }
    float W(int i, int j, int k) const { return _w[(size_t)i + (size_t)_isize * (j + (size_t)_jsize * k)]; }

private:
    friend class FluidSimulation;
    int _isize = 0, _jsize = 0, _ksize = 0;
    double _dx = 0.0;
    std::vector<float> _u, _v, _w;
};

// the per-stage wall-clock buckets of the reference's log (fluidsimulation.h:1172-1190), in seconds, of the last frame
struct TimingData {
    double updateObstacleObjects = 0.0, updateLiquidLevelSet = 0.0, advectVelocityField = 0.0, saveVelocityField = 0.0,
           calculateFluidCurvatureGrid = 0.0, applyBodyForcesToVelocityField = 0.0, applyViscosityToVelocityField = 0.0,
           pressureSolve = 0.0, constrainVelocityFields = 0.0, updateDiffuseMaterial = 0.0, updateSheetSeeding = 0.0,
           updateMarkerParticleVelocities = 0.0, deleteSavedVelocityField = 0.0, advanceMarkerParticles = 0.0,
           updateFluidObjects = 0.0, outputNonMeshSimulationData = 0.0, outputMeshSimulationData = 0.0, frameTime = 0.0;
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
    // surface reconstruction: the subdivision level of the mesher's scalar field (:1012-1023)
    void setSurfaceSubdivisionLevel(int n) {
        if (n < 1) throw std::domain_error("Error: subdivision level must be greater than or equal to 1.\n");
        _subdivisionLevel = n;
        flip_set_surface_subdivision_level(_c, n);
    }
    int getSurfaceSubdivisionLevel() const { return _subdivisionLevel; }
    void addBodyForce(double fx, double fy, double fz) { check(flip_add_body_force(_c, fx, fy, fz)); }
    // addMeshFluid (:1573-1590): queued, seeded on the device at the end of the next substep
    void addMeshFluid(MeshObject fluid) { addMeshFluid(fluid, vmath::vec3(0.f, 0.f, 0.f)); }
    void addMeshFluid(MeshObject fluid, vmath::vec3 velocity) {
        int i, j, k;
        fluid.getGridDimensions(&i, &j, &k);
        if (i != _isize || j != _jsize || k != _ksize) throw std::domain_error("Error: mesh object dimensions must be equal to simulation dimensions.\n");
        const double v[3] = {velocity.x, velocity.y, velocity.z};
        if (fluid.isAxisAlignedBox()) {
            vmath::vec3 lo, hi;
            fluid.bounds(lo, hi);
            const double l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
            check(flip_add_fluid_box(_c, l, h, v));
            return;
        }
        std::vector<float> phi;
        int clo[3], chi[3];
        fluid.signedDistanceField(phi, clo, chi);
        check(flip_add_fluid_sdf(_c, phi.data(), clo, chi, v));
    }
    void addMeshFluidBox(const double lo[3], const double hi[3], const double velocity[3]) { check(flip_add_fluid_box(_c, lo, hi, velocity)); }
    // addMeshFluidSource / removeMeshFluidSource (:1953-1985)
    void addMeshFluidSource(MeshFluidSource *source) {
        if (source->_ctx == _c) throw std::runtime_error("Error: Mesh fluid source has already been added.\n");
        const vmath::vec3 vv = source->getVelocity();
        const double v[3] = {vv.x, vv.y, vv.z};
        int id = 0;
        if (source->_object.isAxisAlignedBox()) {
            vmath::vec3 lo, hi;
            source->_object.bounds(lo, hi);
            const double l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
            check(flip_add_fluid_source_box(_c, source->isOutflow() ? 1 : 0, l, h, v, &id));
        } else {
            std::vector<float> phi;
            int clo[3], chi[3];
            source->_object.signedDistanceField(phi, clo, chi);
            vmath::vec3 lo, hi;
            source->_object.bounds(lo, hi);
            const double l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
            check(flip_add_fluid_source_sdf(_c, source->isOutflow() ? 1 : 0, phi.data(), clo, chi, l, h, v, &id));
        }
        source->_ctx = _c; source->_id = id;
        _sources.push_back(source);
        source->push();
    }
    void removeMeshFluidSource(MeshFluidSource *source) {
        if (source->_ctx != _c) throw std::runtime_error("Error: could not find mesh fluid source to remove.\n");
        check(flip_remove_fluid_source(_c, source->_id));
        source->_ctx = nullptr;
        for (size_t q = 0; q < _sources.size(); q++)
            if (_sources[q] == source) { _sources.erase(_sources.begin() + q); break; }
    }
    void removeMeshFluidSources() {                                // :1987-1992
        while (!_sources.empty()) removeMeshFluidSource(_sources.back());
    }
    // setBoundaryFriction / getBoundaryFriction (:1743-1759): std::domain_error outside [0, 1]
    void setBoundaryFriction(double f) { check(flip_set_boundary_friction(_c, f)); _boundaryFriction = f; }
    double getBoundaryFriction() const { return _boundaryFriction; }
    // addMeshObstacle / removeMeshObstacle (:1994-2031): merged into the solid SDF on the device side
    void addMeshObstacle(MeshObject *obstacle) {
        if (obstacle->_ctx == _c) throw std::runtime_error("Error: mesh obstacle has already been added.\n");
        int id = 0;
        if (obstacle->isAxisAlignedBox() && (!obstacle->_animated || obstacle->_boxMotion)) {
            vmath::vec3 lo, hi;
            obstacle->bounds(lo, hi);
            const double l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
            check(flip_add_obstacle_box(_c, l, h, &id));
            for (int a = 0; a < 3; a++) obstacle->_baseLo[a] = l[a];
            obstacle->_addedAsBox = true;
        } else {
            // the library keeps the mesh (its signed distance field in the band of three cells, _solidLevelSetExactBand, is
            // flip_mesh_sdf's) so that it can be animated later
            static_assert(sizeof(vmath::vec3) == 12 && sizeof(Triangle) == 12, "packed float / int triplets");
            const TriangleMesh &m = obstacle->_mesh;
            check(flip_add_obstacle_mesh(_c, reinterpret_cast<const float *>(m.vertices.data()), (int)m.vertices.size(),
                                         reinterpret_cast<const int *>(m.triangles.data()), (int)m.triangles.size(), &id));
            obstacle->_addedAsBox = false;
            obstacle->_boxMotion = false;
        }
        obstacle->_ctx = _c; obstacle->_obstacleId = id;
        if (obstacle->_friction != 0.0f) check(flip_set_obstacle_friction(_c, id, obstacle->_friction));
        obstacle->pushMotion();                                    // an animation set before the object was added
        _obstacles.push_back(obstacle);
        if (!obstacle->isEnabled()) check(flip_enable_obstacle(_c, id, 0));
    }
    void removeMeshObstacle(MeshObject *obstacle) {
        if (obstacle->_ctx != _c) throw std::invalid_argument("Error: could not find mesh obstacle to remove.\n");
        check(flip_remove_obstacle(_c, obstacle->_obstacleId));
        obstacle->_ctx = nullptr;
        for (size_t q = 0; q < _obstacles.size(); q++)
            if (_obstacles[q] == obstacle) { _obstacles.erase(_obstacles.begin() + q); break; }
    }
    void removeMeshObstacles() {                                   // :2032-2036
        while (!_obstacles.empty()) removeMeshObstacle(_obstacles.back());
    }
    void addMarkerParticle(vmath::vec3 p, vmath::vec3 v = vmath::vec3()) {      // _addMarkerParticle :2637
        const float pp[3] = {p.x, p.y, p.z}, vv[3] = {v.x, v.y, v.z};
        check(flip_add_marker_particle(_c, pp, vv));
    }
    void loadMarkerParticleData(FluidSimulationMarkerParticleData d) {
        check(flip_load_particles(_c, d.size, reinterpret_cast<const float *>(d.positions), reinterpret_cast<const float *>(d.velocities)));
    }
    void setPICFLIPRatio(double r) { check(flip_set_pic_flip_ratio(_c, r)); _picflip = r; }
    double getPICFLIPRatio() const { return _picflip; }
    void setCFLConditionNumber(double n) { check(flip_set_cfl(_c, n)); _cfl = n; }
    double getCFLConditionNumber() const { return _cfl; }
    // setMin / MaxTimeStepsPerFrame (:1801-1831)
    void setMinTimeStepsPerFrame(int n) {
        if (n < 1) throw std::domain_error("Error: min step count must be greater than or equal to 1.\n");
        check(flip_set_substep_limits(_c, n, _maxSteps)); _minSteps = n;
    }
    void setMaxTimeStepsPerFrame(int n) {
        if (n < 1) throw std::domain_error("Error: max step count must be greater than or equal to 1.\n");
        check(flip_set_substep_limits(_c, _minSteps, n)); _maxSteps = n;
    }
    int getMinTimeStepsPerFrame() const { return _minSteps; }
    int getMaxTimeStepsPerFrame() const { return _maxSteps; }
    void resetBodyForce() { check(flip_reset_body_force(_c)); }
    void enableExtremeVelocityRemoval() { check(flip_set_extreme_velocity_removal(_c, 1)); _extreme = true; }
    void disableExtremeVelocityRemoval() { check(flip_set_extreme_velocity_removal(_c, 0)); _extreme = false; }
    bool isExtremeVelocityRemovalEnabled() const { return _extreme; }
    void setMarkerParticleScale(double s) { check(flip_set_marker_particle_scale(_c, s)); _particleScale = s; }
    double getMarkerParticleScale() const { return _particleScale; }
    // setSurfaceSmoothingValue / Iterations (fluidsimulation.h:1607-1608 defaults: 0.5, 2)
    void setSurfaceSmoothingValue(double v) { _smoothValue = v; check(flip_set_surface_smoothing(_c, _smoothValue, _smoothIterations)); }
    void setSurfaceSmoothingIterations(int n) { _smoothIterations = n; check(flip_set_surface_smoothing(_c, _smoothValue, _smoothIterations)); }
    double getSurfaceSmoothingValue() const { return _smoothValue; }
    int getSurfaceSmoothingIterations() const { return _smoothIterations; }
    // thread-count knobs of the host engine: accepted, without effect on the device path
    void setMaxThreadCount(int) {}
    int getGridWidth() const { return _isize; }
    int getGridHeight() const { return _jsize; }
    int getGridDepth() const { return _ksize; }
    double getSimulationWidth() const { return _isize * _dx; }
    double getSimulationHeight() const { return _jsize * _dx; }
    double getSimulationDepth() const { return _ksize * _dx; }
    void getVersion(int *major, int *minor, int *revision) const { *major = 1; *minor = 0; *revision = 9; }     // the restated engine
    bool isInitialized() const { return _initialized; }
    bool isCurrentFrameFinished() const { return true; }        // update() returns when the frame is done
    void initialize() { check(flip_initialize(_c)); _initialized = true; }
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
    unsigned int getMarkerParticleVelocityDataSize() { return getMarkerParticlePositionDataSize(); }
    // ...DataRange (:2314-2338): particles [start, end) of the current order
    void getMarkerParticlePositionDataRange(int start_idx, int end_idx, char *data) { dataRange(start_idx, end_idx, data, true); }
    void getMarkerParticleVelocityDataRange(int start_idx, int end_idx, char *data) { dataRange(start_idx, end_idx, data, false); }
    // getVelocityField (:2045-2047): the MAC field as it stands after the last step
    MACVelocityField *getVelocityField() {
        _mac._isize = _isize; _mac._jsize = _jsize; _mac._ksize = _ksize; _mac._dx = _dx;
        _mac._u.resize((size_t)(_isize + 1) * _jsize * _ksize);
        _mac._v.resize((size_t)_isize * (_jsize + 1) * _ksize);
        _mac._w.resize((size_t)_isize * _jsize * (_ksize + 1));
        check(flip_get_velocity_field(_c, _mac._u.data(), _mac._v.data(), _mac._w.data()));
        return &_mac;
    }
    void getVelocityField(std::vector<float> &U, std::vector<float> &V, std::vector<float> &W) {
        MACVelocityField *m = getVelocityField();
        U = m->_u; V = m->_v; W = m->_w;
    }
    // getIsomesh (fluidsimulation.h:1126): the surface of the liquid as of the last update(), reconstructed on the
    // device from the marker particles (ParticleMesher::meshParticles, particlemesher.cpp:36)
    TriangleMesh &getIsomesh() {
        int nv = 0, nt = 0;
        check(flip_get_isomesh_size(_c, &nv, &nt));
        _isomesh.vertices.resize((size_t)nv);
        _isomesh.triangles.resize((size_t)nt);
        if (nv > 0 || nt > 0)
            check(flip_get_isomesh(_c, reinterpret_cast<float *>(_isomesh.vertices.data()), reinterpret_cast<int *>(_isomesh.triangles.data())));
        return _isomesh;
    }
    // the stage buckets of the last substep under the reference's names, seconds (CUDA events on the context's stream)
    TimingData getTimingData() {
        float ms[FLIP_NUM_STAGES];
        check(flip_get_stage_times_ms(_c, ms));
        TimingData t;
        t.updateObstacleObjects = 1e-3 * ms[FLIP_STAGE_OBSTACLES];
        t.updateLiquidLevelSet = 1e-3 * ms[FLIP_STAGE_LIQUID_SDF];
        t.advectVelocityField = 1e-3 * (ms[FLIP_STAGE_P2G] + ms[FLIP_STAGE_EXTRAPOLATE_A]);
        t.saveVelocityField = 1e-3 * ms[FLIP_STAGE_SAVE];
        t.applyBodyForcesToVelocityField = 1e-3 * ms[FLIP_STAGE_BODY_FORCE];
        t.pressureSolve = 1e-3 * ms[FLIP_STAGE_PRESSURE];
        t.constrainVelocityFields = 1e-3 * (ms[FLIP_STAGE_EXTRAPOLATE_B] + ms[FLIP_STAGE_CONSTRAIN]);
        t.updateMarkerParticleVelocities = 1e-3 * ms[FLIP_STAGE_G2P];
        t.advanceMarkerParticles = 1e-3 * ms[FLIP_STAGE_ADVANCE];
        t.updateFluidObjects = 1e-3 * ms[FLIP_STAGE_TAIL];
        for (int s = 0; s < FLIP_NUM_STAGES; s++) t.frameTime += 1e-3 * ms[s];
        return t;
    }
    // per-substep log integers of the last frame (_logStepInfo :5710-5745)
    int getNumSubsteps() { int n = 0; check(flip_get_num_substeps(_c, &n)); return n; }
    flip_step_stats getStepStats(int substep) { flip_step_stats s; check(flip_get_step_stats(_c, substep, &s)); return s; }
    flip_ctx *handle() { return _c; }

private:
    void dataRange(int start_idx, int end_idx, char *data, bool positions) {
        const int n = (int)getNumMarkerParticles();
        if (start_idx < 0 || end_idx > n || start_idx > end_idx) throw std::domain_error("Error: invalid range.\n");
        std::vector<float> all((size_t)3 * n);
        if (n > 0) check(positions ? flip_get_particle_positions(_c, all.data(), n) : flip_get_particle_velocities(_c, all.data(), n));
        std::memcpy(data, all.data() + (size_t)3 * start_idx, sizeof(float) * 3 * (size_t)(end_idx - start_idx));
    }
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
    int _subdivisionLevel = 1;
    double _picflip = 0.05, _cfl = 5.0, _particleScale = 3.0, _smoothValue = 0.5;      // the engine's defaults (fluidsimulation.h)
    int _minSteps = 1, _maxSteps = 6, _smoothIterations = 2;
    bool _extreme = true, _initialized = false;
    std::vector<MeshObject *> _obstacles;
    double _boundaryFriction = 0.0;
    std::vector<MeshFluidSource *> _sources;
    MACVelocityField _mac;
    TriangleMesh _isomesh;
};

}  // namespace flipb200

#ifndef FLIPB200_NO_GLOBAL_NAMES
namespace vmath = flipb200::vmath;
using flipb200::AABB;
using flipb200::FluidSimulation;
using flipb200::FluidSimulationMarkerParticleData;
using flipb200::MACVelocityField;
using flipb200::MarkerParticle;
using flipb200::MeshFluidSource;
using flipb200::MeshObject;
using flipb200::Triangle;
using flipb200::TriangleMesh;
#endif
