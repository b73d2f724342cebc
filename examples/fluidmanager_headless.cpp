// The reference's FluidManager scene (src/FluidManager.cpp:47-69 with the sizes of src/main.cpp:12-16) run
// headless -- no DXViewer, no Win32 -- on libflip_b200 through the C++ façade: BASELINE config 1.
//   fluidmanager_headless [frames=100] [isize=30] [device=0]
// Prints one line per 10 frames and a summary (ms per frame, particles); exit code 0 on success, 2 when no
// CUDA device is present (there is no CPU fallback).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "fluidsimulation_b200.hpp"      // the ONLY change against the reference's includes (engine/fluidsimulation.h etc.)

class FluidManager {      // the members of the reference's FluidManager that do not draw; bodies copied call for call
public:
    FluidManager(int isize, int jsize, int ksize, double dx, double timeStep, int device)
        : _isize(isize), _jsize(jsize), _ksize(ksize), _dx(dx), _timeStep(timeStep) {
        _fluidsim = new FluidSimulation(_isize, _jsize, _ksize, _dx, device);
    }
    ~FluidManager() { delete _fluidsim; }

    TriangleMesh getTriangleMeshFromAABB(AABB bbox) {     // FluidManager::getTriangleMeshFromAABB  :20-44
        vmath::vec3 p = bbox.position;
        std::vector<vmath::vec3> verts{
            vmath::vec3(p.x, p.y, p.z),
            vmath::vec3(p.x + bbox.width, p.y, p.z),
            vmath::vec3(p.x + bbox.width, p.y, p.z + bbox.depth),
            vmath::vec3(p.x, p.y, p.z + bbox.depth),
            vmath::vec3(p.x, p.y + bbox.height, p.z),
            vmath::vec3(p.x + bbox.width, p.y + bbox.height, p.z),
            vmath::vec3(p.x + bbox.width, p.y + bbox.height, p.z + bbox.depth),
            vmath::vec3(p.x, p.y + bbox.height, p.z + bbox.depth)
        };
        std::vector<Triangle> tris{
            Triangle(0, 1, 2), Triangle(0, 2, 3), Triangle(4, 7, 6), Triangle(4, 6, 5),
            Triangle(0, 3, 7), Triangle(0, 7, 4), Triangle(1, 5, 6), Triangle(1, 6, 2),
            Triangle(0, 4, 5), Triangle(0, 5, 1), Triangle(3, 2, 6), Triangle(3, 6, 7)
        };
        TriangleMesh m;
        m.vertices = verts;
        m.triangles = tris;
        return m;
    }

    void initialize() {                                   // FluidManager::initialize  :47-69
        _fluidsim->setSurfaceSubdivisionLevel(2);

        double x, y, z;
        _fluidsim->getSimulationDimensions(&x, &y, &z);

        double boxWidth = (1.0 / 3.0) * x;
        double boxHeight = (1.0 / 3.0) * y;
        double boxDepth = (1.0 / 3.0) * z;
        vmath::vec3 boxPosition(0.5 * (x - boxWidth), 0.5 * (y - boxHeight), 0.5 * (z - boxDepth));
        AABB box(boxPosition, boxWidth, boxHeight, boxDepth);
        TriangleMesh boxMesh = getTriangleMeshFromAABB(box);
        MeshObject boxFluidObject(_isize, _jsize, _ksize, _dx);
        boxFluidObject.updateMeshStatic(boxMesh);
        _fluidsim->addMeshFluid(boxFluidObject);

        _fluidsim->addBodyForce(0.0, -25.0, 0.0);
        _fluidsim->initialize();
    }
    void iUpdate() {                                      // FluidManager::iUpdate  :74-83
        _simFrame = _fluidsim->getCurrentFrame();
        auto t0 = std::chrono::steady_clock::now();
        _fluidsim->update(_timeStep);
        _simTime += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    size_t iGetVertices() {                               // FluidManager::iGetVertices  :102-150, without the D3D vertex type
        _normal.clear();
        TriangleMesh isomesh = _fluidsim->getIsomesh();
        for (size_t i = 0; i < isomesh.vertices.size(); i++) _normal.push_back(vmath::vec3(0.0f, 0.0f, 0.0f));
        for (size_t i = 0; i < isomesh.triangles.size(); i++) {
            int i0 = isomesh.triangles[i].tri[0];
            int i1 = isomesh.triangles[i].tri[1];
            int i2 = isomesh.triangles[i].tri[2];
            vmath::vec3 v0 = isomesh.vertices[i0];
            vmath::vec3 v1 = isomesh.vertices[i1];
            vmath::vec3 v2 = isomesh.vertices[i2];
            vmath::vec3 e0 = v1 - v0;
            vmath::vec3 e1 = v2 - v0;
            vmath::vec3 faceN = vmath::cross(e0, e1);
            _normal[i0] += faceN;
            _normal[i1] += faceN;
            _normal[i2] += faceN;
        }
        for (size_t i = 0; i < isomesh.vertices.size(); i++) _normal[i] = vmath::normalize(_normal[i]);
        _triangles = isomesh.triangles.size();
        return isomesh.vertices.size();
    }
    FluidSimulation *sim() { return _fluidsim; }
    double simTimeMs() const { return _simTime; }
    int simFrame() const { return _simFrame; }
    size_t triangles() const { return _triangles; }

private:
    FluidSimulation *_fluidsim = nullptr;
    int _isize, _jsize, _ksize;
    double _dx, _timeStep, _simTime = 0.0;
    int _simFrame = 0;
    std::vector<vmath::vec3> _normal;
    size_t _triangles = 0;
};

int main(int argc, char **argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 100;
    const int n = argc > 2 ? atoi(argv[2]) : 30;
    const int device = argc > 3 ? atoi(argv[3]) : 0;
    const double dx = 0.125, timestep = 1.0 / 30.0;       // main.cpp:12-16, FPS30_D
    try {
        FluidManager fm(n, n, n, dx, timestep, device);
        fm.initialize();
        std::printf("initialized: %d^3 cells, dx %.3f, %u marker particles\n", n, dx, fm.sim()->getNumMarkerParticles());
        for (int f = 0; f < frames; f++) {
            fm.iUpdate();
            if ((f + 1) % 10 == 0 || f + 1 == frames) {
                const int ns = fm.sim()->getNumSubsteps();
                flip_step_stats s = fm.sim()->getStepStats(ns - 1);
                std::printf("frame %4d  substeps %d  particles %d  fluid cells %d  pcg iterations %d\n", fm.sim()->getCurrentFrame(), ns,
                            s.particles, s.fluid_cells, s.pcg_iterations);
            }
        }
        std::vector<flipb200::MarkerParticle> p = fm.sim()->getMarkerParticles();
        double ymin = 1e30, ymax = -1e30;
        for (const auto &q : p) { ymin = q.position.y < ymin ? q.position.y : ymin; ymax = q.position.y > ymax ? q.position.y : ymax; }
        std::printf("done: %d frames, %.3f ms per frame (wall, host side included), %zu particles, y in [%.4f, %.4f]\n", frames,
                    fm.simTimeMs() / frames, p.size(), ymin, ymax);
        const size_t nv = fm.iGetVertices();
        std::printf("isomesh: %zu vertices, %zu triangles (subdivision level %d)\n", nv, fm.triangles(), fm.sim()->getSurfaceSubdivisionLevel());
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "fluidmanager_headless: %s\n", e.what());
        return 2;
    }
}
