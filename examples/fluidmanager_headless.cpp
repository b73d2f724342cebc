// The reference's FluidManager scene (src/FluidManager.cpp:47-69 with the sizes of src/main.cpp:12-16) run
// headless -- no DXViewer, no Win32 -- on libflip_b200 through the C++ façade: BASELINE config 1.
//   fluidmanager_headless [frames=100] [isize=30] [device=0]
// Prints one line per 10 frames and a summary (ms per frame, particles); exit code 0 on success, 2 when no
// CUDA device is present (there is no CPU fallback).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "fluidsimulation_b200.hpp"

using flipb200::FluidSimulation;

class FluidManager {      // the members of the reference's FluidManager that do not draw
public:
    FluidManager(int isize, int jsize, int ksize, double dx, double timeStep, int device)
        : _isize(isize), _jsize(jsize), _ksize(ksize), _dx(dx), _timeStep(timeStep) {
        _fluidsim = new FluidSimulation(_isize, _jsize, _ksize, _dx, device);
    }
    ~FluidManager() { delete _fluidsim; }
    void initialize() {                                   // FluidManager::initialize  :47-69
        _fluidsim->setSurfaceSubdivisionLevel(2);
        double x, y, z;
        _fluidsim->getSimulationDimensions(&x, &y, &z);
        double boxWidth = (1.0 / 3.0) * x, boxHeight = (1.0 / 3.0) * y, boxDepth = (1.0 / 3.0) * z;
        double lo[3] = {0.5 * (x - boxWidth), 0.5 * (y - boxHeight), 0.5 * (z - boxDepth)};
        double hi[3] = {lo[0] + boxWidth, lo[1] + boxHeight, lo[2] + boxDepth};
        double v[3] = {0.0, 0.0, 0.0};
        _fluidsim->addMeshFluidBox(lo, hi, v);
        _fluidsim->addBodyForce(0.0, -25.0, 0.0);
        _fluidsim->initialize();
    }
    void iUpdate() {                                      // FluidManager::iUpdate  :74-83
        _simFrame = _fluidsim->getCurrentFrame();
        auto t0 = std::chrono::steady_clock::now();
        _fluidsim->update(_timeStep);
        _simTime += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    FluidSimulation *sim() { return _fluidsim; }
    double simTimeMs() const { return _simTime; }
    int simFrame() const { return _simFrame; }

private:
    FluidSimulation *_fluidsim = nullptr;
    int _isize, _jsize, _ksize;
    double _dx, _timeStep, _simTime = 0.0;
    int _simFrame = 0;
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
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "fluidmanager_headless: %s\n", e.what());
        return 2;
    }
}
