// A scene a user of the reference would write against engine/fluidsimulation.h, built here over the façade: an inflow
// MeshFluidSource pouring onto a static ramp (a general closed mesh -> nodal signed distance field) and a box obstacle,
// an outflow along one wall.  Exercises addMeshObstacle / removeMeshObstacle, addMeshFluidSource, MeshObject::disable.
//   obstacles_and_sources [frames=40] [isize=40] [device=0]
// Exit code 0 on success, 2 when no CUDA device is present (there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fluidsimulation_b200.hpp"

static TriangleMesh boxMesh(vmath::vec3 p, float w, float h, float d) {
    TriangleMesh m;
    m.vertices = {vmath::vec3(p.x, p.y, p.z), vmath::vec3(p.x + w, p.y, p.z), vmath::vec3(p.x + w, p.y, p.z + d), vmath::vec3(p.x, p.y, p.z + d),
                  vmath::vec3(p.x, p.y + h, p.z), vmath::vec3(p.x + w, p.y + h, p.z), vmath::vec3(p.x + w, p.y + h, p.z + d), vmath::vec3(p.x, p.y + h, p.z + d)};
    m.triangles = {Triangle(0, 1, 2), Triangle(0, 2, 3), Triangle(4, 7, 6), Triangle(4, 6, 5), Triangle(0, 3, 7), Triangle(0, 7, 4),
                   Triangle(1, 5, 6), Triangle(1, 6, 2), Triangle(0, 4, 5), Triangle(0, 5, 1), Triangle(3, 2, 6), Triangle(3, 6, 7)};
    return m;
}

// a wedge (triangular prism): not an axis-aligned box, so it takes the signed-distance path
static TriangleMesh wedgeMesh(vmath::vec3 p, float w, float h, float d) {
    TriangleMesh m;
    m.vertices = {vmath::vec3(p.x, p.y, p.z), vmath::vec3(p.x + w, p.y, p.z), vmath::vec3(p.x, p.y + h, p.z),
                  vmath::vec3(p.x, p.y, p.z + d), vmath::vec3(p.x + w, p.y, p.z + d), vmath::vec3(p.x, p.y + h, p.z + d)};
    m.triangles = {Triangle(0, 2, 1), Triangle(3, 4, 5),                         // the two triangular ends
                   Triangle(0, 1, 4), Triangle(0, 4, 3),                         // bottom
                   Triangle(0, 3, 5), Triangle(0, 5, 2),                         // back
                   Triangle(1, 2, 5), Triangle(1, 5, 4)};                        // the slope
    return m;
}

int main(int argc, char **argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 40;
    const int n = argc > 2 ? atoi(argv[2]) : 40;
    const int device = argc > 3 ? atoi(argv[3]) : 0;
    const double dx = 0.125;
    try {
        FluidSimulation sim(n, n, n, dx, device);
        const float L = (float)(n * dx);
        MeshObject ramp(n, n, n, dx), block(n, n, n, dx);
        ramp.updateMeshStatic(wedgeMesh(vmath::vec3(0.30f * L, 0.04f * L, 0.2f * L), 0.4f * L, 0.3f * L, 0.6f * L));
        block.updateMeshStatic(boxMesh(vmath::vec3(0.72f * L, 0.04f * L, 0.35f * L), 0.12f * L, 0.2f * L, 0.3f * L));
        sim.addMeshObstacle(&ramp);
        sim.addMeshObstacle(&block);

        MeshFluidSource inflow(n, n, n, dx), drain(n, n, n, dx);
        inflow.updateMeshStatic(boxMesh(vmath::vec3(0.33f * L, 0.62f * L, 0.4f * L), 0.1f * L, 0.1f * L, 0.2f * L));
        inflow.setInflow();
        inflow.setVelocity(vmath::vec3(0.5f, -1.0f, 0.0f));
        drain.updateMeshStatic(boxMesh(vmath::vec3(0.86f * L, 0.05f * L, 0.1f * L), 0.08f * L, 0.15f * L, 0.8f * L));
        drain.setOutflow();
        sim.addMeshFluidSource(&inflow);
        sim.addMeshFluidSource(&drain);

        sim.setBoundaryFriction(0.2);       // the floor and the walls drag the flow (setBoundaryFriction / MeshObject::setFriction)
        block.setFriction(0.5f);
        sim.addBodyForce(0.0, -25.0, 0.0);
        sim.initialize();
        printf("initialized: %d^3 cells, 2 obstacles, 1 inflow, 1 outflow\n", n);
        int inRamp = 0, inBlock = 0, peak = 0;
        for (int f = 1; f <= frames; f++) {
            if (f == frames / 2) { block.disable(); printf("frame %4d  block disabled\n", f); }
            sim.update(1.0 / 30.0);
            const int np = sim.getNumMarkerParticles();
            peak = np > peak ? np : peak;
            if (f % 10 == 0 || f == frames) {
                std::vector<MarkerParticle> parts = sim.getMarkerParticles();
                inRamp = inBlock = 0;
                for (const auto &mp : parts) {
                    const vmath::vec3 p = mp.position;
                    // under the slope of the wedge by more than half a cell / inside the block by more than half a cell
                    const float rx = (p.x - 0.30f * L) / (0.4f * L), ry = (p.y - 0.04f * L) / (0.3f * L);
                    const bool inZ = p.z > 0.2f * L + 0.5f * dx && p.z < 0.8f * L - 0.5f * dx;
                    if (inZ && rx > 0.1f && ry > 0.1f && rx + ry < 0.8f) inRamp++;
                    if (p.x > 0.72f * L + 0.5f * dx && p.x < 0.84f * L - 0.5f * dx && p.y > 0.04f * L + 0.5f * dx && p.y < 0.24f * L - 0.5f * dx &&
                        p.z > 0.35f * L + 0.5f * dx && p.z < 0.65f * L - 0.5f * dx) inBlock++;
                }
                printf("frame %4d  particles %d  inside ramp %d  inside block %d\n", f, np, inRamp, inBlock);
            }
        }
        sim.removeMeshObstacle(&ramp);
        sim.update(1.0 / 30.0);
        printf("done: %d frames, peak %d particles, final %d\n", frames, peak, sim.getNumMarkerParticles());
        // an animated obstacle (MeshObject::updateMeshAnimated once per frame): a plate that sweeps along the floor
        MeshObject plate(n, n, n, dx);
        const vmath::vec3 p0(0.12f * L, 0.04f * L, 0.2f * L);
        const float stepx = 0.012f * L;
        auto plateAt = [&](int f) { return boxMesh(vmath::vec3(p0.x + stepx * f, p0.y, p0.z), 0.06f * L, 0.25f * L, 0.6f * L); };
        plate.updateMeshAnimated(plateAt(-1), plateAt(0), plateAt(1));
        sim.addMeshObstacle(&plate);
        // ... and a general animated mesh: a wedge that rises out of the floor (per-vertex velocities)
        MeshObject lift(n, n, n, dx);
        auto liftAt = [&](int f) { return wedgeMesh(vmath::vec3(0.55f * L, 0.02f * L + 0.006f * L * f, 0.25f * L), 0.2f * L, 0.12f * L, 0.5f * L); };
        lift.updateMeshAnimated(liftAt(-1), liftAt(0), liftAt(1));
        sim.addMeshObstacle(&lift);
        for (int f = 0; f < 12; f++) {
            plate.updateMeshAnimated(plateAt(f - 1), plateAt(f), plateAt(f + 1));
            lift.updateMeshAnimated(liftAt(f - 1), liftAt(f), liftAt(f + 1));
            sim.update(1.0 / 30.0);
        }
        int inPlate = 0;
        for (const auto &mp : sim.getMarkerParticles()) {
            const vmath::vec3 p = mp.position;
            if (p.x > p0.x + stepx * 12 + 0.5f * dx && p.x < p0.x + stepx * 12 + 0.06f * L - 0.5f * dx && p.y > p0.y + 0.5f * dx &&
                p.y < p0.y + 0.25f * L - 0.5f * dx && p.z > p0.z + 0.5f * dx && p.z < p0.z + 0.6f * L - 0.5f * dx) inPlate++;
        }
        printf("animated plate: 12 frames, particles %d, inside plate %d\n", sim.getNumMarkerParticles(), inPlate);
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return strstr(e.what(), "no CUDA device") ? 2 : 1;
    }
    return 0;
}
