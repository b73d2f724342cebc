"""Developer driver for compute-sanitizer, second half: the device paths outside the per-substep chain -- seeding of a queued
fluid object, an inflow source with the constrained velocity, an outflow source, and the surface reconstruction -- on a 24^3
grid through the C-ABI (no torch).  Usage: sanitize_objects.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flipengine3d_b200 import engine as fe

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, dx = 24, 0.125
sim = fe.FluidSimulation(n, n, n, dx)
sim.addBodyForce(0, -25, 0)
sim.addMeshFluidBox((3.3 * dx, 2.0 * dx, 3.3 * dx), (11.7 * dx, 9.6 * dx, 12.2 * dx))
sid = sim.addMeshFluidSourceBox((14.3 * dx, 15.4 * dx, 14.3 * dx), (18.7 * dx, 18.6 * dx, 18.7 * dx), velocity=(0.0, -2.0, 0.0))
sim.addMeshFluidSourceBox((2.2 * dx, 2.3 * dx, 16.2 * dx), (21.8 * dx, 4.7 * dx, 21.8 * dx), outflow=True)
sim.setSurfaceSubdivisionLevel(2)
sim.initialize()
for f in range(frames):
    sim.update(1 / 30)
    v, t = sim.getIsomesh()
    print("frame", f, "particles", sim.getNumMarkerParticles(), "vertices", v.shape[0], "triangles", t.shape[0])
sim.constrainMeshFluidSourceVelocity(sid, False)
sim.update(1 / 30)
sim.removeMeshFluidSource(sid)
sim.update(1 / 30)
sim.synchronize()
print("done", sim.getNumMarkerParticles())
