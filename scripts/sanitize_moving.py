"""Developer driver for compute-sanitizer, third part: the moving-solid paths -- an animated box obstacle (per-substep solid SDF,
solid fractions, normalisation + extrapolation of the solid velocities), the enclosed-pocket conditioning, the solid terms of
the pressure system and the moving constraint, then prescribed solid velocities over a closed chamber -- on a 24^3 grid
through the C-ABI (no torch, no oracle).  Usage: sanitize_moving.py [frames]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flipengine3d_b200 import engine as fe, scenes

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = scenes.dam_break(24)
n, dx = 24, sc["dx"]
sim = fe.FluidSimulation(n, n, n, dx)
sim.addBodyForce(0, -25, 0)
lo, hi = (11.3 * dx, 1.8 * dx, 4.2 * dx), (14.4 * dx, 12.7 * dx, 19.6 * dx)
oid = sim.addMeshObstacleBox(lo, hi)
# a closed chamber of six walls, brim-full (an enclosed pocket)
o0, o1, t = np.array([15.3, 12.3, 6.3]), np.array([22.7, 21.7, 17.7]), 2.4
for a in range(3):
    l, h = o0.copy(), o1.copy(); h[a] = o0[a] + t; sim.addMeshObstacleBox(tuple(l * dx), tuple(h * dx))
    l, h = o0.copy(), o1.copy(); l[a] = o1[a] - t; sim.addMeshObstacleBox(tuple(l * dx), tuple(h * dx))
sim.initialize()
cells = scenes.box_cells(18, 20, 15, 19, 9, 15)
pos, vel = scenes.seed_cells(cells, dx, 5)
P = np.concatenate([np.concatenate([sc["pos"], sc["vel"]], axis=1), np.concatenate([pos, vel], axis=1)], axis=0).astype(np.float32)
sim.setMarkerParticles(P)
step = np.array([-0.04, 0.0, 0.0])
for f in range(frames):
    sim.setMeshObstacleBoxMotion(oid, step * (f - 1), step * f, step * (f + 1))
    sim.update(1 / 30)
    st = sim.substep_stats()[-1]
    print("frame", f, "particles", sim.getNumMarkerParticles(), "rows", st["pressure_rows"], "its", st["pcg_iterations"],
          "solid |u| max", float(np.abs(sim.array("solidU")).max()), "zero faces", int(np.count_nonzero(sim.array("solidU") == 0)))
sim.enableMeshObstacle(oid, False)
sim.update(1 / 30)
shapes = {k: sim.shape_of("solid" + k) for k in "UVW"}
sim.setSolidVelocity(*[np.full(shapes[k], v, dtype=np.float32) for k, v in zip("UVW", (0.4, -0.2, 0.3))])
sim.update(1 / 30)
print("prescribed: zero faces", int(np.count_nonzero(sim.array("solidU") == 0)), "its", sim.substep_stats()[-1]["pcg_iterations"])
sim.setSolidVelocity()
sim.update(1 / 30)
sim.synchronize()
print("done", sim.getNumMarkerParticles())
