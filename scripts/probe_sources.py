import sys, numpy as np
sys.path.insert(0, "tests")
import parity_common as pc
sys.path.insert(0, ".")
from flipengine3d_b200 import engine as fe
n, dx = 30, 0.125
inflow = ((12.3 * dx, 20.4 * dx, 12.3 * dx), (17.7 * dx, 23.6 * dx, 17.7 * dx))
outflow = ((3.2 * dx, 2.3 * dx, 3.2 * dx), (26.8 * dx, 4.7 * dx, 26.8 * dx))
vel = (0.0, -2.0, 0.0)
ref = pc.refengine.RefEngine((n, n, n), dx, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), threads=1)
rid = ref.add_fluid_source_box(*inflow, velocity=vel)
ref.add_fluid_source_box(*outflow, outflow=True)
gpu = fe.FluidSimulation(n, n, n, dx)
gpu.addBodyForce(0, -25, 0)
sid = gpu.addMeshFluidSourceBox(*inflow, velocity=vel)
gpu.addMeshFluidSourceBox(*outflow, outflow=True)
gpu.initialize()
for f in range(14):
    ref.update(1.0 / 30.0); gpu.update(1.0 / 30.0)
    a, b = ref.particles(), gpu.getMarkerParticles()
    print(f, a.shape[0], b.shape[0], ref.substeps, len(gpu.substep_stats()), "miny", a[:, 1].min() / dx, b[:, 1].min() / dx, "minvy", a[:, 4].min(), b[:, 4].min())
