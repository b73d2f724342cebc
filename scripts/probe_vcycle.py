"""Developer probe: frame / pressure / V-cycle time on spheredrop256 (10 frames after 3 warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipengine3d_b200 import scenes, engine as fe
sc = scenes.sphere_drop(256)
sim = fe.FluidSimulation(256, 256, 256, sc["dx"]); sim.addBodyForce(0, -25, 0)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"])); sim.initialize()
for _ in range(3): sim.update(1 / 30)
sim.synchronize()
sim.enable_kernel_timing(True); sim.reset_kernel_timing()
st = torch.cuda.ExternalStream(sim.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
its = 0; pres = 0.0
for _ in range(10):
    sim.update(1 / 30)
    its += sum(s["pcg_iterations"] for s in sim.substep_stats()); pres += sim.stage_times_ms()["pressure"]
e1.record(st); sim.synchronize(); torch.cuda.synchronize()
kt = sim.kernel_timing()
print(os.environ.get("FLIP_MG_COARSE_BLOCKS", "default"), "ms/frame", round(e0.elapsed_time(e1) / 10, 3), "pressure", round(pres / 10, 3), "its", its,
      "vcycle", round(kt["precond"][0] / kt["precond"][1], 4), "pcg_iter", round(kt["pcg_iter"][0] / kt["pcg_iter"][1], 4))
