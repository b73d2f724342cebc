"""Developer driver for compute-sanitizer: a few frames of a small scene through the C-ABI (no torch).
Usage: sanitize_step.py [dam32|default|sphere32] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flipengine3d_b200 import scenes, engine as fe

which = sys.argv[1] if len(sys.argv) > 1 else "dam32"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = scenes.dam_break(int(which[3:])) if which.startswith("dam") else (scenes.sphere_drop(int(which[6:])) if which.startswith("sphere") else scenes.default_scene(30))
I, J, K = sc["dims"]
sim = fe.FluidSimulation(I, J, K, sc["dx"])
sim.addBodyForce(0, -25, 0)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
sim.initialize()
for f in range(frames):
    sim.update(1 / 30)
sim.synchronize()
st = sim.substep_stats()
print("frames", frames, "particles", st[-1]["particles"], "rows", st[-1]["pressure_rows"], "pcg", [s["pcg_iterations"] for s in st])
