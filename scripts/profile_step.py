"""Profiling driver: W warm-up frames, then P frames between cudaProfilerStart/Stop (run under
`ncu --profile-from-start off ...` on a GPU box).  Usage: profile_step.py [scene] [warm] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipengine3d_b200 import scenes, engine as fe

which = sys.argv[1] if len(sys.argv) > 1 else "sphere256"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1
prec = sys.argv[4] if len(sys.argv) > 4 else None
if which.startswith("dam"):
    sc = scenes.dam_break(int(which[3:]))
elif which.startswith("sphere"):
    sc = scenes.sphere_drop(int(which[6:]))
else:
    sc = scenes.default_scene(30)
I, J, K = sc["dims"]
sim = fe.FluidSimulation(I, J, K, sc["dx"])
sim.addBodyForce(0, -25, 0)
if prec:
    sim.setPreconditioner(prec)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
sim.initialize()
for f in range(warm):
    sim.update(1 / 30)
sim.synchronize()
torch.cuda.profiler.start()
for f in range(frames):
    sim.update(1 / 30)
sim.synchronize()
torch.cuda.profiler.stop()
st = sim.substep_stats()
print("particles", st[-1]["particles"], "rows", st[-1]["pressure_rows"], "pcg", [s["pcg_iterations"] for s in st])
