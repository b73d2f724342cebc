"""Where the end-to-end frame goes: host->device (pinned AoS), update, device->host, each timed alone on spheredrop256."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flipengine3d_b200 import scenes, engine as fe
sc = scenes.sphere_drop(int(sys.argv[1]) if len(sys.argv) > 1 else 256)
I, J, K = sc["dims"]
sim = fe.FluidSimulation(I, J, K, sc["dx"]); sim.addBodyForce(0, -25, 0)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"])); sim.initialize()
for _ in range(3): sim.update(1 / 30)
n = sim.getNumMarkerParticles()
host = torch.empty((n + 65536, 6), dtype=torch.float32).pin_memory().numpy()
sim.getMarkerParticles(out=host); sim.synchronize()
# the bench.py e2e loop verbatim, per-iteration wall time
per = []
for _ in range(10):
    t = time.perf_counter()
    nn = sim.getNumMarkerParticles()
    sim.setMarkerParticles(host[:nn])
    sim.update(1 / 30)
    st = sim.substep_stats()
    nn = sim.getNumMarkerParticles()
    sim.getMarkerParticles(out=host)
    per.append(round(1e3 * (time.perf_counter() - t), 2))
sim.synchronize()
print("bench-style e2e iterations (ms):", per, "substeps", len(st), "pcg", [q["pcg_iterations"] for q in st])
n = sim.getNumMarkerParticles()
acc = {"set": 0.0, "update": 0.0, "get": 0.0}
R = 5
for _ in range(R):
    t = time.perf_counter(); sim.setMarkerParticles(host[:n]); sim.synchronize(); acc["set"] += time.perf_counter() - t
    t = time.perf_counter(); sim.update(1 / 30); sim.synchronize(); acc["update"] += time.perf_counter() - t
    n = sim.getNumMarkerParticles()
    t = time.perf_counter(); sim.getMarkerParticles(out=host); sim.synchronize(); acc["get"] += time.perf_counter() - t
mb = n * 24 / 1e6
print({k: round(1e3 * v / R, 3) for k, v in acc.items()}, "MB each way", round(mb, 1), "set GB/s", round(mb / (acc["set"] / R) / 1e3, 1), "get GB/s", round(mb / (acc["get"] / R) / 1e3, 1))
# raw copies for comparison
d = torch.empty((n, 6), dtype=torch.float32, device="cuda")
h = torch.from_numpy(host[:n])
torch.cuda.synchronize()
for name, fn in (("raw H2D", lambda: d.copy_(h, non_blocking=True)), ("raw D2H", lambda: h.copy_(d, non_blocking=True))):
    t = time.perf_counter()
    for _ in range(R): fn()
    torch.cuda.synchronize()
    el = (time.perf_counter() - t) / R
    print(name, round(1e3 * el, 3), "ms", round(mb / el / 1e3, 1), "GB/s")
