"""Developer probe: stage times of free-running frames on a scene (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flipengine3d_b200 import scenes, engine as fe

which = sys.argv[1] if len(sys.argv) > 1 else "dam128"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "jacobi"
t = time.time()
if which.startswith("dam"):
    sc = scenes.dam_break(int(which[3:]))
elif which.startswith("sphere"):
    sc = scenes.sphere_drop(int(which[6:]))
else:
    sc = scenes.default_scene(30)
print("scene", sc["name"], sc["pos"].shape, "gen", time.time() - t, flush=True)
I, J, K = sc["dims"]
t = time.time()
sim = fe.FluidSimulation(I, J, K, sc["dx"])
sim.addBodyForce(0, -25, 0)
sim.setPreconditioner(prec)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
sim.initialize()
print("init", time.time() - t, flush=True)
for f in range(frames):
    t = time.time()
    sim.update(1 / 30)
    sim.synchronize()
    el = time.time() - t
    st = sim.substep_stats()
    tm = sim.stage_times_ms()
    print(f"frame {f}: {el*1e3:.1f} ms, substeps {len(st)}, particles {st[-1]['particles']}, rows {st[-1]['pressure_rows']}, "
          f"pcg {[s['pcg_iterations'] for s in st]} conv {[s['pcg_converged'] for s in st]}")
    print("   last substep stage ms:", {k: round(v, 3) for k, v in tm.items() if v > 0.001}, flush=True)
