"""Developer probe: multigrid parameter sweep on a scene (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flipengine3d_b200 import scenes, engine as fe

which = sys.argv[1] if len(sys.argv) > 1 else "sphere256"
sc = scenes.sphere_drop(int(which[6:])) if which.startswith("sphere") else scenes.dam_break(int(which[3:]))
I, J, K = sc["dims"]
sim = fe.FluidSimulation(I, J, K, sc["dx"])
sim.addBodyForce(0, -25, 0)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
sim.initialize()
for f in range(3):
    sim.update(1 / 30)
combos = [("jacobi",), (2, 0.9, 1.8, 8), (1, 0.9, 1.8, 8), (3, 0.9, 1.8, 8), (2, 0.9, 1.8, 4), (2, 1.0, 1.8, 8), (2, 0.9, 2.0, 8)]
for persistent in (False, True):
  sim.setSolverMode(persistent)
  print("persistent", persistent)
  for cb in combos:
    if cb[0] == "jacobi":
        sim.setPreconditioner("jacobi")
    else:
        sim.setPreconditioner("multigrid")
        sim.setMultigrid(*cb)
    sim.enable_kernel_timing(True); sim.reset_kernel_timing()
    sim.update(1 / 30)
    st = sim.substep_stats()
    tm = sim.stage_times_ms()
    kt = sim.kernel_timing()
    print(cb, "solve ms", round(kt["pcg_solve"][0], 3), "pcg", [s["pcg_iterations"] for s in st], "conv", [s["pcg_converged"] for s in st], "err/rhs", [f"{s['pcg_error']/max(s['rhs_max'],1e-300):.2e}" for s in st],
          "pressure ms", round(tm["pressure"], 3), "vcycle avg ms", round(kt["precond"][0] / max(kt["precond"][1], 1), 4),
          "iter avg ms", round(kt["pcg_iter"][0] / max(kt["pcg_iter"][1], 1), 4), flush=True)
