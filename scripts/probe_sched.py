"""Developer probe: damping schedules of the multigrid smoother on a scene (run on a GPU box).
Each combination runs one frame from the evolving state; reports PCG iterations and per-iteration times."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flipengine3d_b200 import scenes, engine as fe


def cheb(alpha, beta, k):
    mid, half = 0.5 * (beta + alpha), 0.5 * (beta - alpha)
    roots = [mid + half * math.cos(math.pi * (2 * q + 1) / (2 * k)) for q in range(k)]
    return [1.0 / r for r in roots]      # ascending damping: largest root first


which = sys.argv[1] if len(sys.argv) > 1 else "sphere256"
sc = scenes.sphere_drop(int(which[6:])) if which.startswith("sphere") else scenes.dam_break(int(which[3:]))
I, J, K = sc["dims"]
sim = fe.FluidSimulation(I, J, K, sc["dx"])
sim.addBodyForce(0, -25, 0)
sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
sim.initialize()
for f in range(3):
    sim.update(1 / 30)
combos = [([0.9, 0.9], 1.8)]
for alpha in (0.67, 0.5, 0.4, 0.33):
    for scale in (1.8, 1.5, 2.0):
        combos.append((cheb(alpha, 2.0, 2), scale))
combos += [(cheb(0.5, 1.9, 2), 1.8), (list(reversed(cheb(0.5, 2.0, 2))), 1.8), (cheb(0.4, 2.0, 3), 1.8), (cheb(0.3, 2.0, 3), 1.8),
           ([0.9], 1.8), (cheb(0.6, 2.0, 1), 1.8), ([0.9, 0.9], 1.8)]
for damp, scale in combos:
    sim.setMultigrid(len(damp), 0.9, scale, 8)
    sim.setMultigridSchedule(damp)
    sim.enable_kernel_timing(True); sim.reset_kernel_timing()
    sim.update(1 / 30)
    st = sim.substep_stats()
    tm = sim.stage_times_ms()
    kt = sim.kernel_timing()
    print([round(w, 4) for w in damp], scale, "pcg", [s["pcg_iterations"] for s in st], "conv", [s["pcg_converged"] for s in st],
          "pressure ms", round(tm["pressure"], 3), "vcycle ms", round(kt["precond"][0] / max(kt["precond"][1], 1), 4),
          "iter ms", round(kt["pcg_iter"][0] / max(kt["pcg_iter"][1], 1), 4), flush=True)
