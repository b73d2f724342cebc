"""Developer probe (GPU + oracle/_ref): the animated box obstacle of tests/test_moving_solids_gpu.py frame by frame against the
reference -- substeps, counts, pressure rows, position / centre-of-mass / mean-velocity differences, and the solid SDF,
weights and solid face velocities after every frame.  Prints; asserts nothing."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc
from flipengine3d_b200 import scenes
import test_moving_solids_gpu as T

sc = scenes.dam_break(32)
dx = sc["dx"]
lo, hi = T._PADDLE
ref, gpu = pc.make_pair(sc, obstacles=[T._PADDLE], own_solid=True)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for f in range(frames):
    offs = [T._PADDLE_STEP * (f - 1), T._PADDLE_STEP * f, T._PADDLE_STEP * (f + 1)]
    ref.animate_obstacle_box(0, lo, hi, *offs)
    gpu.setMeshObstacleBoxMotion(1, *offs)
    ref.update(1.0 / 30.0)
    gpu.update(1.0 / 30.0)
    st = gpu.substep_stats()
    p, ids = pc.particles_by_id(gpu)
    a = ref.particles()
    line = dict(frame=f, substeps=(ref.substeps, len(st)), particles=(ref.num_particles, st[-1]["particles"]),
                rows=(ref.num_fluid_cells, st[-1]["pressure_rows"]), its=(ref.pcg_iterations, st[-1]["pcg_iterations"]),
                conv=[s["pcg_converged"] for s in st])
    if p.shape[0] == a.shape[0]:
        line["pos_rel_l2"] = pc.rel_l2(p[np.argsort(ids), :3], a[:, :3])
    line["com_diff"] = float(np.abs(p[:, :3].mean(0) - a[:, :3].mean(0)).max())
    line["meanvel_diff"] = float(np.abs(p[:, 3:].mean(0) - a[:, 3:].mean(0)).max())
    line["vmax"] = (float(np.abs(a[:, 3:]).max()), float(np.abs(p[:, 3:]).max()))
    R, G = ref.array("solid_phi"), gpu.array("solid_phi")
    near = np.abs(R) < 2.5 * dx
    line["phi_sign_diff"] = int(np.count_nonzero((R < 0) != (G < 0)))
    line["phi_near_maxdiff_dx"] = float(np.abs(R[near] - G[near]).max() / dx)
    line["weights_maxdiff"] = [float(np.abs(ref.array(n) - gpu.array(n)).max()) for n in ("weightU", "weightV", "weightW", "weightC")]
    line["solidvel_maxdiff"] = [float(np.abs(ref.array("solid" + n) - gpu.array("solid" + n)).max()) for n in "UVW"]
    line["solidvel_max"] = [float(np.abs(ref.array("solid" + n)).max()) for n in "UVW"]
    print(line, flush=True)
