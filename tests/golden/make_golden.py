"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference engine
(oracle/_ref/libflipref_golden.so, built by oracle/Makefile from /root/reference).  Run in the build
container:  python tests/golden/make_golden.py

Fixtures (small, committed):
  default30_frames.npz   config 1 (FluidManager scene, 30^3, 8000 particles): per-frame particle
                         hashes/counts/PCG iterations over 6 frames + final particle state.
  default30_static.npz   the static inputs of that scene (solid SDF, face weights, near-solid mask).
  dam24_stages.npz       a 24^3 dam break, third frame, every intermediate array of one substep
                         (inputs and outputs of each stage), for per-stage known-answer tests.
The reference is bit-deterministic for injected particles with any thread count.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flipengine3d_b200 import scenes  # noqa: E402
from oracle.refengine import RefEngine  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STAGES = ("obstacles", "liquid_sdf", "p2g", "extrapolate_a", "save", "body_force", "pressure", "extrapolate_b",
          "constrain", "g2p", "advance", "tail")


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def default30():
    sc = scenes.default_scene(30)
    e = RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"], threads=4)
    hashes, counts, iters, cells, substeps = [], [], [], [], []
    for f in range(6):
        e.update(1.0 / 30.0)
        p = e.particles()
        hashes.append(md5(p))
        counts.append(e.num_particles)
        iters.append(e.pcg_iterations)
        cells.append(e.num_fluid_cells)
        substeps.append(e.substeps)
    np.savez_compressed(os.path.join(HERE, "default30_frames.npz"), hashes=np.array(hashes), counts=np.array(counts),
                        pcg_iterations=np.array(iters), fluid_cells=np.array(cells), substeps=np.array(substeps),
                        final_particles=e.particles())
    print("default30", hashes[-1], counts, iters, cells, substeps)


def default30_static():
    """Static inputs of the default scene (they come from the reference's mesh code, outside the hot path): nodal
    solid SDF, face weights, near-solid mask -- what oracle/restatement.py's whole-step engine takes as given."""
    sc = scenes.default_scene(30)
    e = RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"], threads=2)
    e.begin_frame(1.0 / 30.0)
    dt = e.begin_substep()
    e.stage("obstacles", dt)
    e.update_weight_grid()
    np.savez_compressed(os.path.join(HERE, "default30_static.npz"), solid_phi=e.array("solid_phi"), near_solid=e.array("near_solid"),
                        weightU=e.array("weightU"), weightV=e.array("weightV"), weightW=e.array("weightW"))
    e.close()
    print("default30_static written")


def dam24():
    sc = scenes.dam_break(24)
    e = RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"], threads=4)
    for f in range(2):
        e.update(1.0 / 30.0)
    out = {"dims": np.array(sc["dims"]), "dx": np.array(sc["dx"])}
    e.begin_frame(1.0 / 30.0)
    dt = e.begin_substep()
    out["dt"] = np.array(dt)
    out["particles_in"] = e.particles()
    e.stage("obstacles", dt)
    out["solid_phi"] = e.array("solid_phi")
    out["near_solid"] = e.array("near_solid")
    e.update_weight_grid()
    for n in ("weightU", "weightV", "weightW"):
        out[n] = e.array(n)
    grid_names = ("U", "V", "W", "validU", "validV", "validW", "liquid_phi")
    for st in STAGES[1:]:
        e.stage(st, dt)
        if st in ("liquid_sdf",):
            out["liquid_phi"] = e.array("liquid_phi")
        elif st in ("g2p", "advance"):
            out[f"{st}.particles"] = e.particles()
        elif st != "tail":
            for n in grid_names[:6]:
                out[f"{st}.{n}"] = e.array(n)
            if st in ("save", "constrain"):
                for n in ("savedU", "savedV", "savedW"):
                    out[f"{st}.{n}"] = e.array(n)
        if st == "pressure":
            out["pressure.iterations"] = np.array(e.pcg_iterations)
            out["pressure.error"] = np.array(e.pcg_error)
            out["pressure.fluid_cells"] = np.array(e.num_fluid_cells)
    np.savez_compressed(os.path.join(HERE, "dam24_stages.npz"), **out)
    print("dam24", {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]}, "iters", out["pressure.iterations"])


if __name__ == "__main__":
    default30()
    default30_static()
    dam24()
