"""The oracle (the unmodified reference engine built by oracle/Makefile into oracle/_ref/) against the committed
golden fixtures: the fixtures are reproducible from the reference, with a different thread count than the one
that generated them (bit-deterministic for injected particles).  CPU only; skipped when
oracle/_ref has not been built (it always is by __graft_entry__.build() where /root/reference exists, and it
travels to the GPU box)."""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flipengine3d_b200 import scenes  # noqa: E402
from oracle import refengine  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.skipif(not refengine.available("golden"), reason="oracle/_ref/libflipref_golden.so not built")


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_default_scene_reproduces_the_golden_frames():
    g = np.load(os.path.join(GOLDEN, "default30_frames.npz"))
    sc = scenes.default_scene(30)
    e = refengine.RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"], threads=2)
    for f in range(6):
        e.update(1.0 / 30.0)
        assert md5(e.particles()) == str(g["hashes"][f]), f
        assert e.num_particles == int(g["counts"][f])
        assert e.pcg_iterations == int(g["pcg_iterations"][f])
        assert e.num_fluid_cells == int(g["fluid_cells"][f])
        assert e.substeps == int(g["substeps"][f])
    assert np.array_equal(e.particles(), g["final_particles"])
    e.close()


def test_dam_break_stage_arrays_reproduce_the_golden_substep():
    g = np.load(os.path.join(GOLDEN, "dam24_stages.npz"))
    sc = scenes.dam_break(24)
    assert tuple(g["dims"]) == tuple(sc["dims"]) and float(g["dx"]) == sc["dx"]
    e = refengine.RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"], threads=3)
    for f in range(2):
        e.update(1.0 / 30.0)
    e.begin_frame(1.0 / 30.0)
    dt = e.begin_substep()
    assert dt == float(g["dt"])
    assert np.array_equal(e.particles(), g["particles_in"])
    e.stage("obstacles", dt)
    assert np.array_equal(e.array("solid_phi"), g["solid_phi"])
    e.update_weight_grid()
    for n in ("weightU", "weightV", "weightW"):
        assert np.array_equal(e.array(n), g[n]), n
    for st in ("liquid_sdf", "p2g", "extrapolate_a", "save", "body_force", "pressure", "extrapolate_b", "constrain", "g2p", "advance"):
        e.stage(st, dt)
        if st == "liquid_sdf":
            assert np.array_equal(e.array("liquid_phi"), g["liquid_phi"])
        elif st in ("g2p", "advance"):
            assert np.array_equal(e.particles(), g[f"{st}.particles"]), st
        else:
            for n in ("U", "V", "W", "validU", "validV", "validW"):
                assert np.array_equal(e.array(n), g[f"{st}.{n}"]), (st, n)
        if st == "pressure":
            assert e.pcg_iterations == int(g["pressure.iterations"])
            assert e.num_fluid_cells == int(g["pressure.fluid_cells"])
    e.close()
