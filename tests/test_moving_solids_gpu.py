"""Moving solids, the hot-path half of SURVEY §8f rank 4 (-m gpu, through the C-ABI): solid face velocities in the
right-hand side of the pressure system (pressuresolver.cpp:595-613), the enclosed-pocket conditioning
(_conditionSolidVelocityField :124-244) and the solid constraint (fluidsimulation.cpp:3884-3933), against the unmodified
reference engine whose solid SDF carries the same face velocities (VelocityDataGrid, meshlevelset.h:69-87)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import parity_common as pc
from flipengine3d_b200 import engine as fe
from flipengine3d_b200 import scenes

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not pc.refengine.available("golden"), reason="oracle/_ref/libflipref_golden.so not built")

_WALL = ((12.3 * 0.125, 0.0, 6.2 * 0.125), (16.7 * 0.125, 9.4 * 0.125, 25.9 * 0.125))       # across the flow


def _scene():
    sc = scenes.dam_break_with_chamber(32)
    return sc, [_WALL] + list(sc["obstacles"])


@needs_ref
def test_lockstep_with_moving_solids():
    """Every stage from identical inputs over six frames of a dam break that runs into a wall, with a brim-full closed
    chamber elsewhere in the domain: the conditioned solid velocities (the chamber's faces zeroed, everything else kept)
    and the centre weights bit for bit, the projected field within the pressure tolerance, the constrained field bit for
    bit (zero-weight faces carry the solid's velocity)."""
    sc, obstacles = _scene()
    reps = pc.lockstep_frames(sc, frames=6, isolate=True, obstacles=obstacles, solid_velocity=True)
    assert len(reps) >= 6
    for rep in reps:
        pc.check_report(rep, dx=sc["dx"], isolate=True)
        # the chamber is found (some hundred faces per component zeroed) and the solid terms act on the flow
        assert rep["solid.U.zero_faces"] >= 300 and rep["solid.V.zero_faces"] >= 300 and rep["solid.W.zero_faces"] >= 300, rep
    # once the liquid has reached the floor and the wall the solid terms dominate the projected field
    assert max(r["pressure.scale"] for r in reps) > 1.0, [r["pressure.scale"] for r in reps]


@needs_ref
def test_moving_solids_free_running_against_reference():
    """flip_update against FluidSimulation::update with the same solid velocities (handed over BEFORE flip_initialize
    here): same substeps, particle counts, pressure rows; positions to 1e-4 rel-L2 over the first frames.  Then the
    velocities are withdrawn and the run goes on as a static scene."""
    sc, obstacles = _scene()
    ref, gpu = pc.make_pair(sc, obstacles=obstacles, solid_velocity=True, solid_velocity_early=True)
    for f in range(5):
        ref.update(1.0 / 30.0)
        gpu.update(1.0 / 30.0)
        st = gpu.substep_stats()
        assert ref.substeps == len(st) and ref.num_particles == st[-1]["particles"], (f, ref.substeps, len(st), ref.num_particles, st[-1]["particles"])
        assert abs(ref.num_fluid_cells - st[-1]["pressure_rows"]) <= 2, (f, ref.num_fluid_cells, st[-1]["pressure_rows"])
        assert all(s["pcg_converged"] == 1 for s in st), st
        if f < 3:
            p, ids = pc.particles_by_id(gpu)
            a = ref.particles()
            assert pc.rel_l2(p[np.argsort(ids), :3], a[:, :3]) <= 1e-4, (f, pc.rel_l2(p[np.argsort(ids), :3], a[:, :3]))
    for name in "UVW":      # the conditioning acted in place in both engines
        assert np.array_equal(gpu.array("solid" + name), ref.array("solid" + name)), name
    gpu.setSolidVelocity()
    gpu.update(1.0 / 30.0)
    assert gpu.substep_stats()[-1]["pcg_converged"] == 1
    ref.close()
    gpu.close()


def test_solid_velocity_api_errors():
    sim = fe.FluidSimulation(12, 12, 12, 0.125)
    sim.initialize()
    U = np.zeros(sim.shape_of("solidU"), dtype=np.float32)
    rc = sim.L.flip_set_solid_velocity(sim.h, U.ctypes.data, None, None)           # three arrays or none
    assert rc == fe.FLIP_ERR_RUNTIME
    with pytest.raises(Exception):
        sim.array("weightC")                                                       # not built while every solid is at rest
    sim.setSolidVelocity(U, np.zeros(sim.shape_of("solidV"), np.float32), np.zeros(sim.shape_of("solidW"), np.float32))
    assert sim.array("weightC").shape == sim.shape_of("weightC")
    sim.update(1.0 / 30.0)
    sim.setSolidVelocity()
    sim.update(1.0 / 30.0)
    sim.close()
