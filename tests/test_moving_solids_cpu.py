"""Moving solids on the CPU: the host-side centre weights of the library and the restated enclosed-pocket conditioning and
solid divergence terms (oracle/restatement.py), pinned to the unmodified reference engine (needs oracle/_ref)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import parity_common as pc
from flipengine3d_b200 import engine as fe
from flipengine3d_b200 import scenes
from oracle import refengine
from oracle import restatement as rs

needs_ref = pytest.mark.skipif(not refengine.available("golden"), reason="oracle/_ref/libflipref_golden.so not built")
_WALL = ((12.3 * 0.125, 0.0, 6.2 * 0.125), (16.7 * 0.125, 9.4 * 0.125, 25.9 * 0.125))


@needs_ref
def test_center_weights_equal_the_reference():
    """flip_center_weights (host code of csrc/static_host.cpp) against _updateWeightGridThread's CENTER branch
    (fluidsimulation.cpp:3720-3727): box walls and obstacles, and a tilted, noisy ellipsoid whose cells show every
    sign pattern of the eight corners (tetrahedron and prism cases of LevelsetUtils::volumeFraction) -- bit for bit."""
    sc = scenes.dam_break_with_chamber(32)
    ref = refengine.RefEngine(sc["dims"], sc["dx"], sc["pos"][:8], sc["vel"][:8])
    for lo, hi in [_WALL] + list(sc["obstacles"]):
        ref.add_obstacle_box(lo, hi)
    ref.stage("obstacles", 1.0 / 30.0)
    ref.update_weight_grid()
    phi, wc = ref.array("solid_phi"), ref.array("weightC")
    assert np.count_nonzero((wc > 0) & (wc < 1)) > 3000
    assert np.array_equal(fe.center_weights(sc["dims"], sc["dx"], phi), wc)
    I, J, K = sc["dims"]
    dx = sc["dx"]
    z, y, x = np.meshgrid(np.arange(K + 1) * dx, np.arange(J + 1) * dx, np.arange(I + 1) * dx, indexing="ij")
    e = np.sqrt(((x - 2.0 + 0.3 * (y - 2)) / 0.9) ** 2 + ((y - 1.6) / 0.55) ** 2 + ((z - 2.1 - 0.2 * (x - 2)) / 0.7) ** 2) - 1.0
    e = (e * 0.5 + 0.02 * np.random.default_rng(3).standard_normal(e.shape)).astype(np.float32)
    phi2 = np.minimum(phi, e)
    ref.set_array("solid_phi", phi2)
    ref.update_weight_grid(force=True)
    wc2 = ref.array("weightC")
    assert np.count_nonzero(wc2 != wc) > 1000
    assert np.array_equal(fe.center_weights(sc["dims"], dx, phi2), wc2)
    ref.close()


@needs_ref
def test_pocket_conditioning_and_solid_terms_equal_the_reference():
    """Three frames of the dam break with the brim-full chamber and a smooth solid velocity field in the reference; in the
    next substep the restated conditioning (spread of "reaches air" instead of the reference's flood fill) zeroes exactly
    the faces the reference zeroes, and the restated projection with the solid terms and the reference's own MIC(0)-PCG
    reproduces iteration count, residual and the projected field bit for bit -- and does not without the terms."""
    sc = scenes.dam_break_with_chamber(32)
    ref = refengine.RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"])
    for lo, hi in [_WALL] + list(sc["obstacles"]):
        ref.add_obstacle_box(lo, hi)
    ref.stage("obstacles", 1.0 / 30.0)
    vel = pc.solid_velocity_field({n: ref.shape_of("solid" + n) for n in "UVW"})
    for n in "UVW":
        ref.set_array("solid" + n, vel[n])
    for _ in range(3):
        ref.update(1.0 / 30.0)
    zeroed = {n: ref.array("solid" + n) != vel[n] for n in "UVW"}
    assert all(z.sum() >= 300 for z in zeroed.values())
    # put the velocities back: the conditioning of the next substep has to find the chamber again
    for n in "UVW":
        ref.set_array("solid" + n, vel[n])
    ref.begin_frame(1.0 / 30.0)
    dt = ref.begin_substep()
    for st in ("obstacles", "liquid_sdf", "p2g", "extrapolate_a", "save", "body_force"):
        ref.stage(st, dt)
    before = {n: ref.array(n) for n in "UVW"}
    ref.update_weight_grid()
    phi = ref.array("liquid_phi")
    w = {n: ref.array("weight" + n) for n in "UVWC"}
    ref.stage("pressure", dt)
    sU, sV, sW, pocket = rs.condition_solid_velocities(phi, w["U"], w["V"], w["W"], vel["U"], vel["V"], vel["W"])
    assert pocket.sum() >= 100
    for mine, n in zip((sU, sV, sW), "UVW"):
        assert np.array_equal(mine, ref.array("solid" + n)), n
        assert np.array_equal(mine != vel[n], zeroed[n]), n
    info = {}
    out = rs.pressure_project(before["U"], before["V"], before["W"], phi, w["U"], w["V"], w["W"], sc["dims"], sc["dx"], dt,
                              solver="mic0_pcg", info=info, solid=(sU, sV, sW, w["C"]))
    assert info["iterations"] == ref.pcg_iterations and info["converged"]
    for a, n in zip(out[:3], "UVW"):
        assert np.array_equal(a, ref.array(n)), n
    plain = rs.pressure_project(before["U"], before["V"], before["W"], phi, w["U"], w["V"], w["W"], sc["dims"], sc["dx"], dt,
                                solver="mic0_pcg")
    assert max(np.abs(a - ref.array(n)).max() for a, n in zip(plain[:3], "UVW")) > 0.1
    ref.close()
