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


@needs_ref
def test_mesh_velocity_data_equals_the_reference():
    """flip_mesh_velocity_data (host code) against the reference's solid velocity field for an obstacle animated with
    updateMeshAnimated -- a wedge that turns and translates, so every vertex has its own velocity: the library's per-face
    solid fractions and fraction x nearest-surface velocity, summed with the domain's fractions, normalised and extrapolated
    over 5 layers (restated), equal the field the reference's merged MeshLevelSet holds after _updateSolidLevelSet to float
    rounding; the mesh's signed distance field likewise."""
    sc = scenes.dam_break(32)
    dx = sc["dx"]
    I, J, K = sc["dims"]
    tris = scenes.WEDGE_TRIANGLES

    def frame(f):
        return scenes.wedge_vertices((2.3 - 0.04 * f, 1.1 + 0.01 * f, 2.0), 0.05 * f)
    ref = refengine.RefEngine(sc["dims"], dx, sc["pos"], sc["vel"])
    idx = ref.add_obstacle_mesh(frame(0), tris)
    f = 2
    ref.animate_obstacle_mesh(idx, frame(f - 1), frame(f), frame(f + 1), tris)
    ref.begin_frame(1.0 / 30.0)
    dt = ref.begin_substep()
    ref.stage("obstacles", dt)                  # frame progress 0: the mesh stands at frame(f), velocities (cur - prev) / dt
    vel = ((frame(f) - frame(f - 1)).astype(np.float64) * 30.0).astype(np.float32)
    md = fe.mesh_velocity_data(sc["dims"], dx, frame(f), tris, vel, band=3, far=3.0e38)
    dom = fe.static_inputs(I, J, K, dx)
    R, M = ref.array("solid_phi"), np.minimum(dom["solid_phi"], md["phi"])
    assert np.array_equal(R < 0, M < 0)
    near = np.abs(R) < 2.5 * dx
    assert np.abs(R[near] - M[near]).max() <= 1e-4 * dx
    seen = 0.0
    for n in "UVW":
        wsum = (1.0 - dom["weight" + n]) + md["fraction" + n]
        ok = wsum > 1e-6
        u = np.where(ok, md["field" + n] / np.where(ok, wsum, 1.0), 0.0).astype(np.float32)
        u = rs.extrapolate(u, ok.astype(np.uint8), layers=5)
        u = u[0] if isinstance(u, tuple) else u
        want = ref.array("solid" + n)
        seen = max(seen, float(np.abs(want).max()))
        assert np.abs(u - want).max() <= 5e-6, (n, float(np.abs(u - want).max()))         # measured: 2.4e-7
    assert seen > 1.5           # the turning wedge's surface moves at up to 1.9 units/s
    ref.close()


_FRICTION_BOXES = [((12.3 * 0.125, 0.0, 6.2 * 0.125), (16.7 * 0.125, 9.4 * 0.125, 25.9 * 0.125)),                    # a wall across the flow
                   ((20.27 * 0.125, 5.52 * 0.125, 12.13 * 0.125), (24.91 * 0.125, 11.64 * 0.125, 19.86 * 0.125)),    # a block above the floor
                   ((14.93 * 0.125, 6.31 * 0.125, 9.44 * 0.125), (19.58 * 0.125, 12.17 * 0.125, 15.62 * 0.125)),     # ... one that overlaps the wall
                   ((22.41 * 0.125, 0.0, 3.37 * 0.125), (29.83 * 0.125, 4.56 * 0.125, 9.12 * 0.125))]                # a step into the domain walls
_FRICTIONS = (0.35, [0.7, 0.2, 0.45, 0.9])


@needs_ref
def test_face_friction_equals_the_reference_where_the_constraint_reads_it():
    """flip_face_friction (host code; the merge order and take-over rule of MeshLevelSet::calculateUnion restated) against
    _getFaceFrictionU/V/W of the reference for a domain with boundary friction and four boxes of different friction, two
    of them overlapping and one reaching into the domain walls: equal on every partly open face (the only faces whose
    friction the constraint reads, fluidsimulation.cpp:3895)."""
    sc = scenes.dam_break(32)
    dx = sc["dx"]
    I, J, K = sc["dims"]
    ref = refengine.RefEngine(sc["dims"], dx, sc["pos"][:8], sc["vel"][:8])
    ref.set_boundary_friction(_FRICTIONS[0])
    for (lo, hi), f in zip(_FRICTION_BOXES, _FRICTIONS[1]):
        ref.set_obstacle_friction(ref.add_obstacle_box(lo, hi), f)
    ref.stage("obstacles", 1.0 / 30.0)
    ref.update_weight_grid()
    want = ref.face_friction()
    phis = [fe.static_inputs(I, J, K, dx)["solid_phi"]] + [fe.box_obstacle_sdf(sc["dims"], dx, lo, hi) for lo, hi in _FRICTION_BOXES]
    mine = fe.face_friction(sc["dims"], dx, phis, [_FRICTIONS[0]] + list(_FRICTIONS[1]))
    for n in "UVW":
        w = ref.array("weight" + n)
        partial = (w > 0) & (w < 1)
        assert partial.sum() > 3000
        assert len(np.unique(want[n][partial])) >= 10           # blends of the five frictions around edges and overlaps
        assert np.array_equal(mine[n][partial], want[n][partial]), (n, int(np.count_nonzero(mine[n][partial] != want[n][partial])))
    ref.close()


_BOX_TRIANGLES = np.array([[0, 1, 2], [0, 2, 3], [4, 7, 6], [4, 6, 5], [0, 3, 7], [0, 7, 4], [1, 5, 6], [1, 6, 2], [0, 4, 5], [0, 5, 1],
                           [3, 2, 6], [3, 6, 7]], dtype=np.int32)


def _box_vertices(lo, hi):
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    return np.array([[x0, y0, z0], [x1, y0, z0], [x1, y0, z1], [x0, y0, z1], [x0, y1, z0], [x1, y1, z0], [x1, y1, z1], [x0, y1, z1]],
                    dtype=np.float32)


@needs_ref
def test_box_obstacle_field_equals_the_reference_level_set_of_the_box_mesh():
    """flip_box_obstacle_sdf (the analytic box distances the library gives a box obstacle, and every substep an animated
    one) against MeshLevelSet::fastCalculateSignedDistanceField of the twelve-triangle box mesh: the same sign wherever both
    computed a distance, the same value within 2e-4 dx near the surface -- for a box inside the domain and one cut by it."""
    dims, dx = (32, 32, 32), 0.125
    for lo, hi in (((1.56, 0.22, 0.81), (1.96, 1.63, 3.17)), ((2.81, -0.4, 1.13), (4.3, 0.97, 2.26))):
        mine = fe.box_obstacle_sdf(dims, dx, lo, hi, band=3)
        want = refengine.mesh_level_set(dims, dx, _box_vertices(lo, hi), _BOX_TRIANGLES, band=3)
        both = (mine < 1e30) & (np.abs(want) < 3.0 * dx)
        assert both.sum() > 1500
        assert np.array_equal(mine[both] < 0, want[both] < 0)
        near = both & (np.abs(want) < 2.5 * dx)
        assert np.abs(mine[near] - want[near]).max() <= 2e-4 * dx, np.abs(mine[near] - want[near]).max() / dx


def test_mesh_velocity_data_of_a_mesh_at_rest_and_bad_input():
    """Host code, no oracle: a mesh whose vertices do not move adds solid fractions and no velocity (the reference's
    isStatic branch, meshlevelset.cpp:1410-1417); its fractions are 1 - the face weights flip_static_inputs derives from the
    same field; a triangle that names a vertex out of range is refused."""
    dims, dx = (24, 24, 24), 0.125
    v = scenes.wedge_vertices((1.5, 1.4, 1.5), 0.3)
    md = fe.mesh_velocity_data(dims, dx, v, scenes.WEDGE_TRIANGLES, np.zeros_like(v), band=3, far=3.0e38)
    for n in "UVW":
        assert not md["field" + n].any()
        assert md["fraction" + n].max() == 1.0 and md["fraction" + n].min() == 0.0
    w = fe.static_inputs(*dims, dx, solid_phi=md["phi"])
    for n in "UVW":
        assert np.abs((1.0 - w["weight" + n]) - md["fraction" + n]).max() <= 1e-6, n
    bad = scenes.WEDGE_TRIANGLES.copy()
    bad[3, 1] = 17
    with pytest.raises(IndexError):
        fe.mesh_velocity_data(dims, dx, v, bad, np.zeros_like(v))
    # a uniform translation: every face with a positive fraction carries fraction x that velocity
    vel = np.tile(np.array([[0.4, -0.7, 0.25]], dtype=np.float32), (v.shape[0], 1))
    md = fe.mesh_velocity_data(dims, dx, v, scenes.WEDGE_TRIANGLES, vel, band=3, far=3.0e38)
    for c, n in enumerate("UVW"):
        fr, fl = md["fraction" + n], md["field" + n]
        assert np.abs(fl - fr * vel[0, c]).max() <= 2e-6, n


@needs_ref
def test_restated_constraint_with_solid_velocities_and_friction_equals_the_reference():
    """oracle/restatement.constrain with solid velocities and friction against the reference's _constrainVelocityFields on a
    developed flow (boundary friction, four boxes of different friction, a smooth solid velocity field): bit for bit."""
    sc = scenes.dam_break(32)
    ref = refengine.RefEngine(sc["dims"], sc["dx"], sc["pos"], sc["vel"])
    ref.set_boundary_friction(_FRICTIONS[0])
    for (lo, hi), f in zip(_FRICTION_BOXES, _FRICTIONS[1]):
        ref.set_obstacle_friction(ref.add_obstacle_box(lo, hi), f)
    ref.stage("obstacles", 1.0 / 30.0)
    vel = pc.solid_velocity_field({n: ref.shape_of("solid" + n) for n in "UVW"})
    for n in "UVW":
        ref.set_array("solid" + n, vel[n])
    for _ in range(4):
        ref.update(1.0 / 30.0)
    ref.begin_frame(1.0 / 30.0)
    dt = ref.begin_substep()
    for st in ("obstacles", "liquid_sdf", "p2g", "extrapolate_a", "save", "body_force", "pressure", "extrapolate_b"):
        ref.stage(st, dt)
    before = {n: ref.array(n) for n in "UVW"}
    saved = {n: ref.array("saved" + n) for n in "UVW"}
    solid = {n: ref.array("solid" + n) for n in "UVW"}          # as the pocket conditioning of this substep left them
    fric = ref.face_friction()
    ref.stage("constrain", dt)
    changed = 0
    for n in "UVW":
        w = ref.array("weight" + n)
        assert np.array_equal(rs.constrain(before[n], w, solid[n], fric[n]), ref.array(n)), n
        assert np.array_equal(rs.constrain(saved[n], w, solid[n], fric[n]), ref.array("saved" + n)), n
        changed += int(np.count_nonzero((ref.array(n) != before[n]) & (w > 0) & (w < 1)))
    assert changed > 500            # the friction blend acted on the partly open faces
    ref.close()
