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


def test_particle_container_aliasing_matches_the_harness_model():
    """above 208 333 particles FragmentedVector::operator[] maps some logical indices
    i = k * 208333 onto the slot of i - 208333 (fragmentedvector.h:141-153).  tests/parity_common.aliased_slots is the
    model the GPU parity harness uses for it (the in-place G2P loop updates such a slot twice); here it is pinned to
    the reference itself: the logical view after initialize() holds a duplicate exactly at the modelled indices."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_common as pc
    sc = scenes.dam_break(64)
    n = 3 * pc.FRAGMENT_ELEMENTS + 1000
    reps = -(-n // sc["pos"].shape[0])
    # distinct in-domain positions: copies of the column shifted by a sub-cell amount
    pos = np.concatenate([sc["pos"] + np.float32(1e-3 * r) for r in range(reps)])[:n]
    vel = np.zeros_like(pos)
    e = refengine.RefEngine(sc["dims"], sc["dx"], pos, vel, threads=2)
    got = e.particles()
    assert got.shape[0] == n
    hi, lo = pc.aliased_slots(n)
    assert len(hi) >= 1 and set(hi.tolist()) <= {pc.FRAGMENT_ELEMENTS * k for k in (1, 2, 3)}
    differs = np.flatnonzero((got[:, :3] != pos).any(axis=1))
    assert np.array_equal(differs, hi)                       # every other logical entry is the injected particle
    # slot map s(i): i - 208333 for the modelled indices, i otherwise.  The load queue is a FragmentedVector too and is
    # read by index (fluidsimulation.cpp:2773-2785), so storage slot j holds injected[s(j)], and the logical view reads
    # storage[s(i)] = injected[s(s(i))]
    slot = np.arange(n)
    slot[hi] = lo
    assert np.array_equal(got[:, :3], pos[slot[slot]])
    e.close()


def _wedge(p, w, h, d):
    x, y, z = p
    v = np.array([(x, y, z), (x + w, y, z), (x, y + h, z), (x, y, z + d), (x + w, y, z + d), (x, y + h, z + d)], np.float32)
    t = np.array([(0, 2, 1), (3, 4, 5), (0, 1, 4), (0, 4, 3), (0, 3, 5), (0, 5, 2), (1, 2, 5), (1, 5, 4)], np.int32)
    return v, t


def _octahedron(c, r):
    c = np.array(c, np.float32)
    v = np.array([c + (r, 0, 0), c - (r, 0, 0), c + (0, 0.8 * r, 0), c - (0, 0.8 * r, 0), c + (0, 0, 1.2 * r), c - (0, 0, 1.2 * r)], np.float32)
    t = np.array([(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)], np.int32)
    return v, t


@pytest.mark.parametrize("mesh", ["wedge", "octahedron"])
def test_mesh_signed_distance_field_matches_the_reference_level_set(mesh):
    """flip_mesh_sdf (host code of the library: what the façade hands over for fluid objects, sources and obstacles that
    are not boxes) against MeshLevelSet::fastCalculateSignedDistanceField of the unmodified reference on the same mesh:
    the same sign at every node that is not ON the surface, the same distance within the band the engine reads."""
    from flipengine3d_b200 import engine as fe
    n, dx = 28, 0.125
    v, t = _wedge((0.83, 0.41, 0.67), 1.31, 0.9, 1.7) if mesh == "wedge" else _octahedron((1.77, 1.63, 1.71), 0.93)
    ref = refengine.mesh_level_set((n, n, n), dx, v, t, band=3)
    mine, lo, hi = fe.mesh_sdf((n, n, n), dx, v, t, band=3)
    off = np.abs(ref) > 1e-6
    assert (ref < 0).sum() > 300
    assert np.array_equal((ref < 0)[off], (mine < 0)[off])
    near = np.abs(ref) < 2.5 * dx
    assert near.sum() > 1500 and np.abs(ref - mine)[near].max() <= 1e-6
    # the cell range handed to the seeding kernel covers every node inside the mesh
    kk, jj, ii = np.nonzero(ref < 0)
    assert lo[0] <= ii.min() and ii.max() <= hi[0] and lo[1] <= jj.min() and jj.max() <= hi[1] and lo[2] <= kk.min() and kk.max() <= hi[2]
