"""Lock-step parity harness: drives the reference engine (oracle/_ref, through oracle/refengine.py)
and the CUDA engine (libflip_b200.so through flipengine3d_b200.engine) stage by stage from identical
state and reports per-stage differences.  Test infrastructure only."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from flipengine3d_b200 import engine as fe  # noqa: E402
from oracle import refengine  # noqa: E402

GRID_STAGE_ARRAYS = ("U", "V", "W", "validU", "validV", "validW", "liquid_phi")
ALL_STAGES = ("obstacles", "liquid_sdf", "p2g", "extrapolate_a", "save", "body_force", "pressure",
              "extrapolate_b", "constrain", "g2p", "advance", "tail")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.linalg.norm(b)
    if den == 0.0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / den)


def max_abs(a, b):
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


FRAGMENT_ELEMENTS = int(5e6) // 24     # MarkerParticles per FragmentedVector node  fragmentedvector.h:296,321


def aliased_slots(n):
    """Logical particle indices that share storage in the reference (fragmentedvector.h:141-153):
    operator[] picks the node with `int(i * (1.0 / elementsPerFragment))`, which for most exact multiples
    i = k * 208333 rounds down to k - 1, so logical index i reads AND WRITES the slot of index i - 208333.
    Returns (aliased i's, the indices i - 208333 they share a slot with)."""
    inv = 1.0 / float(FRAGMENT_ELEMENTS)
    hi = [i for i in range(FRAGMENT_ELEMENTS, n, FRAGMENT_ELEMENTS) if int(i * inv) != i // FRAGMENT_ELEMENTS]
    hi = np.asarray(hi, dtype=np.int64)
    return hi, hi - FRAGMENT_ELEMENTS


def solid_velocity_field(shapes):
    """A smooth, nowhere-constant face velocity field for the moving-solid tests: dict(U, V, W) of float32 arrays of the given
    MAC shapes (k, j, i); every face differs from its neighbours, so an index slip shows."""
    out = {}
    for name, (a, b, c, amp) in zip("UVW", ((0.31, 0.17, 0.11, 0.8), (0.13, 0.29, 0.19, 0.6), (0.23, 0.07, 0.37, 0.7))):
        k, j, i = np.meshgrid(*[np.arange(n) for n in shapes[name]], indexing="ij")
        out[name] = (amp * np.sin(a * i + b * j + c * k + 0.3)).astype(np.float32)
    return out


def make_pair(scene, tol=None, threads=None, gravity=(0.0, -25.0, 0.0), preconditioner=None, sampling=None, obstacles=(),
              own_solid=False, settings=None, solid_velocity=False, solid_velocity_early=False, friction=None):
    """Returns (ref, gpu) engines initialised from the same scene; the GPU engine gets the oracle's
    own static solid SDF and the oracle's LOGICAL particle list.  obstacles: (lo, hi) boxes added to
    the reference with addMeshObstacle (their distances arrive with the oracle's solid SDF); own_solid: the GPU engine builds
    its solid SDF itself instead (its domain box, and the obstacles through flip_add_obstacle_box).  solid_velocity: both
    engines get the face velocities of solid_velocity_field for their solids (the reference in the VelocityDataGrid of its
    solid SDF, the GPU engine through flip_set_solid_velocity -- before flip_initialize with solid_velocity_early).
    friction = (boundary, [per obstacle]): setBoundaryFriction / MeshObject::setFriction on the reference; the GPU engine gets
    the same settings with own_solid, else the reference's face friction itself (flip_set_face_friction)."""
    ref = refengine.RefEngine(scene["dims"], scene["dx"], scene["pos"], scene["vel"], gravity=gravity, threads=threads, tol=tol)
    if friction:
        ref.set_boundary_friction(friction[0])
    for q, (lo, hi) in enumerate(obstacles):
        idx = ref.add_obstacle_box(lo, hi)
        if friction:
            ref.set_obstacle_friction(idx, friction[1][q])
    if settings:      # dict(cfl=, picflip=, min_steps=, max_steps=): the step's settings on both engines
        ref.set_step_settings(**settings)
    ref.stage("obstacles", 1.0 / 30.0)   # builds the solid SDF / near-solid grid exactly as the first step would
    vel = None
    if solid_velocity:
        vel = solid_velocity_field({n: ref.shape_of("solid" + n) for n in "UVW"})
        for n in "UVW":
            ref.set_array("solid" + n, vel[n])
    I, J, K = scene["dims"]
    gpu = fe.FluidSimulation(I, J, K, scene["dx"])
    gpu.addBodyForce(*gravity)
    if tol is not None:
        gpu.setPressureSolver(tolerance=tol)
    if preconditioner is not None:
        gpu.setPreconditioner(preconditioner)
    if sampling is not None:
        gpu.setSamplingMode(sampling)
    gpu.enableParticleIds(True)
    if settings:
        if settings.get("cfl"):
            gpu.setCFLConditionNumber(settings["cfl"])
        if settings.get("picflip", -1.0) >= 0.0:
            gpu.setPICFLIPRatio(settings["picflip"])
        if settings.get("min_steps") or settings.get("max_steps"):
            gpu.setTimeStepsPerFrame(settings.get("min_steps") or 1, settings.get("max_steps") or 6)
    if own_solid:
        if friction:
            gpu.setBoundaryFriction(friction[0])
        for q, (lo, hi) in enumerate(obstacles):
            oid = gpu.addMeshObstacleBox(lo, hi)
            if friction:
                gpu.setMeshObstacleFriction(oid, friction[1][q])
    else:
        gpu.setSolidSDF(ref.array("solid_phi"))
        if friction:
            F = ref.face_friction()
            gpu.setFaceFriction(F["U"], F["V"], F["W"])
    if vel is not None and solid_velocity_early:
        gpu.setSolidVelocity(vel["U"], vel["V"], vel["W"])
    gpu.initialize()
    if vel is not None and not solid_velocity_early:
        gpu.setSolidVelocity(vel["U"], vel["V"], vel["W"])
    gpu.setMarkerParticles(ref.particles())
    return ref, gpu


def particles_by_id(gpu):
    """GPU particles re-ordered to the order of the last setMarkerParticles call; returns (aos, ids)."""
    p = gpu.getMarkerParticles()
    ids = gpu.getParticleIds()
    return p, ids


def lockstep_substep(ref, gpu, dt, isolate=True, report=None):
    """Runs one substep of both engines stage by stage.  With isolate=True the GPU's grid state is
    overwritten with the oracle's before every stage, so each stage is compared from identical
    inputs.  Returns a dict of metrics."""
    rep = report if report is not None else {}
    P0 = ref.particles()
    gpu.setMarkerParticles(P0)
    n0 = P0.shape[0]

    def sync_grids():
        for name in GRID_STAGE_ARRAYS:
            gpu.set_array(name, ref.array(name))

    def sync_saved():
        for name in ("savedU", "savedV", "savedW"):
            gpu.set_array(name, ref.array(name))

    def cmp_fields(tag, names=("U", "V", "W")):
        scale = 0.0
        for name in names:
            a, b = gpu.array(name), ref.array(name)
            rep[f"{tag}.{name}.rel_l2"] = rel_l2(a, b)
            rep[f"{tag}.{name}.max_abs"] = max_abs(a, b)
            scale = max(scale, float(np.max(np.abs(b))) if b.size else 0.0)
        rep[f"{tag}.scale"] = scale     # largest |component| of the reference's MAC field at this stage

    def cmp_valid(tag):
        for name in ("validU", "validV", "validW"):
            a, b = gpu.array(name), ref.array(name)
            rep[f"{tag}.{name}.hamming"] = int(np.count_nonzero((a != 0) != (b != 0)))
            rep[f"{tag}.{name}.count"] = int(np.count_nonzero(b))

    # S1/S2: liquid SDF (the GPU produces SDF and P2G in one gather)
    ref.stage("obstacles", dt)
    ref.stage("liquid_sdf", dt)
    gpu.stage("liquid_sdf", dt)
    a, b = gpu.array("liquid_phi"), ref.array("liquid_phi")
    rep["sdf.mismatch_cells"] = int(np.count_nonzero(a != b))
    rep["sdf.max_abs"] = max_abs(a, b)
    rep["sdf.sign_flips"] = int(np.count_nonzero((a < 0) != (b < 0)))
    rep["sdf.liquid_cells"] = int(np.count_nonzero(b < 0))

    # S3a: P2G
    ref.stage("p2g", dt)
    gpu.stage("p2g", dt)
    cmp_fields("p2g")
    cmp_valid("p2g")
    if isolate:
        sync_grids()

    # S3b: extrapolation
    ref.stage("extrapolate_a", dt)
    gpu.stage("extrapolate_a", dt)
    cmp_fields("extrapolate_a")
    for name in ("U", "V", "W"):
        rep[f"extrapolate_a.{name}.mismatch"] = int(np.count_nonzero(gpu.array(name) != ref.array(name)))
    if isolate:
        sync_grids()

    # S4, S5
    ref.stage("save", dt)
    gpu.stage("save", dt)
    ref.stage("body_force", dt)
    gpu.stage("body_force", dt)
    cmp_fields("body_force")
    if isolate:
        sync_grids()
        sync_saved()

    # S7a: pressure
    ref.stage("pressure", dt)
    gpu.stage("pressure", dt)
    cmp_fields("pressure")
    cmp_valid("pressure")
    rep["pressure.ref_iterations"] = ref.pcg_iterations
    rep["pressure.ref_error"] = ref.pcg_error
    if gpu.hasSolidVelocity():
        # moving solids: the velocities as the enclosed-pocket conditioning left them (pressuresolver.cpp:124-244, in place
        # in both engines), and the cell-centre weights that multiply them in the divergence
        for name in "UVW":
            a, b = gpu.array("solid" + name), ref.array("solid" + name)
            rep[f"solid.{name}.mismatch"] = int(np.count_nonzero(a != b))
            rep[f"solid.{name}.zero_faces"] = int(np.count_nonzero(b == 0))
        rep["solid.weightC.mismatch"] = int(np.count_nonzero(gpu.array("weightC") != ref.array("weightC")))
    if isolate:
        sync_grids()

    # S7b
    ref.stage("extrapolate_b", dt)
    gpu.stage("extrapolate_b", dt)
    cmp_fields("extrapolate_b")
    if isolate:
        sync_grids()

    # S8
    ref.stage("constrain", dt)
    gpu.stage("constrain", dt)
    cmp_fields("constrain")
    for name in ("savedU", "savedV", "savedW"):
        rep[f"constrain.{name}.rel_l2"] = rel_l2(gpu.array(name), ref.array(name))
    if isolate:
        sync_grids()
        sync_saved()

    # S10: G2P
    ref.stage("g2p", dt)
    gpu.stage("g2p", dt)
    pg, ids = particles_by_id(gpu)
    pr = ref.particles()
    # Above 208 333 particles some logical indices share a slot in the reference (aliased_slots).  Its G2P loop updates
    # particles IN PLACE by logical index (fluidsimulation.cpp:4082-4092), so a shared slot is updated twice -- or once,
    # the two indices belong to different worker threads -- and both logical entries read the result.  Those few
    # entries (2 per 208 333 particles) are compared against the two possible outcomes instead of the tolerance:
    # the once-updated velocity v1 (what the CUDA path holds) or f(v1) = v1 + 0.95 (v1 - v0).
    a_hi, a_lo = aliased_slots(n0)
    shared = np.zeros(n0, dtype=bool)
    shared[a_hi] = True
    shared[a_lo] = True
    keep = ~shared[ids]
    rep["g2p.shared_slots"] = int(shared.sum())
    if shared.any():
        once = pg[~keep, 3:].astype(np.float64)
        v0 = P0[ids[~keep], 3:].astype(np.float64)
        got = pr[ids[~keep], 3:].astype(np.float64)
        twice = once + 0.95 * (once - v0)
        err = np.minimum(np.abs(got - once), np.abs(got - twice)).max(axis=1)
        scale = max(float(np.abs(once).max()), 1.0)
        rep["g2p.shared_slot_err"] = float(err.max() / scale)
    rep["g2p.vel.rel_l2"] = rel_l2(pg[keep, 3:], pr[ids[keep], 3:])
    rep["g2p.vel.max_abs"] = max_abs(pg[keep, 3:], pr[ids[keep], 3:])
    rep["g2p.vel.mismatch"] = int(np.count_nonzero(pg[keep, 3:] != pr[ids[keep], 3:]))
    if isolate:
        # put the oracle's post-G2P velocities in (keeps ids = index in pr)
        gpu.setMarkerParticles(pr)

    # S12: advance (+ removal)
    ref.stage("advance", dt)
    gpu.stage("advance", dt)
    ref.stage("tail", dt)
    gpu.stage("tail", dt)
    rep["advance.ref_particles"] = ref.num_particles
    rep["advance.gpu_particles"] = gpu.getNumMarkerParticles()
    rep["n0"] = n0
    if ref.num_particles == n0 and gpu.getNumMarkerParticles() == n0:
        pg, ids = particles_by_id(gpu)
        pr2 = ref.particles()
        rep["advance.pos.rel_l2"] = rel_l2(pg[:, :3], pr2[ids, :3])
        rep["advance.pos.max_abs"] = max_abs(pg[:, :3], pr2[ids, :3])
        rep["advance.pos.mismatch"] = int(np.count_nonzero(pg[:, :3] != pr2[ids, :3]))
    return rep


def developed_scene(scene, frames, preconditioner=None):
    """The scene after `frames` free-running frames of the CUDA engine: a developed flow (splash, waves, particles
    against the walls) as the starting state of a lock-step comparison, at sizes where stepping the CPU oracle
    that far would take minutes."""
    I, J, K = scene["dims"]
    sim = fe.FluidSimulation(I, J, K, scene["dx"])
    sim.addBodyForce(0.0, -25.0, 0.0)
    if preconditioner is not None:
        sim.setPreconditioner(preconditioner)
    sim.loadMarkerParticleData(fe.MarkerParticleData(scene["pos"], scene["vel"]))
    sim.initialize()
    for _ in range(frames):
        sim.update(1.0 / 30.0)
    p = sim.getMarkerParticles().copy()
    sim.close()
    return dict(scene, name=f"{scene['name']}+{frames}f", pos=np.ascontiguousarray(p[:, :3]), vel=np.ascontiguousarray(p[:, 3:]))


def lockstep_frames(scene, frames=1, isolate=True, tol=None, threads=None, preconditioner=None, verbose=False,
                    sampling=None, max_substeps=None, obstacles=(), own_solid=False, settings=None, solid_velocity=False,
                    friction=None):
    """max_substeps: stop after that many lock-step substeps in total (the large scenes cost tens of CPU seconds each)."""
    ref, gpu = make_pair(scene, tol=tol, threads=threads, preconditioner=preconditioner, sampling=sampling, obstacles=obstacles,
                         own_solid=own_solid, settings=settings, solid_velocity=solid_velocity, friction=friction)
    reports = []
    for f in range(frames):
        ref.begin_frame(1.0 / 30.0)
        gpu.begin_frame(1.0 / 30.0)
        more = True
        while more:
            if max_substeps is not None and len(reports) >= max_substeps:
                break
            dt = ref.begin_substep()
            gpu.begin_substep()   # bookkeeping only; the oracle's dt is used for both
            rep = {"frame": f, "dt": dt}
            lockstep_substep(ref, gpu, dt, isolate=isolate, report=rep)
            more = ref.end_substep()
            gpu.end_substep()
            st = gpu.substep_stats()[-1]
            rep["gpu.pcg_iterations"] = st["pcg_iterations"]
            rep["gpu.pcg_error"] = st["pcg_error"]
            rep["gpu.pcg_converged"] = st["pcg_converged"]
            rep["gpu.pressure_rows"] = st["pressure_rows"]
            rep["gpu.rhs_max"] = st["rhs_max"]
            rep["ref.fluid_cells"] = ref.num_fluid_cells
            reports.append(rep)
            if verbose:
                print_report(rep)
        ref.end_frame()
        gpu.end_frame()
    ref.close()
    gpu.close()
    return reports


def print_report(rep):
    for k in sorted(rep):
        v = rep[k]
        print(f"  {k:38s} {v:.6g}" if isinstance(v, float) else f"  {k:38s} {v}")
    print(flush=True)


# Tolerances of the parity contract (SURVEY §8d "Parity report"; north_star: integer bookkeeping
# bit-exact, float fields rel L2 <= 1e-4 per step).
TOL_REL_L2 = 1e-4
TOL_P2G_REL_L2 = 1e-5      # P2G differs from the reference by float summation order only
TOL_POS_MAX_ABS_DX = 1e-3  # particle positions: max-abs <= 1e-3 dx
# FLIP_SAMPLING_FAST (the default): exact indices/weights, single-precision 8-point blend -> a few ulp of the
# sampled velocities per sample instead of bit-identity
TOL_FAST_VEL_REL_L2 = 5e-6
TOL_FAST_POS_REL_L2 = 1e-6
# A component that is itself rounding noise (e.g. U and W of a column that only falls: |U| ~ 1e-7 next to
# |V| ~ 1) has no meaningful relative error: it passes if its max-abs error is below the float-storage floor of
# the velocities, 1e-5 * max|u| (SURVEY §8d parity report).
TOL_FIELD_FLOOR = 1e-5


def field_ok(rep, tag, comp, tol):
    return rep[f"{tag}.{comp}.rel_l2"] <= tol or rep[f"{tag}.{comp}.max_abs"] <= TOL_FIELD_FLOOR * rep.get(f"{tag}.scale", 0.0)


def check_report(rep, dx=0.125, isolate=True, exact_sampling=False):
    """Asserts the parity contract on one lock-step substep report.  exact_sampling: the engine ran with
    FLIP_SAMPLING_EXACT, so G2P and RK3 must be bit-identical from identical inputs."""
    # integer bookkeeping: exact
    assert rep["sdf.sign_flips"] == 0, rep
    assert rep["gpu.pressure_rows"] == rep["ref.fluid_cells"] or rep["ref.fluid_cells"] == 0, rep
    assert rep["advance.gpu_particles"] == rep["advance.ref_particles"], rep
    for comp in "UVW":
        assert rep[f"p2g.valid{comp}.hamming"] == 0, rep
        assert rep[f"pressure.valid{comp}.hamming"] == 0, rep
    # order-independent stages: bit-exact from identical inputs
    assert rep["sdf.mismatch_cells"] == 0, rep
    if "solid.weightC.mismatch" in rep:
        assert rep["solid.weightC.mismatch"] == 0, rep
        for comp in "UVW":
            assert rep[f"solid.{comp}.mismatch"] == 0, rep
    if isolate:
        for comp in "UVW":
            assert rep[f"extrapolate_a.{comp}.mismatch"] == 0, rep
            assert rep[f"extrapolate_b.{comp}.max_abs"] == 0.0, rep
            assert rep[f"body_force.{comp}.max_abs"] == 0.0, rep
            assert rep[f"constrain.{comp}.max_abs"] == 0.0, rep
        if exact_sampling:
            assert rep["g2p.vel.mismatch"] == 0, rep
            if "advance.pos.mismatch" in rep:
                assert rep["advance.pos.mismatch"] == 0, rep
        else:
            assert rep["g2p.vel.rel_l2"] <= TOL_FAST_VEL_REL_L2, rep
            if "advance.pos.rel_l2" in rep:
                assert rep["advance.pos.rel_l2"] <= TOL_FAST_POS_REL_L2, rep
    # entries that share a slot in the reference's particle container: once- or twice-updated, nothing else
    if rep.get("g2p.shared_slots", 0):
        assert rep["g2p.shared_slots"] <= 2 * (rep["n0"] // FRAGMENT_ELEMENTS), rep
        assert rep["g2p.shared_slot_err"] <= 1e-5, rep
    # float fields
    for comp in "UVW":
        assert field_ok(rep, "p2g", comp, TOL_P2G_REL_L2), rep
        assert field_ok(rep, "pressure", comp, TOL_REL_L2), rep
        assert field_ok(rep, "extrapolate_b", comp, TOL_REL_L2), rep
    assert rep["g2p.vel.rel_l2"] <= TOL_REL_L2, rep
    if "advance.pos.rel_l2" in rep:
        assert rep["advance.pos.rel_l2"] <= TOL_REL_L2, rep
        assert rep["advance.pos.max_abs"] <= TOL_POS_MAX_ABS_DX * dx, rep
    assert rep["gpu.pcg_converged"] == 1, rep


def restatement_check(sc, frames=3):
    """The CUDA stages against oracle/restatement.py on a seeded scene, from the CUDA path's own inputs: needs
    neither oracle/_ref nor /root/reference.  Order-independent stages bit for bit, P2G and the projection to
    their tolerances (literal P2G weights and double-precision sampling: FLIP_SAMPLING_EXACT)."""
    from oracle import restatement as R
    dims, dx = sc["dims"], sc["dx"]
    sim = fe.FluidSimulation(*dims, dx)
    sim.addBodyForce(0, -25, 0)
    sim.setSamplingMode("exact")
    sim.enableParticleIds(True)
    sim.loadMarkerParticleData(fe.MarkerParticleData(sc["pos"], sc["vel"]))
    sim.initialize()
    for _ in range(frames):
        sim.update(1 / 30)
    sim.begin_frame(1 / 30)
    dt = sim.begin_substep()
    P = sim.getMarkerParticles().copy()
    sim.stage("obstacles", dt)
    solid = sim.array("solid_phi")
    sim.stage("liquid_sdf", dt)
    assert np.array_equal(sim.array("liquid_phi"), R.liquid_sdf(P[:, :3], dims, dx, solid))
    sim.stage("p2g", dt)
    ref = R.p2g(P[:, :3], P[:, 3:], dims, dx)
    scale = max(float(np.abs(a).max()) for a in ref[:3])
    for n, a, v in zip("UVW", ref[:3], ref[3:]):
        assert np.array_equal(sim.array(f"valid{n}") != 0, v != 0), n
        assert rel_l2(sim.array(n), a) <= 1e-5 or max_abs(sim.array(n), a) <= 1e-5 * scale, n
    fields = {n: sim.array(n) for n in "UVW"}
    valids = {n: sim.array(f"valid{n}") for n in "UVW"}
    sim.stage("extrapolate_a", dt)
    for n in "UVW":
        assert np.array_equal(sim.array(n), R.extrapolate(fields[n], valids[n])), n
    sim.stage("save", dt)
    before = {n: sim.array(n) for n in "UVW"}
    sim.stage("body_force", dt)
    assert np.array_equal(sim.array("V"), R.body_force(before["V"], -25.0, dt))
    pre = {n: sim.array(n) for n in "UVW"}
    phi = sim.array("liquid_phi")
    w = {n: sim.array(f"weight{n}") for n in "UVW"}
    sim.stage("pressure", dt)
    U, V, W, vU, vV, vW, rows, bmax = R.pressure_project(pre["U"], pre["V"], pre["W"], phi, w["U"], w["V"], w["W"], dims, dx, dt)
    scale = max(float(np.abs(a).max()) for a in (U, V, W))
    for n, a, v in (("U", U, vU), ("V", V, vV), ("W", W, vW)):
        assert np.array_equal(sim.array(f"valid{n}") != 0, v != 0), n
        assert rel_l2(sim.array(n), a) <= 1e-4 or max_abs(sim.array(n), a) <= 1e-5 * scale, n
    fields = {n: sim.array(n) for n in "UVW"}
    valids = {n: sim.array(f"valid{n}") for n in "UVW"}
    sim.stage("extrapolate_b", dt)
    for n in "UVW":
        assert np.array_equal(sim.array(n), R.extrapolate(fields[n], valids[n])), n
    fields = {n: sim.array(n) for n in "UVW"}
    saved = {n: sim.array(f"saved{n}") for n in "UVW"}
    sim.stage("constrain", dt)
    for n in "UVW":
        assert np.array_equal(sim.array(n), R.constrain(fields[n], w[n])), n
        assert np.array_equal(sim.array(f"saved{n}"), R.constrain(saved[n], w[n])), n
    new = tuple(sim.array(n) for n in "UVW")
    old = tuple(sim.array(f"saved{n}") for n in "UVW")
    ids0 = sim.getParticleIds().copy()
    sim.stage("g2p", dt)
    P1 = sim.getMarkerParticles().copy()
    assert np.array_equal(sim.getParticleIds(), ids0)
    assert np.array_equal(P1[:, 3:], R.g2p(P[:, :3], P[:, 3:], new, old, dims, dx))
    sim.stage("advance", dt)
    P2, ids2 = sim.getMarkerParticles(), sim.getParticleIds()
    # the advance stage re-sorts the store: match by particle id.  RK3 + _resolveCollision, bit for bit
    ns = sim.array("near_solid")
    nsn = [-(-d // 3) for d in dims]                      # coarse cells of 3dx (fluidsimulation.cpp:3083-3092)
    p1 = R.advance(P[:, :3], new, solid, ns.reshape(nsn[2], nsn[1], nsn[0]), dims, dx, dt)
    lookup = np.full(int(ids0.max()) + 1, -1, dtype=np.int64)
    lookup[ids0] = np.arange(ids0.size)
    src = lookup[ids2]
    assert np.array_equal(P2[:, :3], p1[src])
    sim.end_substep(); sim.end_frame(); sim.close()
