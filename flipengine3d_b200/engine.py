"""ctypes binding of libflip_b200.so and a Python mirror of the reference's FluidSimulation API.

The product is the C-ABI library (include/flip_b200.h); this module is the thin host-side mirror
used by tests/ and bench.py.  It never falls back to a CPU implementation: if the library is
missing, or no CUDA device is present, construction raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflip_b200.so")

STAGES = ("obstacles", "liquid_sdf", "p2g", "extrapolate_a", "save", "body_force", "pressure",
          "extrapolate_b", "constrain", "g2p", "advance", "tail")
STAGE_ID = {n: i for i, n in enumerate(STAGES)}
ARRAY_ID = dict(U=0, V=1, W=2, validU=3, validV=4, validW=5, liquid_phi=6, solid_phi=7,
                weightU=8, weightV=9, weightW=10, weightC=11, savedU=12, savedV=13, savedW=14,
                near_solid=15, pressure=16, solidU=17, solidV=18, solidW=19)

KERNEL_CLASSES = ("sdf_p2g", "g2p", "advance", "sort", "extrapolate", "pcg_spmv", "pcg_iter", "pressure_build",
                  "pressure_apply", "precond", "pcg_solve", "pcg_dir_spmv", "g2p_advance")

FLIP_OK, FLIP_ERR_RUNTIME, FLIP_ERR_DOMAIN, FLIP_ERR_OUT_OF_RANGE, FLIP_ERR_CUDA, FLIP_ERR_UNSUPPORTED = range(6)


class FlipCudaError(RuntimeError):
    pass


class FlipUnsupported(RuntimeError):
    pass


# the exception types FluidSimulation throws for the same misuse (SURVEY §8b "Errors")
_EXC = {FLIP_ERR_RUNTIME: RuntimeError, FLIP_ERR_DOMAIN: ValueError, FLIP_ERR_OUT_OF_RANGE: IndexError,
        FLIP_ERR_CUDA: FlipCudaError, FLIP_ERR_UNSUPPORTED: FlipUnsupported}


class StepStats(C.Structure):
    _fields_ = [("particles", C.c_int32), ("fluid_cells", C.c_int32), ("pressure_rows", C.c_int32),
                ("pcg_iterations", C.c_int32), ("pcg_converged", C.c_int32), ("removed_solid", C.c_int32),
                ("removed_crowded", C.c_int32), ("removed_fast", C.c_int32), ("pcg_error", C.c_double),
                ("rhs_max", C.c_double), ("dt", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_lib = None


def load_library():
    """Loads libflip_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    L.flip_create.argtypes = [C.POINTER(vp), ci, ci, ci, cd, ci]
    L.flip_destroy.argtypes = [vp]
    L.flip_destroy.restype = None
    L.flip_last_error.argtypes = [vp]
    L.flip_last_error.restype = C.c_char_p
    L.flip_create_error.restype = C.c_char_p
    L.flip_add_body_force.argtypes = [vp, cd, cd, cd]
    L.flip_set_pic_flip_ratio.argtypes = [vp, cd]
    L.flip_set_cfl.argtypes = [vp, cd]
    L.flip_set_substep_limits.argtypes = [vp, ci, ci]
    L.flip_set_pressure_solver.argtypes = [vp, cd, cd, ci]
    L.flip_set_preconditioner.argtypes = [vp, ci]
    L.flip_set_multigrid.argtypes = [vp, ci, cd, cd, ci]
    L.flip_set_solver_mode.argtypes = [vp, ci]
    L.flip_set_pressure_warm_start.argtypes = [vp, ci]
    L.flip_set_multigrid_schedule.argtypes = [vp, ci, C.POINTER(cd)]
    L.flip_set_sampling_mode.argtypes = [vp, ci]
    L.flip_load_particles.argtypes = [vp, ci, vp, vp]
    L.flip_add_fluid_box.argtypes = [vp, C.POINTER(cd), C.POINTER(cd), C.POINTER(cd)]
    L.flip_add_fluid_sdf.argtypes = [vp, vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(cd)]
    L.flip_add_fluid_source_box.argtypes = [vp, ci, C.POINTER(cd), C.POINTER(cd), C.POINTER(cd), C.POINTER(ci)]
    L.flip_add_fluid_source_sdf.argtypes = [vp, ci, vp, C.POINTER(ci), C.POINTER(ci), C.POINTER(cd), C.POINTER(ci)]
    L.flip_enable_fluid_source.argtypes = [vp, ci, ci]
    L.flip_remove_fluid_source.argtypes = [vp, ci]
    L.flip_constrain_fluid_source_velocity.argtypes = [vp, ci, ci]
    L.flip_mesh_sdf.argtypes = [ci, ci, ci, cd, vp, ci, vp, ci, ci, C.c_float, vp, C.POINTER(ci), C.POINTER(ci)]
    L.flip_reset_body_force.argtypes = [vp]
    L.flip_set_extreme_velocity_removal.argtypes = [vp, ci]
    L.flip_set_marker_particle_scale.argtypes = [vp, cd]
    L.flip_add_obstacle_box.argtypes = [vp, C.POINTER(cd), C.POINTER(cd), C.POINTER(ci)]
    L.flip_add_obstacle_sdf.argtypes = [vp, vp, C.POINTER(ci)]
    L.flip_enable_obstacle.argtypes = [vp, ci, ci]
    L.flip_remove_obstacle.argtypes = [vp, ci]
    L.flip_set_surface_subdivision_level.argtypes = [vp, ci]
    L.flip_set_surface_smoothing.argtypes = [vp, cd, ci]
    L.flip_get_isomesh_size.argtypes = [vp, C.POINTER(ci), C.POINTER(ci)]
    L.flip_get_isomesh.argtypes = [vp, vp, vp]
    L.flip_get_isomesh_field.argtypes = [vp, vp, vp, vp]
    L.flip_add_marker_particle.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.flip_set_solid_sdf.argtypes = [vp, vp]
    L.flip_initialize.argtypes = [vp]
    L.flip_update.argtypes = [vp, cd]
    L.flip_get_current_frame.argtypes = [vp, C.POINTER(ci)]
    L.flip_set_current_frame.argtypes = [vp, ci]
    L.flip_static_inputs.argtypes = [ci, ci, ci, C.c_double, vp, ci, vp, vp, vp, vp, C.POINTER(ci)]
    L.flip_center_weights.argtypes = [ci, ci, ci, C.c_double, vp, vp]
    L.flip_face_friction.argtypes = [ci, ci, ci, C.c_double, ci, ci, vp, vp, vp, vp, vp]
    L.flip_box_obstacle_sdf.argtypes = [ci, ci, ci, C.c_double, ci, C.POINTER(C.c_double), C.POINTER(C.c_double), vp]
    L.flip_set_solid_velocity.argtypes = [vp, vp, vp, vp]
    L.flip_add_obstacle_mesh.argtypes = [vp, vp, ci, vp, ci, C.POINTER(ci)]
    L.flip_set_obstacle_mesh_motion.argtypes = [vp, ci, vp, vp, vp]
    L.flip_mesh_velocity_data.argtypes = [ci, ci, ci, cd, vp, ci, vp, ci, vp, ci, C.c_float, vp, vp, vp, vp, vp, vp, vp]
    L.flip_set_boundary_friction.argtypes = [vp, cd]
    L.flip_set_obstacle_friction.argtypes = [vp, ci, cd]
    L.flip_set_face_friction.argtypes = [vp, vp, vp, vp]
    L.flip_get_face_friction.argtypes = [vp, vp, vp, vp]
    L.flip_set_obstacle_box_motion.argtypes = [vp, ci, C.POINTER(cd), C.POINTER(cd), C.POINTER(cd)]
    L.flip_get_num_substeps.argtypes = [vp, C.POINTER(ci)]
    L.flip_get_step_stats.argtypes = [vp, ci, C.POINTER(StepStats)]
    L.flip_get_num_particles.argtypes = [vp, C.POINTER(ci)]
    L.flip_get_particles.argtypes = [vp, vp, ci]
    L.flip_set_particles.argtypes = [vp, ci, vp]
    L.flip_get_particle_positions.argtypes = [vp, vp, ci]
    L.flip_get_particle_velocities.argtypes = [vp, vp, ci]
    L.flip_get_velocity_field.argtypes = [vp, vp, vp, vp]
    L.flip_enable_particle_ids.argtypes = [vp, ci]
    L.flip_set_particle_id_base.argtypes = [vp, ci]
    L.flip_get_particle_ids.argtypes = [vp, vp, ci]
    L.flip_begin_frame.argtypes = [vp, cd]
    L.flip_begin_substep.argtypes = [vp, C.POINTER(cd)]
    L.flip_run_stage.argtypes = [vp, ci, cd]
    L.flip_end_substep.argtypes = [vp, C.POINTER(ci)]
    L.flip_end_frame.argtypes = [vp]
    L.flip_array_bytes.argtypes = [vp, ci, C.POINTER(C.c_int64)]
    L.flip_get_array.argtypes = [vp, ci, vp]
    L.flip_set_array.argtypes = [vp, ci, vp]
    L.flip_get_stage_times_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.flip_get_kernel_launches.argtypes = [vp, C.POINTER(C.c_int64)]
    L.flip_enable_kernel_timing.argtypes = [vp, ci]
    L.flip_reset_kernel_timing.argtypes = [vp]
    L.flip_get_kernel_timing.argtypes = [vp, ci, C.POINTER(cd), C.POINTER(C.c_int64)]
    L.flip_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.flip_synchronize.argtypes = [vp]
    L.flip_set_slab.argtypes = [vp, ci, ci, vp, ci]
    L.flip_get_nccl_unique_id.argtypes = [vp, ci]
    L.flip_slab_range.argtypes = [ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]
    L.flip_get_slab_info.argtypes = [vp] + [C.POINTER(ci)] * 4
    L.flip_set_halo.argtypes = [vp, ci]
    _lib = L
    return L


def static_inputs(isize, jsize, ksize, dx, solid_phi=None):
    """Host-side static inputs of a box domain (flip_static_inputs; no CUDA device needed):
    dict(solid_phi (K+1,J+1,I+1), weightU, weightV, weightW, near_solid (nk,nj,ni)).  solid_phi given: the weights and
    the mask are derived from it (as after flip_set_solid_sdf) instead of from the built-in box."""
    L = load_library()
    I, J, K = int(isize), int(jsize), int(ksize)
    given = 0 if solid_phi is None else 1
    phi = np.empty((K + 1, J + 1, I + 1), dtype=np.float32) if solid_phi is None else np.ascontiguousarray(solid_phi, dtype=np.float32).copy()
    assert phi.shape == (K + 1, J + 1, I + 1)
    wU = np.empty((K, J, I + 1), dtype=np.float32)
    wV = np.empty((K, J + 1, I), dtype=np.float32)
    wW = np.empty((K + 1, J, I), dtype=np.float32)
    nd = (C.c_int * 3)()
    rc = L.flip_static_inputs(I, J, K, float(dx), phi.ctypes.data if given else None, given, None, None, None, None, nd)
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)("flip_static_inputs failed")
    ns = np.empty((nd[2], nd[1], nd[0]), dtype=np.uint8)
    rc = L.flip_static_inputs(I, J, K, float(dx), phi.ctypes.data, given, wU.ctypes.data, wV.ctypes.data, wW.ctypes.data, ns.ctypes.data, nd)
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)("flip_static_inputs failed")
    return dict(solid_phi=phi, weightU=wU, weightV=wV, weightW=wW, near_solid=ns)


def box_obstacle_sdf(dims, dx, lo, hi, band=3):
    """flip_box_obstacle_sdf (host code): the nodal field (K+1, J+1, I+1) of a box obstacle, FLT_MAX outside its band."""
    L = load_library()
    I, J, K = (int(d) for d in dims)
    phi = np.empty((K + 1, J + 1, I + 1), dtype=np.float32)
    d3 = C.c_double * 3
    rc = L.flip_box_obstacle_sdf(I, J, K, float(dx), int(band), d3(*lo), d3(*hi), phi.ctypes.data)
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)("flip_box_obstacle_sdf failed")
    return phi


def face_friction(dims, dx, phis, frictions, band=3):
    """flip_face_friction (host code): dict(U, V, W) of the face friction of solids given as nodal fields in merge order
    (phis[0]: the domain) with their frictions."""
    L = load_library()
    I, J, K = (int(d) for d in dims)
    arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in phis]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    fr = np.ascontiguousarray(frictions, dtype=np.float32)
    out = dict(U=np.empty((K, J, I + 1), np.float32), V=np.empty((K, J + 1, I), np.float32), W=np.empty((K + 1, J, I), np.float32))
    rc = L.flip_face_friction(I, J, K, float(dx), int(band), len(arrs), ptrs, fr.ctypes.data, out["U"].ctypes.data, out["V"].ctypes.data,
                              out["W"].ctypes.data)
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)("flip_face_friction failed")
    return out


def center_weights(dims, dx, solid_phi):
    """flip_center_weights (host code, no CUDA device needed): the cell-centre weights (K, J, I) of a nodal solid SDF."""
    L = load_library()
    I, J, K = (int(d) for d in dims)
    phi = np.ascontiguousarray(solid_phi, dtype=np.float32)
    assert phi.shape == (K + 1, J + 1, I + 1)
    wC = np.empty((K, J, I), dtype=np.float32)
    rc = L.flip_center_weights(I, J, K, float(dx), phi.ctypes.data, wC.ctypes.data)
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)("flip_center_weights failed")
    return wC


def mesh_velocity_data(dims, dx, vertices, triangles, vertex_velocities, band=3, far=0.0):
    """flip_mesh_velocity_data (host code, no CUDA device needed): dict(phi (K+1,J+1,I+1), fractionU/V/W, fieldU/V/W in the MAC
    shapes) of a closed mesh whose vertices move with vertex_velocities."""
    L = load_library()
    I, J, K = (int(d) for d in dims)
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    t = np.ascontiguousarray(triangles, dtype=np.int32)
    w = np.ascontiguousarray(vertex_velocities, dtype=np.float32)
    assert w.shape == v.shape
    out = dict(phi=np.empty((K + 1, J + 1, I + 1), dtype=np.float32))
    shapes = dict(U=(K, J, I + 1), V=(K, J + 1, I), W=(K + 1, J, I))
    for kind in ("fraction", "field"):
        for n in "UVW":
            out[kind + n] = np.empty(shapes[n], dtype=np.float32)
    rc = L.flip_mesh_velocity_data(I, J, K, float(dx), v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], w.ctypes.data, int(band),
                                   float(far), out["phi"].ctypes.data, *[out[k + n].ctypes.data for k in ("fraction", "field") for n in "UVW"])
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)("flip_mesh_velocity_data failed")
    return out


def mesh_sdf(dims, dx, vertices, triangles, band=3, far=0.0):
    """flip_mesh_sdf (host code, no CUDA device needed): the nodal signed distance field of a closed triangle mesh on the
    grid -- (phi of shape (K+1, J+1, I+1), cell_lo, cell_hi)."""
    L = load_library()
    I, J, K = (int(d) for d in dims)
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    t = np.ascontiguousarray(triangles, dtype=np.int32)
    phi = np.empty((K + 1, J + 1, I + 1), dtype=np.float32)
    lo, hi = (C.c_int * 3)(), (C.c_int * 3)()
    rc = L.flip_mesh_sdf(I, J, K, float(dx), v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], int(band), float(far), phi.ctypes.data, lo, hi)
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)("flip_mesh_sdf failed")
    return phi, tuple(lo), tuple(hi)


def nccl_unique_id():
    """128-byte ncclUniqueId (rank 0 creates it, the caller distributes it)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.flip_get_nccl_unique_id(buf, 128)
    if rc != FLIP_OK:
        raise _EXC.get(rc, RuntimeError)(L.flip_create_error().decode())
    return bytes(buf.raw)


def slab_range(K, nranks, rank):
    """Owned global planes [k0, k1) of `rank` (flip_slab_range)."""
    L = load_library()
    a, b = C.c_int(), C.c_int()
    rc = L.flip_slab_range(int(K), int(nranks), int(rank), C.byref(a), C.byref(b))
    if rc != FLIP_OK:
        raise IndexError("bad rank / nranks")
    return a.value, b.value


class MarkerParticleData:
    """FluidSimulationMarkerParticleData (fluidsimulation.h:102-106): float xyz triplets."""

    def __init__(self, positions, velocities):
        self.positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        self.velocities = np.ascontiguousarray(velocities, dtype=np.float32).reshape(-1, 3)
        assert self.positions.shape == self.velocities.shape
        self.size = self.positions.shape[0]


class FluidSimulation:
    """Mirror of the subset of the reference's FluidSimulation that FluidManager and the north star
    call: constructor, addBodyForce, addMeshFluid (axis-aligned boxes),
    loadMarkerParticleData, initialize, update, getCurrentFrame, getNumMarkerParticles,
    getMarkerParticles, getMarkerParticlePositionData/VelocityData, getVelocityField."""

    def __init__(self, isize, jsize, ksize, dx, device=0):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.flip_create(C.byref(self.h), int(isize), int(jsize), int(ksize), float(dx), int(device))
        if rc != FLIP_OK:
            raise _EXC.get(rc, RuntimeError)(self.L.flip_create_error().decode())
        self.dims = (int(isize), int(jsize), int(ksize))
        self.dx = float(dx)
        self.local_K, self.k_offset, self.k_own = int(ksize), 0, (0, int(ksize))

    # -- plumbing
    def _check(self, rc):
        if rc != FLIP_OK:
            raise _EXC.get(rc, RuntimeError)(self.L.flip_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.flip_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- z-slab decomposition (one process per GPU)
    def setSlab(self, rank, nranks, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self.L.flip_set_slab(self.h, int(rank), int(nranks), buf, 128))
        ko, kl, a, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.flip_get_slab_info(self.h, C.byref(ko), C.byref(kl), C.byref(a), C.byref(b)))
        self.local_K, self.k_offset, self.k_own = kl.value, ko.value, (a.value, b.value)

    def setHalo(self, planes):
        self._check(self.L.flip_set_halo(self.h, int(planes)))

    # -- configuration
    def addBodyForce(self, fx, fy, fz):
        self._check(self.L.flip_add_body_force(self.h, fx, fy, fz))

    def setPICFLIPRatio(self, r):
        self._check(self.L.flip_set_pic_flip_ratio(self.h, r))

    def setCFLConditionNumber(self, n):
        self._check(self.L.flip_set_cfl(self.h, n))

    def setTimeStepsPerFrame(self, mn, mx):
        self._check(self.L.flip_set_substep_limits(self.h, mn, mx))

    def setPressureSolver(self, tolerance=1e-9, acceptable=1.0, max_iterations=1000):
        self._check(self.L.flip_set_pressure_solver(self.h, tolerance, acceptable, max_iterations))

    def setPreconditioner(self, kind):
        self._check(self.L.flip_set_preconditioner(self.h, {"jacobi": 0, "multigrid": 1}.get(kind, kind)))

    def setMultigrid(self, sweeps=2, damping=0.8, coarse_weight=1.8, coarsest_sweeps=8):
        self._check(self.L.flip_set_multigrid(self.h, sweeps, damping, coarse_weight, coarsest_sweeps))

    def setMultigridSchedule(self, dampings):
        """Per-sweep damping factors of the pre-smoothing passes (post-smoothing mirrors them)."""
        arr = (C.c_double * len(dampings))(*[float(w) for w in dampings])
        self._check(self.L.flip_set_multigrid_schedule(self.h, len(dampings), arr))

    def setPressureWarmStart(self, on=True):
        self._check(self.L.flip_set_pressure_warm_start(self.h, 1 if on else 0))

    def setSamplingMode(self, mode):
        """'exact': the reference's double-precision trilinear blend (bit-identical G2P / RK3);
        'fast' (default): exact indices and weights, single-precision blend."""
        kinds = {"exact": 0, "fast": 1}
        if mode not in kinds:
            raise ValueError(f"sampling mode must be one of {sorted(kinds)}")
        self._check(self.L.flip_set_sampling_mode(self.h, kinds[mode]))

    def setSolverMode(self, persistent=True):
        self._check(self.L.flip_set_solver_mode(self.h, 1 if persistent else 0))

    def addMeshFluidSourceBox(self, lo, hi, velocity=(0.0, 0.0, 0.0), outflow=False):
        """addMeshFluidSource with a static box MeshFluidSource (inflow, or outflow=True); returns the source id."""
        a, b, v = (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_double * 3)(*velocity)
        sid = C.c_int()
        self._check(self.L.flip_add_fluid_source_box(self.h, 1 if outflow else 0, a, b, v, C.byref(sid)))
        return sid.value

    def enableMeshFluidSource(self, sid, on=True):
        self._check(self.L.flip_enable_fluid_source(self.h, int(sid), 1 if on else 0))

    def meshSDF(self, vertices, triangles, band=3, far=0.0):
        """flip_mesh_sdf (host code): (phi of shape (K+1, J+1, I+1), cell_lo, cell_hi) of a closed triangle mesh."""
        return mesh_sdf(self.dims, self.dx, vertices, triangles, band, far)

    def addMeshFluidMesh(self, vertices, triangles, velocity=(0.0, 0.0, 0.0)):
        """addMeshFluid(MeshObject) for a mesh that is not a box: host signed distance field -> flip_add_fluid_sdf."""
        phi, lo, hi = self.meshSDF(vertices, triangles)
        self.addMeshFluidSDF(phi, velocity, lo, hi)

    def setBoundaryFriction(self, f):
        """FluidSimulation::setBoundaryFriction (ValueError outside [0, 1])."""
        self._check(self.L.flip_set_boundary_friction(self.h, float(f)))

    def setMeshObstacleFriction(self, oid, f):
        """MeshObject::setFriction of an obstacle."""
        self._check(self.L.flip_set_obstacle_friction(self.h, int(oid), float(f)))

    def setFaceFriction(self, U=None, V=None, W=None):
        """flip_set_face_friction: the face friction of the constraint handed in directly; no arguments: derived again."""
        if U is None and V is None and W is None:
            self._check(self.L.flip_set_face_friction(self.h, None, None, None))
            return
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (U, V, W)]
        for x, name in zip(a, ("solidU", "solidV", "solidW")):
            assert x.shape == self.shape_of(name), (name, x.shape)
        self._check(self.L.flip_set_face_friction(self.h, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data))

    def getFaceFriction(self):
        out = {n: np.empty(self.shape_of("solid" + n), dtype=np.float32) for n in "UVW"}
        self._check(self.L.flip_get_face_friction(self.h, out["U"].ctypes.data, out["V"].ctypes.data, out["W"].ctypes.data))
        return out

    def addMeshObstacleMesh(self, vertices, triangles):
        phi, _, _ = self.meshSDF(vertices, triangles, band=3, far=3.0e38)
        return self.addMeshObstacleSDF(phi)

    def resetBodyForce(self):
        self._check(self.L.flip_reset_body_force(self.h))

    def enableExtremeVelocityRemoval(self, on=True):
        self._check(self.L.flip_set_extreme_velocity_removal(self.h, 1 if on else 0))

    def setMarkerParticleScale(self, s):
        self._check(self.L.flip_set_marker_particle_scale(self.h, float(s)))

    def addMeshObstacleBox(self, lo, hi):
        """FluidSimulation::addMeshObstacle with a static box MeshObject; returns the obstacle's handle."""
        oid = C.c_int()
        self._check(self.L.flip_add_obstacle_box(self.h, (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), C.byref(oid)))
        return oid.value

    def addMeshObstacleSDF(self, nodal_sdf):
        """... with the nodal signed distance field of the obstacle, shape (K+1, J+1, I+1), negative inside."""
        a = np.ascontiguousarray(nodal_sdf, dtype=np.float32)
        I, J, K = self.dims
        assert a.size == (I + 1) * (J + 1) * (K + 1)
        oid = C.c_int()
        self._check(self.L.flip_add_obstacle_sdf(self.h, a.ctypes.data, C.byref(oid)))
        return oid.value

    def setSolidVelocity(self, U=None, V=None, W=None):
        """flip_set_solid_velocity: the face velocities of the solids (MeshLevelSet::getFaceVelocityU/V/W), shapes
        (K, J, I+1), (K, J+1, I), (K+1, J, I); no arguments: every solid at rest again."""
        if U is None and V is None and W is None:
            self._check(self.L.flip_set_solid_velocity(self.h, None, None, None))
            self._solid_velocity = False
            return
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (U, V, W)]
        for x, name in zip(a, ("solidU", "solidV", "solidW")):
            assert x.shape == self.shape_of(name), (name, x.shape)
        self._check(self.L.flip_set_solid_velocity(self.h, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data))
        self._solid_velocity = True

    def addMeshObstacleMesh(self, vertices, triangles):
        """FluidSimulation::addMeshObstacle with a closed triangle mesh the library keeps (it can be animated)."""
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        oid = C.c_int()
        self._check(self.L.flip_add_obstacle_mesh(self.h, v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], C.byref(oid)))
        return oid.value

    def setMeshObstacleMeshMotion(self, oid, prev, cur, nxt):
        """MeshObject::updateMeshAnimated for a mesh obstacle: the vertices of the previous / current / next frame."""
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (prev, cur, nxt)]
        self._check(self.L.flip_set_obstacle_mesh_motion(self.h, int(oid), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data))
        self._solid_velocity = True

    def setMeshObstacleBoxMotion(self, oid, off_prev, off_cur, off_next):
        """MeshObject::updateMeshAnimated for a box obstacle that translates: the offsets of the previous, the current and
        the next frame's mesh against the box it was added as (call once per frame, before update)."""
        d3 = C.c_double * 3
        self._check(self.L.flip_set_obstacle_box_motion(self.h, int(oid), d3(*off_prev), d3(*off_cur), d3(*off_next)))
        self._solid_velocity = True

    def hasSolidVelocity(self):
        return bool(getattr(self, "_solid_velocity", False))

    def enableMeshObstacle(self, oid, on=True):
        self._check(self.L.flip_enable_obstacle(self.h, int(oid), 1 if on else 0))

    def removeMeshObstacle(self, oid):
        self._check(self.L.flip_remove_obstacle(self.h, int(oid)))

    def constrainMeshFluidSourceVelocity(self, sid, on=True):
        """MeshFluidSource::enable/disableConstrainedFluidVelocity (on by default)."""
        self._check(self.L.flip_constrain_fluid_source_velocity(self.h, int(sid), 1 if on else 0))

    def removeMeshFluidSource(self, sid):
        self._check(self.L.flip_remove_fluid_source(self.h, int(sid)))

    def getSimulationDimensions(self):
        return tuple(d * self.dx for d in self.dims)

    def getGridDimensions(self):
        return self.dims

    def getCellSize(self):
        return self.dx

    def addMeshFluidBox(self, lo, hi, velocity=(0.0, 0.0, 0.0)):
        a = (C.c_double * 3)(*lo)
        b = (C.c_double * 3)(*hi)
        v = (C.c_double * 3)(*velocity)
        self._check(self.L.flip_add_fluid_box(self.h, a, b, v))

    def addMeshFluidSDF(self, nodal_sdf, velocity=(0.0, 0.0, 0.0), cell_lo=None, cell_hi=None):
        """addMeshFluid(MeshObject, velocity) with the object given as the nodal signed distance field of the grid
        (negative inside), shape (K+1, J+1, I+1): what MeshLevelSet::fastCalculateSignedDistanceField produces."""
        I, J, K = self.dims
        a = np.ascontiguousarray(nodal_sdf, dtype=np.float32)
        if a.size != (I + 1) * (J + 1) * (K + 1):
            raise ValueError("nodal_sdf must have (I+1)(J+1)(K+1) entries")
        lo = (C.c_int * 3)(*cell_lo) if cell_lo is not None else None
        hi = (C.c_int * 3)(*cell_hi) if cell_hi is not None else None
        v = (C.c_double * 3)(*velocity)
        self._check(self.L.flip_add_fluid_sdf(self.h, a.ctypes.data, lo, hi, v))

    def setSurfaceSubdivisionLevel(self, n):
        self._check(self.L.flip_set_surface_subdivision_level(self.h, int(n)))

    def setSurfaceSmoothing(self, value, iterations):
        self._check(self.L.flip_set_surface_smoothing(self.h, float(value), int(iterations)))

    def getIsomesh(self):
        """(vertices (nv,3) float32, triangles (nt,3) int32): FluidSimulation::getIsomesh, reconstructed on the device."""
        nv, nt = C.c_int(), C.c_int()
        self._check(self.L.flip_get_isomesh_size(self.h, C.byref(nv), C.byref(nt)))
        v = np.empty((nv.value, 3), dtype=np.float32)
        t = np.empty((nt.value, 3), dtype=np.int32)
        if nv.value or nt.value:
            self._check(self.L.flip_get_isomesh(self.h, v.ctypes.data, t.ctypes.data))
        return v, t

    def isomesh_field(self, subdivisions):
        """Parity seam: (values, inside, need) of the mesher's scalar field, each of shape (K s+1, J s+1, I s+1)."""
        I, J, K = self.dims
        shp = (K * subdivisions + 1, J * subdivisions + 1, I * subdivisions + 1)
        v = np.zeros(shp, dtype=np.float32)
        a = np.zeros(shp, dtype=np.uint8)
        b = np.zeros(shp, dtype=np.uint8)
        self._check(self.L.flip_get_isomesh_field(self.h, v.ctypes.data, a.ctypes.data, b.ctypes.data))
        return v, a, b

    def loadMarkerParticleData(self, data):
        if data.size == 0:
            return
        self._check(self.L.flip_load_particles(self.h, data.size, data.positions.ctypes.data, data.velocities.ctypes.data))

    def addMarkerParticle(self, p, v):
        a = (C.c_float * 3)(*p)
        b = (C.c_float * 3)(*v)
        self._check(self.L.flip_add_marker_particle(self.h, a, b))

    def setSolidSDF(self, phi_nodal):
        I, J, K = self.dims
        phi = np.ascontiguousarray(phi_nodal, dtype=np.float32)
        assert phi.size == (I + 1) * (J + 1) * (K + 1)
        self._check(self.L.flip_set_solid_sdf(self.h, phi.ctypes.data))

    def initialize(self):
        self._check(self.L.flip_initialize(self.h))

    # -- the hot path
    def update(self, dt):
        self._check(self.L.flip_update(self.h, float(dt)))

    def getCurrentFrame(self):
        v = C.c_int()
        self._check(self.L.flip_get_current_frame(self.h, C.byref(v)))
        return v.value

    def setCurrentFrame(self, frameno):
        self._check(self.L.flip_set_current_frame(self.h, int(frameno)))

    def getNumMarkerParticles(self):
        v = C.c_int()
        self._check(self.L.flip_get_num_particles(self.h, C.byref(v)))
        return v.value

    def getMarkerParticles(self, out=None):
        """(n,6) float32 {px,py,pz,vx,vy,vz}: std::vector<MarkerParticle> by value."""
        n = self.getNumMarkerParticles()
        a = out if out is not None else np.empty((n, 6), dtype=np.float32)
        self._check(self.L.flip_get_particles(self.h, a.ctypes.data, a.shape[0]))
        return a[:n]

    def setMarkerParticles(self, aos):
        aos = np.ascontiguousarray(aos, dtype=np.float32).reshape(-1, 6)
        self._check(self.L.flip_set_particles(self.h, aos.shape[0], aos.ctypes.data))

    def enableParticleIds(self, on=True, base=0):
        self._check(self.L.flip_enable_particle_ids(self.h, 1 if on else 0))
        self._check(self.L.flip_set_particle_id_base(self.h, int(base)))

    def getParticleIds(self):
        n = self.getNumMarkerParticles()
        a = np.empty(n, dtype=np.int32)
        self._check(self.L.flip_get_particle_ids(self.h, a.ctypes.data, n))
        return a

    def getMarkerParticlePositionData(self):
        n = self.getNumMarkerParticles()
        a = np.empty((n, 3), dtype=np.float32)
        self._check(self.L.flip_get_particle_positions(self.h, a.ctypes.data, n))
        return a

    def getMarkerParticleVelocityData(self):
        n = self.getNumMarkerParticles()
        a = np.empty((n, 3), dtype=np.float32)
        self._check(self.L.flip_get_particle_velocities(self.h, a.ctypes.data, n))
        return a

    def getVelocityField(self):
        """(U, V, W) host copies shaped (k,j,i) with the MACVelocityField extents."""
        U, V, W = (np.empty(self.shape_of(n), dtype=np.float32) for n in ("U", "V", "W"))
        self._check(self.L.flip_get_velocity_field(self.h, U.ctypes.data, V.ctypes.data, W.ctypes.data))
        return U, V, W

    # -- stats
    def substep_stats(self):
        n = C.c_int()
        self._check(self.L.flip_get_num_substeps(self.h, C.byref(n)))
        out = []
        for s in range(n.value):
            st = StepStats()
            self._check(self.L.flip_get_step_stats(self.h, s, C.byref(st)))
            out.append(st.as_dict())
        return out

    def stage_times_ms(self):
        a = (C.c_float * len(STAGES))()
        self._check(self.L.flip_get_stage_times_ms(self.h, a))
        return dict(zip(STAGES, list(a)))

    def enable_kernel_timing(self, on=True):
        self._check(self.L.flip_enable_kernel_timing(self.h, 1 if on else 0))

    def reset_kernel_timing(self):
        self._check(self.L.flip_reset_kernel_timing(self.h))

    def kernel_timing(self):
        """{class: (total_ms, launches)}"""
        out = {}
        for i, name in enumerate(KERNEL_CLASSES):
            ms, n = C.c_double(), C.c_int64()
            self._check(self.L.flip_get_kernel_timing(self.h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def kernel_launches(self):
        v = C.c_int64()
        self._check(self.L.flip_get_kernel_launches(self.h, C.byref(v)))
        return v.value

    def stream(self):
        v = C.c_void_p()
        self._check(self.L.flip_get_stream(self.h, C.byref(v)))
        return v.value

    def synchronize(self):
        self._check(self.L.flip_synchronize(self.h))

    # -- stage-wise seams
    def begin_frame(self, dt):
        self._check(self.L.flip_begin_frame(self.h, float(dt)))

    def begin_substep(self):
        v = C.c_double()
        self._check(self.L.flip_begin_substep(self.h, C.byref(v)))
        return v.value

    def stage(self, name, dt):
        self._check(self.L.flip_run_stage(self.h, STAGE_ID[name], float(dt)))

    def end_substep(self):
        v = C.c_int()
        self._check(self.L.flip_end_substep(self.h, C.byref(v)))
        return bool(v.value)

    def end_frame(self):
        self._check(self.L.flip_end_frame(self.h))

    def shape_of(self, name):
        I, J, K = self.dims
        if name != "solid_phi_global":
            K = self.local_K
        if name in ("U", "validU", "weightU", "savedU", "solidU"):
            return (K, J, I + 1)
        if name in ("V", "validV", "weightV", "savedV", "solidV"):
            return (K, J + 1, I)
        if name in ("W", "validW", "weightW", "savedW", "solidW"):
            return (K + 1, J, I)
        if name in ("liquid_phi", "pressure", "weightC"):
            return (K, J, I)
        if name == "solid_phi":
            return (K + 1, J + 1, I + 1)
        if name == "near_solid":
            b = C.c_int64()
            self._check(self.L.flip_array_bytes(self.h, ARRAY_ID[name], C.byref(b)))
            return (b.value,)
        raise KeyError(name)

    def array(self, name):
        dt = np.uint8 if name.startswith("valid") or name == "near_solid" else np.float32
        a = np.empty(self.shape_of(name), dtype=dt)
        b = C.c_int64()
        self._check(self.L.flip_array_bytes(self.h, ARRAY_ID[name], C.byref(b)))
        assert b.value == a.nbytes, (name, b.value, a.nbytes)
        self._check(self.L.flip_get_array(self.h, ARRAY_ID[name], a.ctypes.data))
        return a

    def set_array(self, name, a):
        dt = np.uint8 if name.startswith("valid") or name == "near_solid" else np.float32
        a = np.ascontiguousarray(a, dtype=dt)
        b = C.c_int64()
        self._check(self.L.flip_array_bytes(self.h, ARRAY_ID[name], C.byref(b)))
        assert b.value == a.nbytes, (name, b.value, a.nbytes)
        self._check(self.L.flip_set_array(self.h, ARRAY_ID[name], a.ctypes.data))
